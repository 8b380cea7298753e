"""Mint tests/golden/preproc_gauss5.npz from OpenCV itself (cv2, present in the build container only): the arithmetic of
MTF's default pre-processing -- PreProcBase::processFrame with GaussianSmoothing, Utilities/src/preprocUtils.cc:108-127 --
lives in OpenCV, which MTF does not vendor.  Run here:  python tests/golden/make_preproc_golden.py"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20261017)
    # smooth structure + noise, so that neighbouring pixels differ by realistic amounts; odd sizes exercise the borders
    yy, xx = np.mgrid[0:45, 0:67]
    base = 128 + 80 * np.sin(xx / 7.0) * np.cos(yy / 5.0)
    bgr = np.clip(base[..., None] + rng.normal(0, 25, size=(45, 67, 3)), 0, 255).astype(np.uint8)
    gray = bgr[..., 1].copy()
    out = {"cv2_version": cv2.__version__, "sigma": 3.0, "bgr": bgr, "gray": gray,
           "kernel": cv2.getGaussianKernel(5, 3.0, cv2.CV_32F).ravel()}
    f = bgr.astype(np.float32)                                         # frame_raw.convertTo(frame_rgb, CV_32FC3)
    g = cv2.cvtColor(f, cv2.COLOR_BGR2GRAY)                            # cv::cvtColor(frame_rgb, frame_gs, CV_BGR2GRAY)
    out["bgr_gray_f32"] = g
    out["bgr_out"] = cv2.GaussianBlur(g, (5, 5), 3.0, sigmaY=3.0)      # GaussianSmoothing::apply
    out["gray_out"] = cv2.GaussianBlur(gray.astype(np.float32), (5, 5), 3.0, sigmaY=3.0)
    tiny = rng.integers(0, 256, size=(3, 4), dtype=np.uint8)           # every pixel is a border pixel
    out["tiny"] = tiny
    out["tiny_out"] = cv2.GaussianBlur(tiny.astype(np.float32), (5, 5), 3.0, sigmaY=3.0)
    np.savez_compressed(os.path.join(HERE, "preproc_gauss5.npz"), **out)
    print("wrote preproc_gauss5.npz, cv2", cv2.__version__)


if __name__ == "__main__":
    main()
