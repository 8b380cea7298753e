"""Pre-processing of the raw frame: convertTo(float) -> BGR2GRAY -> GaussianBlur(5, sigma 3), MTF's default
utils::GaussianSmoothing (Utilities/src/preprocUtils.cc:108-127).

The arithmetic is OpenCV's, which MTF does not vendor: the oracle restates the OpenCV 2.4 / 3.x algorithm (the versions MTF
builds against, ReadMe.md:116) and is pinned here against outputs of the OpenCV that IS available in the build container
(cv2 4.13, tests/golden/make_preproc_golden.py).  OpenCV's SIMD builds fuse some multiply-adds and 4.x normalises the
kernel in double, so the anchor holds to a few ulp, not bit for bit:
  * kernel taps: within 1 ulp of cv2's             * filtered frame: within PREPROC_ULPS ulp of cv2's
The CUDA kernel against the oracle (same operation order, no contraction): bit-exact.
"""
import os

import numpy as np
import pytest

import common
from oracle import oracle_lib as O

PREPROC_ULPS = 4
GOLDEN = os.path.join(common.HERE, "golden", "preproc_gauss5.npz")


def _ulps(a, b):
    return (np.abs(a.astype(np.float64) - b) / np.spacing(np.abs(b).astype(np.float32))).max()


def test_oracle_kernel_and_frames_vs_opencv_golden():
    z = np.load(GOLDEN)
    k = O.gaussian_kernel5(float(z["sigma"]))
    assert _ulps(k, z["kernel"]) <= 1 and abs(float(k.sum(dtype=np.float64)) - 1.0) < 1e-7
    assert np.array_equal(k, k[::-1])
    for name in ("bgr", "gray", "tiny"):
        got = O.preproc_gauss5(z[name], float(z["sigma"]))
        assert got.dtype == np.float32 and got.shape == z[name + "_out"].shape
        assert _ulps(got, z[name + "_out"]) <= PREPROC_ULPS, name


def test_oracle_preproc_properties():
    """size-independent properties: a constant frame stays constant (taps sum to 1 within rounding), mirror symmetry, and
    gray == BGR with three equal channels up to the float gray weights"""
    rng = np.random.default_rng(5)
    const = np.full((17, 23), 200, dtype=np.uint8)
    assert np.abs(O.preproc_gauss5(const) - 200.0).max() <= 200 * 3 * 2.0 ** -23
    img = rng.integers(0, 256, size=(31, 40), dtype=np.uint8)
    a = O.preproc_gauss5(img)
    assert np.array_equal(O.preproc_gauss5(img[:, ::-1].copy()), a[:, ::-1])        # symmetric taps, symmetric border rule
    assert np.array_equal(O.preproc_gauss5(img[::-1].copy()), a[::-1])
    bgr = np.repeat(img[..., None], 3, axis=2)
    assert np.abs(O.preproc_gauss5(bgr) - a).max() <= 255 * 4 * 2.0 ** -23


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1024, 1024, 3), (1024, 1024), (37, 53, 3), (3, 3), (64, 16, 3), (17, 129)])
def test_gpu_preproc_bit_exact_vs_oracle(shape):
    from mtf_b200 import api
    rng = np.random.default_rng(sum(shape))
    img = rng.integers(0, 256, size=shape, dtype=np.uint8)
    g = api.BatchTracker(api.make_params("ssd", "homography", "fclk", n_patches=1))
    g.setRawImage(img)
    got = g.image(shape[0], shape[1])
    assert np.array_equal(got, O.preproc_gauss5(img))
    # a padded row stride must not matter
    wide = np.zeros((shape[0], shape[1] + 5) + shape[2:], dtype=np.uint8)
    wide[:, :shape[1]] = img
    g.setRawImage(wide[:, :shape[1]])
    assert np.array_equal(g.image(shape[0], shape[1]), got)


@pytest.mark.gpu
def test_gpu_tracking_from_raw_frames_equals_tracking_from_smoothed_frames(seq384):
    """setRawImage(uint8) == setImage(oracle-smoothed float frame), bit for bit, through initialize and update"""
    from mtf_b200 import api
    frames, _ = seq384
    raw = [np.clip(np.rint(f), 0, 255).astype(np.uint8) for f in frames[:3]]
    cs = common.patches(6, 52.3, 384, 384, seed=13)
    a = api.BatchTracker(api.make_params("ssd", "homography", "fclk", n_patches=len(cs)))
    b = api.BatchTracker(api.make_params("ssd", "homography", "fclk", n_patches=len(cs)))
    a.setRawImage(raw[0]); a.initialize(cs)
    b.initialize(cs, O.preproc_gauss5(raw[0]))
    for r in raw[1:]:
        a.setRawImage(r); a.update()
        b.update(O.preproc_gauss5(r))
        assert np.array_equal(a.getRegion(), b.getRegion())


@pytest.mark.gpu
def test_gpu_preproc_argument_errors():
    from mtf_b200 import api
    g = api.BatchTracker(api.make_params("ssd", "homography", "fclk", n_patches=1))
    img = np.zeros((8, 8), dtype=np.uint8)
    with pytest.raises(api.MTFError) as e:
        g.setRawImage(img, kernel_size=7)
    assert e.value.type == "FunctonNotImplemented"
    with pytest.raises(api.MTFError) as e:
        g.setRawImage(img, sigma=0.0)
    assert e.value.type == "InvalidArgument"
    with pytest.raises(api.MTFError):
        g.setRawImage(np.zeros((8, 8, 4), dtype=np.uint8))
