"""The C++ shim (include/mtf_b200_tracker.h) that subclasses mtf::TrackerBase.  MTF's own headers need Eigen / OpenCV /
Boost, which this image lacks, so the shim is compiled against stand-ins with the same interface (tests/shim_standin).
CPU: it compiles and links against the C-ABI library.  GPU: driven like runMTF / GridTracker drive a tracker, it returns
the corners the Python binding returns for the same inputs (both sit on the same C ABI)."""
import os
import struct
import subprocess

import numpy as np
import pytest

import common

ROOT = common.ROOT


def _build(tmp_path):
    exe = str(tmp_path / "shim_driver")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++11", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "tests", "shim_standin"),
                           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "shim_standin", "shim_driver.cpp"),
                           "-o", exe, "-L", os.path.join(ROOT, "mtf_b200"), "-lmtf_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "mtf_b200")])
    return exe


def test_shim_compiles_and_links(tmp_path):
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
@pytest.mark.parametrize("sm,am,ssm,res", [("fclk", "ssd", "8", 50), ("esm", "ncc", "6", 25)])
def test_shim_drives_the_tracker(tmp_path, seq384, sm, am, ssm, res):
    from mtf_b200 import api
    exe = _build(tmp_path)
    frames = np.stack(seq384[0][:3]).astype(np.float32)
    cs = common.patches(5, res - 1.0 if res == 50 else 24.6, 384, 384, seed=31)
    fb, cb = str(tmp_path / "frames.bin"), str(tmp_path / "corners.bin")
    with open(fb, "wb") as f:
        f.write(struct.pack("iii", *frames.shape)); f.write(frames.tobytes())
    with open(cb, "wb") as f:
        f.write(struct.pack("i", len(cs))); f.write(np.ascontiguousarray(cs, dtype=np.float64).tobytes())
    out = subprocess.check_output([exe, fb, cb, sm, am, ssm, str(res), "raw"], text=True).split("\n")
    vals = np.array([[float(x) for x in l.split()] for l in out if l and l[0] in "-0123456789"])
    single, member = vals[:, 0].reshape(len(cs), 2, 4), vals[:, 1].reshape(len(cs), 2, 4)
    assert "EXC InvalidArgument" in out
    assert not [l for l in out if l.startswith("STALE")], out
    hess = {"fclk": 1, "esm": 2}[sm]
    tr = api.BatchTracker(api.make_params(am, {"8": "homography", "6": "affine"}[ssm], sm, n_patches=len(cs), resx=res, resy=res,
                                          hess_type=hess))
    tr.initialize(cs, frames[0])
    for t in (1, 2):
        tr.update(frames[t])
    ref = tr.getRegion()
    # batch members share one context with the Python run (same work split); single-patch trackers use another thread
    # count per patch, i.e. another summation order
    assert np.abs(member - ref).max() <= 1e-9
    assert np.abs(single - ref).max() <= 1e-6
    # raw uint8 frames through Batch::useRawInput == the Python binding's setRawImage on the same bytes
    raw = np.array([float(l.split()[1]) for l in out if l.startswith("RAW ")]).reshape(2, 4)
    assert "BADTYPE" not in out
    u8 = [np.floor(np.clip(f, 0, 255) + 0.5).astype(np.uint8) for f in frames]
    t8 = api.BatchTracker(api.make_params(am, {"8": "homography", "6": "affine"}[ssm], sm, n_patches=1, resx=res, resy=res,
                                          hess_type=hess))
    t8.setRawImage(u8[0]); t8.initialize(cs[:1])
    for t in (1, 2):
        t8.setRawImage(u8[t]); t8.update()
    assert np.abs(raw - t8.getRegion()[0]).max() <= 1e-9
    # Batch::gridEstimate (points stay on the device) == Batch::estimateWarpFromPts on the same centroids from the host == the
    # oracle's estimateHomography on them
    from oracle import oracle_lib as O
    assert "EST 1 1" in out and not [l for l in out if l.startswith("ESTMASK")]
    su = np.array([[float(x) for x in l.split()[1:]] for l in out if l.startswith("ESTSU ")])
    assert np.array_equal(su[:, 0], su[:, 1])
    pts = np.array([[float(x) for x in l.split()[1:]] for l in out if l.startswith("ESTPT ")], dtype=np.float32)
    orc = O.estimate_warp("homography", pts[:, :2], pts[:, 2:], O.make_est_params("ransac", seed=99))
    assert orc["ok"] and np.allclose(su[:, 0], orc["state_update"], rtol=1e-5, atol=1e-6)
    # mtf::b200::GridTracker (cells + estimate + region + layout + re-initialisation behind the TrackerBase interface) == the Python
    # GridTracker on the same inputs (its first cell layout is NumPy's, the shim's the device's: 1e-9 px)
    from mtf_b200 import grid
    gl = np.array([[int(l.split()[1]), float(l.split()[2])] for l in out if l.startswith("GRID ")])
    gt = grid.GridTracker(api.make_params("ncc", "affine", "esm", n_patches=16, resx=12, resy=12, hess_type=2), grid_size_x=4, grid_size_y=4,
                          patch_size_x=20, patch_size_y=22, reset_at_each_frame=1, ssm="homography",
                          est_params=api.make_est_params("ransac", ransac_reproj_thresh=2.0), seed=5)
    gt.setImage(frames[0]); gt.initialize(np.array([[90.0, 300, 305, 85], [80, 84, 290, 296]]))
    for t in (1, 2):
        gt.setImage(frames[t])
        want = gt.update()
        got_t = gl[gl[:, 0] == t][:, 1].reshape(2, 4)
        assert np.abs(got_t - want).max() <= 1e-7
    gt.close()
    if ssm == "8":
        # mtf::b200::PFTracker == the Python binding's PFTracker with the same seed (device generator: deterministic)
        pfc = np.array([float(l.split()[1]) for l in out if l.startswith("PF ")]).reshape(2, 4)
        assert "EXCPF InvalidArgument" in out
        pt = api.PFTracker(api.make_params(am, "homography", "pf", n_patches=1, resx=res, resy=res), n_particles=300,
                           sigma=[0.5, 0.2, 0, 0, 0, 0, 0, 0], seed=77, mean_type="ssm")
        pt.initialize(cs[:1], frames[0])
        for t in (1, 2):
            pt.update(frames[t])
        assert np.abs(pfc - pt.getRegion()[0]).max() <= 1e-9
