"""CPU: the per-pixel arithmetic the CUDA kernels inline (mtf_b200/csrc/lk_math.cuh, compiled for the host by
tests/host_math) against the oracle.  Bit-exact where the reference's operation order is reproduced
(warped points, sampling indices, pixel values), tolerance-stated where the product evaluates the
eps -> 0 limit of the reference's finite-difference gradient."""
import numpy as np
import pytest

import common
from oracle import oracle_lib as O

# noise of the reference's own difference quotient (imgUtils.cc:233-254, eps = 1e-8): the quotient amplifies the
# ~1e-13 rounding error of two fp64 bilinear samples by 1/(2 eps) and x + eps is itself rounded to ~6e-14/1e-8
FD_RTOL, FD_ATOL = 2e-5, 5e-5


def _stage(ssm, img, xv, yv, dlt, W, eps=1e-8):
    L = common.host_math()
    S = 8 if ssm == "homography" else 6
    N = len(xv) * len(yv)
    ip = np.empty((N, 2)); pts = np.empty((N, 2)); It = np.empty(N); g = np.empty((2, N)); J = np.empty((S, N))
    xv = np.ascontiguousarray(xv); yv = np.ascontiguousarray(yv)
    dlt = np.ascontiguousarray(dlt, dtype=np.float64); W = np.ascontiguousarray(W, dtype=np.float64)
    L.hm_stage(0 if ssm == "homography" else 1, common.fptr(img), img.shape[0], img.shape[1], img.shape[1],
               common.dptr(xv), len(xv), common.dptr(yv), len(yv), common.dptr(dlt), common.dptr(W), eps,
               common.dptr(ip), common.dptr(pts), common.dptr(It), common.dptr(g), common.dptr(J))
    return ip, pts, It, g.T, J.T


def test_sample_pixel_bit_exact(seq384):
    img = seq384[0][0]
    rng = np.random.default_rng(3)
    pts = rng.uniform(-3, 387, size=(50000, 2))
    pts[:10000] = np.round(pts[:10000])                     # integer coordinates (dx == 0 rule, imgUtils.h:103-104)
    pts[10000:15000, 0] = np.round(pts[10000:15000, 0])
    pts[15000:20000, 1] = np.round(pts[15000:20000, 1])
    pts[20000:20010] = [[0, 0], [383, 383], [383, 0], [0, 383], [383.5, 10], [10, 383.5], [384, 5], [5, 384], [-0.0, 5], [382.999, 382.999]]
    ref = O.pix_vals(img, pts)
    L = common.host_math()
    val = np.empty(len(pts))
    L.hm_sample(common.fptr(img), 384, 384, 384, common.dptr(pts), len(pts), common.dptr(val))
    assert np.array_equal(val, ref)
    v2 = np.empty(len(pts)); gx = np.empty(len(pts)); gy = np.empty(len(pts))
    L.hm_sample_grad(common.fptr(img), 384, 384, 384, common.dptr(pts), len(pts), 1e-8, common.dptr(v2), common.dptr(gx), common.dptr(gy))
    assert np.array_equal(v2, ref)
    # gradient: bit-exact against the oracle's analytic mode, within difference-quotient noise of the reference mode
    ga = O.img_grad_analytic(img, pts)
    assert np.array_equal(np.c_[gx, gy], ga)
    gf = O.img_grad(img, pts)
    # skip points within eps of a cell boundary without being on it (measure-zero straddling, not emulated)
    frac = np.abs(pts - np.round(pts))
    ok = ~(((frac > 0) & (frac < 2e-8)).any(axis=1))
    err = np.abs(np.c_[gx, gy] - gf)[ok]
    assert (err <= FD_ATOL + FD_RTOL * np.abs(gf[ok])).all(), err.max()


@pytest.mark.parametrize("ssm", ["homography", "affine"])
@pytest.mark.parametrize("kind", ["axis", "quad"])
def test_stage_matches_oracle(seq384, ssm, kind):
    frames, _ = seq384
    cs = common.patches(4, 49.0, 384, 384) if kind == "axis" else common.quad_patches(4, 384, 384)
    res = 50
    lo, hi = (-0.5, 0.5) if ssm == "homography" else (1 - res / 2.0, res / 2.0)
    grid, _ = O.norm_unit_square_pts(res, res, lo, lo, hi, hi)
    xv = grid[:res, 0].copy(); yv = grid[::res, 1].copy()
    for c in cs:
        for gm in (1, 0):
            p = O.make_params("ssd", ssm, "fclk", max_iters=1, epsilon=-1.0, grad_mode=gm)
            tr = O.OracleTracker(p)
            tr.set_image(frames[0]); tr.initialize(c)
            dlt = tr.init_warp()
            ip, pts, I0, g0, J0 = _stage(ssm, frames[0], xv, yv, dlt, np.eye(3))
            assert np.array_equal(ip, tr.init_pts())
            assert np.array_equal(pts, tr.pts())
            assert np.array_equal(I0, tr.init_pix_vals())
            # one pass on the next frame: It / dIt_dx / dIt_dp of the oracle are those of the identity warp
            tr.set_image(frames[1]); tr.update()
            ip, pts, It, g, J = _stage(ssm, frames[1], xv, yv, dlt, np.eye(3))
            assert np.array_equal(It, tr.curr_pix_vals())
            if gm == 1:
                assert np.array_equal(g, tr.curr_pix_grad())
                assert np.array_equal(J, tr.curr_pix_jacobian())
            else:
                gr = tr.curr_pix_grad()
                assert (np.abs(g - gr) <= FD_ATOL + FD_RTOL * np.abs(gr)).all()
                Jr = tr.curr_pix_jacobian()
                scale = np.abs(Jr).max(axis=0)
                assert (np.abs(J - Jr) <= 1e-4 * scale).all()
            # second pass: a non-identity warp
            W = common.warp_from_state(ssm, tr.state())
            tr.update()
            ip, pts, It, g, J = _stage(ssm, frames[1], xv, yv, dlt, W)
            assert np.array_equal(It, tr.curr_pix_vals())
            if gm == 1:
                assert np.array_equal(g, tr.curr_pix_grad())
                assert np.array_equal(J, tr.curr_pix_jacobian())


@pytest.mark.parametrize("ssm", ["homography", "affine"])
def test_ssm_algebra_matches_oracle(seq384, ssm):
    """compositionalUpdate / invertState / corners through the oracle's public behaviour (ICLK takes both)."""
    frames, _ = seq384
    L = common.host_math()
    sid = 0 if ssm == "homography" else 1
    c = common.quad_patches(1, 384, 384)[0]
    p = O.make_params("ssd", ssm, "iclk", max_iters=1, epsilon=-1.0)
    tr = O.OracleTracker(p)
    tr.set_image(frames[0]); tr.initialize(c)
    tr.set_image(frames[1])
    W = np.eye(3)
    for _ in range(3):
        tr.update()
        dp = tr.log()[-1]["state_update"].copy()
        inv = np.empty(len(dp)); L.hm_invert_state(sid, common.dptr(dp), common.dptr(inv))
        Wn = np.empty((3, 3)); L.hm_compose(sid, common.dptr(np.ascontiguousarray(W)), common.dptr(inv), common.dptr(Wn))
        W = Wn
        assert np.array_equal(common.warp_from_state(ssm, tr.state()), W)
        out = np.empty(8); L.hm_warp_corners(sid, common.dptr(W), common.dptr(np.ascontiguousarray(c).reshape(8)), common.dptr(out))
        assert np.array_equal(out.reshape(2, 4), tr.corners())
