"""CPU: pins the oracle -- against an independent NumPy restatement of the leaf functions (tests/np_ref.py), against
LAPACK for the Eigen algorithms it restates, and through the reference's own Diagnostics identities
(Diagnostics/src/DiagNumeric.cc:6-113 finite-difference Jacobians, Diagnostics.cc:131-166 Hessian equalities)."""
import numpy as np
import pytest

import common
import np_ref
from oracle import oracle_lib as O


def test_bilinear_sampling_matches_numpy(seq384):
    img = seq384[0][0]
    rng = np.random.default_rng(1)
    pts = rng.uniform(-4, 388, size=(40000, 2))
    pts[:8000] = np.round(pts[:8000]); pts[8000:12000, 0] = np.round(pts[8000:12000, 0])
    pts[12000:12006] = [[0, 0], [383, 383], [383.5, 4], [4, 383.2], [384, 384], [-1e-9, 3]]
    assert np.array_equal(O.pix_vals(img, pts), np_ref.pix_vals(img, pts))
    g = O.img_grad(img, pts[:5000])
    assert np.array_equal(g, np_ref.img_grad(img, pts[:5000]))


def test_dlt_matches_svd():
    rng = np.random.default_rng(2)
    unit = np.array([[-0.5, 0.5, 0.5, -0.5], [-0.5, -0.5, 0.5, 0.5]])
    for _ in range(50):
        c = np.array([[100, 160, 165, 95], [80, 75, 140, 150.0]]) + rng.uniform(-10, 10, (2, 4))
        H = O.homography_dlt(unit, c)
        assert np.allclose(H, np_ref.homography_dlt(unit, c), rtol=1e-8, atol=1e-9)
        m = H @ np.vstack([unit, np.ones(4)])
        assert np.allclose(m[:2] / m[2], c, atol=1e-9)


def test_colpiv_qr_solve_matches_lapack():
    rng = np.random.default_rng(3)
    for n in (6, 8):
        for _ in range(30):
            J = rng.normal(size=(40, n)) * rng.uniform(0.1, 30, size=n)
            A = -(J.T @ J); b = rng.normal(size=n)
            assert np.allclose(O.colpiv_qr_solve(A, b), np.linalg.solve(A, b), rtol=1e-8, atol=1e-12)
    A = np.zeros((8, 8)); A[:3, :3] = [[4, 1, 0], [1, 3, 1], [0, 1, 2]]          # rank 3: Eigen zeroes the rest
    x = O.colpiv_qr_solve(A, np.r_[1, 2, 3, 0, 0, 0, 0, 0.0])
    assert np.allclose(x[:3], np.linalg.solve(A[:3, :3], [1, 2, 3])) and np.all(x[3:] == 0)


def test_grid_is_linspaced():
    for res, lo, hi in ((50, -0.5, 0.5), (10, 1 - 5.0, 5.0), (25, 1 - 12.5, 12.5)):
        pts, corners = O.norm_unit_square_pts(res, res, lo, lo, hi, hi)
        assert np.allclose(pts[:res, 0], np.linspace(lo, hi, res), rtol=0, atol=2e-16 * max(abs(lo), abs(hi)) * res)
        assert pts[0, 0] == lo and pts[res - 1, 0] == hi and pts[-1, 1] == hi
        assert np.array_equal(corners, [[lo, hi, hi, lo], [lo, lo, hi, hi]])


@pytest.mark.parametrize("am", ["ssd", "ncc"])
@pytest.mark.parametrize("ssm", ["homography", "affine"])
def test_similarity_matches_numpy(seq384, am, ssm):
    """f at a perturbed state, computed from scratch with NumPy: warp the grid, sample, compare"""
    frames, _ = seq384
    c = common.quad_patches(1, 384, 384, seed=21)[0]
    o = O.OracleTracker(O.make_params(am, ssm, "fclk"))
    o.set_image(frames[0]); o.initialize(c); o.set_image(frames[1])
    S = o.S
    rng = np.random.default_rng(4)
    scale = np.array([1e-2, 1e-2, 1.0, 1e-2, 1e-2, 1.0, 1e-5, 1e-5]) if S == 8 else np.array([1.0, 1.0, 1e-2, 1e-2, 1e-2, 1e-2])
    states = rng.normal(size=(5, S)) * scale
    _, sim = o.pf_evaluate(states)
    ip = o.init_pts(); I0 = o.init_pix_vals()
    dlt = o.init_warp()
    res = 50
    lo, hi = (-0.5, 0.5) if ssm == "homography" else (1 - res / 2.0, res / 2.0)
    grid, _ = O.norm_unit_square_pts(res, res, lo, lo, hi, hi)
    hm0 = dlt @ np.vstack([grid.T, np.ones(len(grid))])
    assert np.allclose((hm0[:2] / hm0[2]).T, ip, rtol=1e-14)
    for s, f in zip(states, sim):
        W = np_ref.warp_from_state(ssm, s)
        hm = W @ (hm0 if ssm == "homography" else np.vstack([ip.T, np.ones(len(ip))]))
        pts = (hm[:2] / hm[2]).T
        It = np_ref.pix_vals(frames[1], pts)
        ref = np_ref.ssd(I0, It) if am == "ssd" else np_ref.ncc(I0, It)
        assert abs(f - ref) <= 1e-9 * max(1.0, abs(ref))


@pytest.mark.parametrize("am", ["ssd", "ncc", "mi"])
@pytest.mark.parametrize("ssm", ["homography", "affine"])
def test_jacobian_is_the_derivative_of_f(seq384, am, ssm):
    """DiagNumeric.cc:6-113: the analytic df/dp of the first FCLK pass = central difference of f along each parameter"""
    frames, _ = seq384
    c = common.patches(1, 52.3, 384, 384, seed=22)[0]
    kw = {"hess_type": 0} if am == "mi" else {}
    o = O.OracleTracker(O.make_params(am, ssm, "fclk", max_iters=1, **kw))
    o.set_image(frames[0]); o.initialize(c); o.set_image(frames[1]); o.update()
    J = o.log()[0]["jacobian"]
    o2 = O.OracleTracker(O.make_params(am, ssm, "fclk", **kw))
    o2.set_image(frames[0]); o2.initialize(c); o2.set_image(frames[1])
    S = o.S
    h = np.array([1e-6, 1e-6, 1e-4, 1e-6, 1e-6, 1e-4, 1e-9, 1e-9]) if S == 8 else np.array([1e-4, 1e-4, 1e-6, 1e-6, 1e-6, 1e-6])
    states = np.concatenate([np.diag(h), -np.diag(h)])
    _, f = o2.pf_evaluate(states)
    fd = (f[:S] - f[S:]) / (2 * h)
    assert np.allclose(fd, J, rtol=2e-3, atol=2e-3 * np.abs(J).max())


@pytest.mark.parametrize("ssm", ["homography", "affine"])
def test_hessians_agree_at_identity(seq384, ssm):
    """Diagnostics::validateHessians (Diagnostics.cc:131-166): with the current image = the template image and the identity
    warp, Self(init) == Self(curr) == Std for SSD"""
    frames, _ = seq384
    c = common.patches(1, 52.3, 384, 384, seed=23)[0]
    H = {}
    for name, sm, ht in (("init_self", "esm", 0), ("curr_self", "fclk", 1), ("std", "fclk", 2)):
        o = O.OracleTracker(O.make_params("ssd", ssm, sm, hess_type=ht, max_iters=1))
        o.set_image(frames[0]); o.initialize(c); o.update()
        H[name] = o.log()[0]["hessian"]
    assert np.allclose(H["init_self"], H["curr_self"], rtol=1e-12) and np.array_equal(H["curr_self"], H["std"])
    assert np.all(np.linalg.eigvalsh(-(H["std"] + H["std"].T) / 2) > 0)          # -J^T J is negative definite


@pytest.mark.parametrize("am,sm", [("ssd", "esm"), ("ssd", "fclk"), ("ssd", "iclk"), ("ncc", "esm"), ("ncc", "fclk")])
def test_tracking_recovers_ground_truth(seq384, am, sm):
    from mtf_b200 import synth
    frames, warps = seq384
    cs = common.patches(4, 52.3, 384, 384, seed=24)
    for c in cs:
        o = O.OracleTracker(O.make_params(am, "homography", sm))
        o.set_image(frames[0]); o.initialize(c)
        for t in (1, 2, 3):
            o.set_image(frames[t]); o.update()
            assert np.abs(o.corners() - synth.warp_corners(warps[t], c)).max() < 0.25
            assert o.n_iters < 30


def test_colpiv_qr_pivots_and_R_match_lapack_dgeqp3():
    """Eigen's ColPivHouseholderQR and LAPACK's dgeqp3 are the same algorithm (largest remaining column norm, norm
    down-dating): same pivot order, R equal up to the sign convention of each row"""
    import scipy.linalg as sl
    rng = np.random.default_rng(31)
    for n in (6, 8):
        for _ in range(40):
            J = rng.normal(size=(60, n)) * rng.uniform(0.1, 30, size=n)
            A = -(J.T @ J)
            R, perm, _, nz = O.colpiv_qr(A)
            _, R2, P2 = sl.qr(A, pivoting=True)
            assert nz == n and np.array_equal(perm, P2)
            assert np.allclose(np.abs(R), np.abs(R2), rtol=1e-10, atol=1e-10 * np.abs(R2).max())
    # the 9 x 8 adjoint of the DLT constraint matrix (warpUtils.cc:171-223 through JacobiSVD's QR preconditioner)
    A = rng.normal(size=(9, 8))
    R, perm, _, nz = O.colpiv_qr(A)
    _, R2, P2 = sl.qr(A, pivoting=True)
    assert nz == 8 and np.array_equal(perm, P2) and np.allclose(np.abs(R[:8]), np.abs(R2[:8]), rtol=1e-10)


def test_colpiv_qr_rank_threshold_is_eigens():
    """Eigen 3.3 ColPivHouseholderQR::computeInPlace: threshold_helper = abs2(maxColNorm * epsilon) / rows, and pivot k counts
    as zero when its squared norm < threshold_helper * (rows - k).  A diagonal matrix whose last entry t sits between
    eps / rows (a misreading with the division inside the square) and eps / sqrt(rows) tells the two apart."""
    eps = np.finfo(float).eps
    for n in (6, 8):
        b = np.arange(1.0, n + 1)
        for t, rank in ((eps / np.sqrt(n) * 1.05, n), (eps / np.sqrt(n) * 0.95, n - 1), (eps / n * 1.5, n - 1), (eps / n * 0.5, n - 1)):
            A = np.diag(np.r_[np.ones(n - 1), t])
            R, perm, _, nz = O.colpiv_qr(A)
            assert nz == rank, (n, t, nz)
            x = O.colpiv_qr_solve(A, b)
            assert np.allclose(x[:n - 1], b[:n - 1])
            assert (x[n - 1] == 0.0) if rank < n else np.isclose(x[n - 1], b[n - 1] / t)


def test_dlt_matches_opencv():
    """the 4-point DLT against cv2.getPerspectiveTransform (an independent 8 x 8 linear solve)"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(32)
    unit = np.array([[-0.5, 0.5, 0.5, -0.5], [-0.5, -0.5, 0.5, 0.5]])
    for _ in range(50):
        c = np.array([[300, 360, 365, 295], [480, 475, 540, 550.0]]) + rng.uniform(-10, 10, (2, 4))
        H = O.homography_dlt(unit, c)
        H2 = cv2.getPerspectiveTransform(unit.T.astype(np.float32), c.T.astype(np.float32))
        # float32 corners on OpenCV's side: compare through the corners they map to
        m = H2 @ np.vstack([unit, np.ones(4)])
        assert np.allclose(m[:2] / m[2], c, atol=2e-4)
        assert np.allclose(H / H[2, 2], H2 / H2[2, 2], rtol=2e-5, atol=2e-5 * np.abs(H).max())


def test_affine_ndlt_against_svd_least_squares():
    """utils::computeAffineNDLT (warpUtils.cc:378-386) as the oracle restates it (Householder QR) against the reference's own
    formulation -- normalizePts, the 8 x 6 system solved through an SVD pseudo-inverse (numpy.linalg.lstsq), inv_norm_mat * warp"""
    from oracle import oracle_lib as O
    rng = np.random.default_rng(3)
    for res in (10, 25, 50):
        inc = np.array([[1 - res / 2, res / 2, res / 2, 1 - res / 2], [1 - res / 2, 1 - res / 2, res / 2, res / 2]])
        for _ in range(50):
            out = np.array([[100, 150, 150, 100], [200, 200, 250, 250.0]]) + rng.uniform(-8, 8, (2, 4)) + rng.uniform(0, 500)
            H = O.affine_ndlt(inc, out)
            c = out.mean(axis=1, keepdims=True); t = out - c
            s = np.sqrt(2) / np.sqrt((t ** 2).sum(axis=0)).mean(); n = t * s
            A = np.zeros((8, 6)); b = np.zeros(8)
            for i in range(4):
                A[2 * i, :3] = [inc[0, i], inc[1, i], 1]; A[2 * i + 1, 3:] = [inc[0, i], inc[1, i], 1]
                b[2 * i] = n[0, i]; b[2 * i + 1] = n[1, i]
            x = np.linalg.lstsq(A, b, rcond=None)[0]
            R = np.array([[1 / s, 0, c[0, 0]], [0, 1 / s, c[1, 0]], [0, 0, 1]]) @ np.array([[x[0], x[1], x[2]], [x[3], x[4], x[5]], [0, 0, 1]])
            assert np.abs(H - R).max() <= 1e-14 * np.abs(R).max()
            assert H[2, 0] == 0 and H[2, 1] == 0 and H[2, 2] == 1
