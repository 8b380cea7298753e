"""GPU: parity against the CPU oracle ON THE BENCHMARK'S OWN INPUTS (mtf_b200/workloads.py = what bench.py times):
1024 x 1024 frames (seeds 1234 / 5678), the seed-42 lattice of 1024 integer-aligned 49 px boxes, hom_normalized_init = 0,
epsilon = 0, 30 passes per frame, frames 1..3 -- and the other BASELINE configurations at their own sizes.

What can and cannot agree here.  With raw pixel coordinates the 8 x 8 Hessian has a condition number
of 1e14..1e17, so Eigen's rank rule (ColPivHouseholderQR: pivot k is zero when its squared norm < (max norm eps)^2 / rows
* (rows - k)) fires or not on rounding noise: the reference's own trajectory is chaotic at the level of which passes take
a rank-truncated step.  Both sides take 30 passes towards the same minimum of the SSD; what is compared is therefore
  * the first pass of every frame from an IDENTICAL state (set by setRegion on both sides): f / J^T r / J^T J to summation
    order (F64) or fp32 tolerance (F32), and
  * the corners after the 30 passes of each frame: the stated per-configuration tolerances below, measured on the B200 and
    recorded by this test in gpurun_out/bench_parity_stats.json (printed with -s).
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle_lib as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SAMPLE = np.arange(0, 1024, 16)            # 64 of the 1024 bench patches, spread over the lattice

# ---- stated tolerances (px unless noted), bench workload, corners after each frame's 30 passes
TOL = {
    # F64 kernel vs the oracle in the same gradient mode: identical arithmetic up to summation order; the residue is the
    # rank rule deciding differently on some passes (see the module docstring)
    "f64_vs_oracle_gm1": 5e-3,
    # vs the reference's finite-difference gradient (grad_mode = 0): its quotient's own noise moves the fixed point
    "f64_vs_oracle_gm0": 2e-2,
    # F32 per-pixel arithmetic, reference-basis QR (MTFB_F32_SOLVE_REFERENCE)
    "f32_reference_vs_oracle_gm1": 2e-2,
    # F32, local-basis Gauss-Jordan (MTFB_F32_SOLVE_LOCAL, NOT the default): the full Gauss-Newton step on every pass, where
    # the reference truncates the step on ~900 of the 1024 patches (rank rule).  The two then stop at different points of
    # the same valley: measured 0.044 / 0.060 / 0.036 px (max), 0.007 px (median) on frames 1..3, with the oracle itself
    # 0.067 .. 0.083 px (max) from the synthetic ground truth.  An explicit opt-in, never a parity claim.
    "f32_local_vs_oracle_gm1": 1e-1,
    "f32_local_median": 1.5e-2,
    # median over the sampled patches (the bulk agrees far better than the worst patch)
    "median": 2e-3,
}
STATS = {}


def _record(name, d):
    STATS[name] = d
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "bench_parity_stats.json"), "w") as f:
            json.dump(STATS, f, indent=1, sort_keys=True)
    except Exception:
        pass
    print(name, json.dumps(d))


def _dist(diff):
    d = np.abs(diff).reshape(diff.shape[0], -1).max(axis=1)
    return {"median": float(np.median(d)), "p90": float(np.percentile(d, 90)), "max": float(d.max())}


@pytest.fixture(scope="module")
def bench_inputs():
    from mtf_b200 import workloads
    frames, warps = workloads.sequence()
    return frames, warps, workloads.config2_patches()


@pytest.fixture(scope="module")
def oracle_tracks(bench_inputs):
    """the oracle's trajectories of the sampled patches through frames 1..3, both gradient modes"""
    frames, _, corners = bench_inputs
    out = {}
    for gm in (1, 0):
        tr = np.empty((3, len(SAMPLE), 2, 4)); nit = np.empty((3, len(SAMPLE)), dtype=int)
        for k, i in enumerate(SAMPLE):
            o = O.OracleTracker(O.make_params("ssd", "homography", "fclk", max_iters=30, epsilon=0.0, grad_mode=gm))
            o.set_image(frames[0]); o.initialize(corners[i])
            for t in range(3):
                o.set_image(frames[1 + t]); o.update()
                tr[t, k] = o.corners(); nit[t, k] = o.n_iters
        out[gm] = (tr, nit)
    return out


def _gpu_track(bench_inputs, **kw):
    from mtf_b200 import api
    frames, _, corners = bench_inputs
    g = api.BatchTracker(api.make_params("ssd", "homography", "fclk", n_patches=len(corners), max_iters=30, epsilon=0.0, **kw))
    g.initialize(corners, frames[0])
    tr = np.empty((3, len(SAMPLE), 2, 4))
    for t in range(3):
        g.update(frames[1 + t])
        tr[t] = g.getRegion()[SAMPLE]
    assert np.isfinite(g.getRegion()).all()
    return tr, g


@pytest.mark.parametrize("arm", ["f64", "f32_reference", "f32_local"])
def test_bench_workload_corners_vs_oracle(bench_inputs, oracle_tracks, arm):
    from mtf_b200 import synth
    frames, warps, corners = bench_inputs
    kw = {"f64": {}, "f32_reference": dict(precision="f32", f32_solve="reference"),
          "f32_local": dict(precision="f32", f32_solve="local")}[arm]
    got, g = _gpu_track(bench_inputs, **kw)
    stats = {}
    for gm in (1, 0):
        want = oracle_tracks[gm][0]
        per_frame = [_dist(got[t] - want[t]) for t in range(3)]
        stats["vs_oracle_gm%d" % gm] = per_frame
    truth = np.stack([synth.warp_corners(warps[1 + t], corners[SAMPLE]) for t in range(3)])
    stats["vs_truth"] = [_dist(got[t] - truth[t]) for t in range(3)]
    stats["oracle_gm0_vs_truth"] = [_dist(oracle_tracks[0][0][t] - truth[t]) for t in range(3)]
    stats["rank_deficient_patches"] = int((g.patch_status() & 2 != 0).sum())
    _record("config2_%s" % arm, stats)
    key = {"f64": "f64_vs_oracle_gm1", "f32_reference": "f32_reference_vs_oracle_gm1", "f32_local": "f32_local_vs_oracle_gm1"}[arm]
    for t in range(3):
        assert stats["vs_oracle_gm1"][t]["max"] <= TOL[key], (arm, t, stats["vs_oracle_gm1"][t])
        assert stats["vs_oracle_gm1"][t]["median"] <= TOL["f32_local_median" if arm == "f32_local" else "median"], \
            (arm, t, stats["vs_oracle_gm1"][t])
        if arm == "f64":
            assert stats["vs_oracle_gm0"][t]["max"] <= TOL["f64_vs_oracle_gm0"], (t, stats["vs_oracle_gm0"][t])


@pytest.mark.parametrize("arm", ["f64", "f32"])
def test_bench_workload_first_pass_from_identical_state(bench_inputs, arm):
    """frame 1, first pass, all 64 sampled patches at the state initialize() left: sums and the state update"""
    from mtf_b200 import api
    frames, _, corners = bench_inputs
    cs = corners[SAMPLE]
    kw = dict(precision="f32") if arm == "f32" else {}
    g = api.BatchTracker(api.make_params("ssd", "homography", "fclk", n_patches=len(cs), max_iters=30, epsilon=0.0, **kw))
    g.initialize(cs, frames[0])
    g.setImage(frames[1])
    J, H, f, dp = g.iterate_once()
    worst = {"f": 0.0, "J": 0.0, "H": 0.0, "corners": 0.0}
    cg = g.getRegion()
    for k, c in enumerate(cs):
        o = O.OracleTracker(O.make_params("ssd", "homography", "fclk", max_iters=1, epsilon=0.0, grad_mode=1))
        o.set_image(frames[0]); o.initialize(c); o.set_image(frames[1]); o.update()
        e = o.log()[0]
        worst["f"] = max(worst["f"], abs(f[k] - e["f"]) / abs(e["f"]))
        worst["J"] = max(worst["J"], np.abs(J[k] - e["jacobian"]).max() / np.abs(e["jacobian"]).max())
        worst["H"] = max(worst["H"], np.abs(H[k] - e["hessian"]).max() / np.abs(e["hessian"]).max())
        worst["corners"] = max(worst["corners"], np.abs(cg[k] - o.corners()).max())
    _record("config2_first_pass_%s" % arm, worst)
    if arm == "f64":
        assert worst["f"] <= 1e-12 and worst["J"] <= 1e-11 and worst["H"] <= 1e-12, worst
    else:
        assert worst["f"] <= 2e-5 and worst["J"] <= 2e-4 and worst["H"] <= 2e-5, worst


@pytest.mark.parametrize("arm", ["f64", "f32_reference", "f32_local"])
def test_bench_workload_iteration_counts(bench_inputs, arm):
    """the reference's stopping rule (epsilon = 1e-4) on the bench patches, against the reference's own gradient mode"""
    from mtf_b200 import api
    frames, _, corners = bench_inputs
    kw = {"f64": {}, "f32_reference": dict(precision="f32", f32_solve="reference"),
          "f32_local": dict(precision="f32", f32_solve="local")}[arm]
    g = api.BatchTracker(api.make_params("ssd", "homography", "fclk", n_patches=len(corners), max_iters=30, epsilon=1e-4, **kw))
    g.initialize(corners, frames[0])
    g.update(frames[1])
    n_gpu, c_gpu = g.n_iters()[SAMPLE], g.getRegion()[SAMPLE]
    n_ref = np.empty(len(SAMPLE), dtype=int); c_ref = np.empty((len(SAMPLE), 2, 4))
    for k, i in enumerate(SAMPLE):
        o = O.OracleTracker(O.make_params("ssd", "homography", "fclk", max_iters=30, epsilon=1e-4, grad_mode=0))
        o.set_image(frames[0]); o.initialize(corners[i]); o.set_image(frames[1]); o.update()
        n_ref[k] = o.n_iters; c_ref[k] = o.corners()
    dn = np.abs(n_gpu.astype(int) - n_ref)
    st = {"n_iters_ref_mean": float(n_ref.mean()), "n_iters_gpu_mean": float(n_gpu.mean()), "dn_max": int(dn.max()),
          "dn_le1_frac": float((dn <= 1).mean()), "corners": _dist(c_gpu - c_ref)}
    _record("config2_eps1e-4_%s" % arm, st)
    # the last accepted step is ~1e-2 px (epsilon = 1e-4 on the squared corner change): corners agree to a few of those
    assert st["corners"]["max"] <= 5e-2, st
    assert st["dn_le1_frac"] >= 0.8 and st["dn_max"] <= 6, st


# ------------------------------------------------------------------------------------------------ ADVICE round 1
@pytest.mark.parametrize("sm", ["iclk", "esm"])
@pytest.mark.parametrize("threads", [32, 0])
def test_stored_hessian_with_one_warp_per_patch(bench_inputs, sm, threads):
    """init_self_hessian must be complete when 32 threads write the 64 entries (ssd_init_kernel): ICLK InitialSelf and ESM
    SumOfSelf (the defaults) on a >= 900-patch batch, where one warp per patch is the automatic F64 split"""
    from mtf_b200 import api
    frames, _, corners = bench_inputs
    g = api.BatchTracker(api.make_params("ssd", "homography", sm, n_patches=len(corners), max_iters=30, epsilon=0.0,
                                         threads_per_patch=threads))
    g.initialize(corners, frames[0])
    g.update(frames[1])
    got = g.getRegion()
    assert np.isfinite(got).all()
    worst = 0.0
    for i in SAMPLE[:16]:
        o = O.OracleTracker(O.make_params("ssd", "homography", sm, max_iters=30, epsilon=0.0, grad_mode=1))
        o.set_image(frames[0]); o.initialize(corners[i]); o.set_image(frames[1]); o.update()
        worst = max(worst, np.abs(got[i] - o.corners()).max())
    _record("stored_hessian_%s_T%d" % (sm, threads), {"max": worst})
    assert worst <= TOL["f64_vs_oracle_gm1"], worst


def test_pf_single_object_odd_resolution(bench_inputs):
    """n_patches = 1 and an odd N put the template row of the bulk copy on an 8-byte boundary only"""
    from mtf_b200 import api, workloads
    frames, _, corners = bench_inputs
    for res, P in ((25, 1), (25, 3), (50, 1)):
        cs = corners[100:100 + P]
        g = api.BatchTracker(api.make_params("ssd", "homography", "pf", n_patches=P, resx=res, resy=res))
        g.initialize(cs, frames[0]); g.setImage(frames[1])
        states = workloads.config5_states(P, 64, seed=3)
        lik, sim = g.pf_evaluate(states)
        for k in range(P):
            o = O.OracleTracker(O.make_params("ssd", "homography", "fclk", resx=res, resy=res))
            o.set_image(frames[0]); o.initialize(cs[k]); o.set_image(frames[1])
            l2, s2 = o.pf_evaluate(states[k])
            assert np.allclose(sim[k], s2, rtol=1e-12) and np.allclose(lik[k], l2, rtol=1e-12)


def test_device_qr_rank_threshold_is_eigens():
    """the device QR (both the literal and the tuned per-pass variant) applies Eigen's rank rule: (max norm eps)^2 / rows"""
    from mtf_b200 import api
    eps = np.finfo(float).eps
    for n in (6, 8):
        ts = np.array([eps / np.sqrt(n) * 1.05, eps / np.sqrt(n) * 0.95, eps / n * 1.5, eps / n * 0.5])
        ranks = [n, n - 1, n - 1, n - 1]
        A = np.stack([np.diag(np.r_[np.ones(n - 1), t]) for t in ts])
        b = np.tile(np.arange(1.0, n + 1), (len(ts), 1))
        for fast in (0, 1, 2, 3):
            x, nz, perm = api.debug_colpiv_qr_solve(A, b, fast=fast)
            assert list(nz) == ranks, (n, fast, nz)
            for k, t in enumerate(ts):
                want = O.colpiv_qr_solve(A[k], b[k])
                assert np.allclose(x[k], want, rtol=1e-12, atol=0), (n, fast, k)
    # generic SPD systems: same pivot order as the oracle (= LAPACK dgeqp3, tests/test_oracle.py), same solution
    rng = np.random.default_rng(5)
    J = rng.normal(size=(64, 60, 8)) * rng.uniform(0.1, 30, size=(64, 1, 8))
    A = -np.einsum("kni,knj->kij", J, J); b = rng.normal(size=(64, 8))
    for fast in (0, 1, 2, 3):
        x, nz, perm = api.debug_colpiv_qr_solve(A, b, fast=fast)
        for k in range(64):
            _, p2, _, nz2 = O.colpiv_qr(A[k])
            assert nz[k] == nz2 and np.array_equal(perm[k], p2)
            assert np.allclose(x[k], O.colpiv_qr_solve(A[k], b[k]), rtol=1e-9)


# ------------------------------------------------------------------------------------------------ configs 3, 4, 5
@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("res,cell", [(10, 10.0), (25, 25.0)])
def test_config3_grid_cells_vs_oracle(bench_inputs, res, cell, precision):
    """ESM + NCC + Affine on all 1024 GridTracker cells of the bench frames, re-initialised on every frame
    (grid_reset_at_each_frame = 1, SM/src/GridTracker.cc:265-277), in both precisions (F32: the one-sweep kernel of
    lk_ncc_f32.cu; stated tolerance: median 1e-4 px, 99th percentile 2e-2 px -- the cells are tiny and some never lock)"""
    from mtf_b200 import api, workloads
    frames, _, _ = bench_inputs
    cells = workloads.grid_cells(32, cell)
    kw = dict(resx=res, resy=res, max_iters=30, epsilon=0.0)
    g = api.BatchTracker(api.make_params("ncc", "affine", "esm", n_patches=len(cells), precision=precision, **kw))
    worst = []
    for t in (0, 1):
        g.initialize(cells, frames[t])
        g.update(frames[t + 1])
        got = g.getRegion()
        assert np.isfinite(got).all()
        d = np.empty(len(cells))
        for i, c in enumerate(cells):
            o = O.OracleTracker(O.make_params("ncc", "affine", "esm", grad_mode=1, **kw))
            o.set_image(frames[t]); o.initialize(c); o.set_image(frames[t + 1]); o.update()
            d[i] = np.abs(got[i] - o.corners()).max()
        worst.append({"median": float(np.median(d)), "p99": float(np.percentile(d, 99)), "max": float(d.max())})
    _record("config3_res%d%s" % (res, "" if precision == "f64" else "_f32"), worst)
    for w in worst:
        if precision == "f64":
            assert w["median"] <= 1e-6 and w["p99"] <= 1e-3, w
        else:
            assert w["median"] <= 1e-4 and w["p99"] <= 2e-2, w


def test_config4_mi_iclk_100x100_vs_oracle():
    """ICLK + MI + Homography, 100 x 100, on the first 32 boxes of config 4's 8192-box lattice (2048 x 2048 frames)"""
    from mtf_b200 import api, synth, workloads
    frames, _ = synth.make_sequence(2, 2048, 2048, seed=1234, walk_seed=5678, sigma=1.0)
    cs = workloads.config4_patches()[::256][:32]
    kw = dict(resx=100, resy=100, max_iters=30, epsilon=0.0, hess_type=0, mi_n_bins=8, mi_pre_seed=10.0)
    g = api.BatchTracker(api.make_params("mi", "homography", "iclk", n_patches=len(cs), **kw))
    g.initialize(cs, frames[0]); g.update(frames[1])
    got = g.getRegion()
    d = np.empty(len(cs))
    for i, c in enumerate(cs):
        o = O.OracleTracker(O.make_params("mi", "homography", "iclk", grad_mode=1, **kw))
        o.set_image(frames[0]); o.initialize(c); o.set_image(frames[1]); o.update()
        d[i] = np.abs(got[i] - o.corners()).max()
    st = {"median": float(np.median(d)), "max": float(d.max())}
    _record("config4_mi_iclk_100x100", st)
    assert st["max"] <= 1e-4, st


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_config5_particles_vs_oracle(bench_inputs, precision):
    """PF + SSD + Homography: 2 objects x 10 000 particles, per-particle similarity and likelihood given the state"""
    from mtf_b200 import api, workloads
    frames, _, _ = bench_inputs
    objs = workloads.config5_objects()[[5, 40]]
    states = workloads.config5_states(2, 10000, seed=1)
    g = api.BatchTracker(api.make_params("ssd", "homography", "pf", n_patches=2, precision=precision))
    g.initialize(objs, frames[0]); g.setImage(frames[1])
    lik, sim = g.pf_evaluate(states)
    worst = 0.0
    for k in range(2):
        o = O.OracleTracker(O.make_params("ssd", "homography", "fclk"))
        o.set_image(frames[0]); o.initialize(objs[k]); o.set_image(frames[1])
        l2, s2 = o.pf_evaluate(states[k])
        worst = max(worst, float((np.abs(sim[k] - s2) / np.abs(s2)).max()))
        if precision == "f64":
            assert np.allclose(sim[k], s2, rtol=1e-12) and np.allclose(lik[k], l2, rtol=1e-12)
        else:
            assert np.allclose(sim[k], s2, rtol=2e-5) and np.allclose(lik[k], l2, rtol=2e-5, atol=1e-300)
    _record("config5_particles_%s" % precision, {"similarity_rel_max": worst})
