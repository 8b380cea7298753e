"""The particle filter search method (nt::PF, SM/src/NT/PF.cc).

CPU: the oracle's PF loop against independent NumPy statements of its pieces (resampling = searchsorted on the normalised
cumulative weights, the dynamic models' algebra).
GPU: the device PF tracker (mtf_b200/csrc/pf_tracker.cu behind mtfb_pf_configure / mtfb_update) against the oracle's loop
given the SAME random stream -- per frame: every particle's state, the weights, max_wt_id, the reported corners."""
import numpy as np
import pytest

from oracle import oracle_lib as O

SIGMA_PLAIN = np.array([2e-3, 2e-3, 0.4, 2e-3, 2e-3, 0.4, 2e-6, 2e-6])
SIGMA_CORNER = np.array([0.5, 0.2, 0, 0, 0, 0, 0, 0])


def _pf_params(**kw):
    from mtf_b200 import api
    return api.make_pf_params(**kw)


@pytest.fixture(scope="module")
def pf_inputs(seq384):
    from mtf_b200 import synth
    frames, warps = seq384
    corners = synth.make_patches(4, 49.0, 384, 384, seed=9)
    return frames, warps, corners


def _warp(s):
    return np.array([[1 + s[0], s[1], s[2]], [s[3], 1 + s[4], s[5]], [s[6], s[7], 1.0]])


def _state(W):
    return np.array([W[0, 0] - 1, W[0, 1], W[0, 2], W[1, 0], W[1, 1] - 1, W[1, 2], W[2, 0], W[2, 1]])


# ------------------------------------------------------------------------------------------------ CPU: the oracle's loop
@pytest.mark.parametrize("dyn,upd", [("random_walk", "additive"), ("random_walk", "compositional"),
                                     ("auto_regression1", "additive"), ("auto_regression1", "compositional")])
def test_oracle_pf_dynamic_models_and_resampling(pf_inputs, dyn, upd):
    """one frame of the oracle's PF against NumPy: perturbed states (ProjectiveBase.cc:255-299, Homography.cc:917-942) and
    multinomial resampling (NT/PF.cc:448-540 = the first particle with normalised cumulative weight >= u)"""
    frames, _, corners = pf_inputs
    n = 64
    pp = _pf_params(n_particles=n, sigma=SIGMA_PLAIN, dynamic_model=dyn, update_type=upd, corner_based_sampling=0,
                    adaptive_resampling_thresh=0.0, resampling_type="binary_multinomial", mean_type="none")
    f = O.OraclePF(O.make_params("ssd", "homography", "fclk"), pp)
    f.set_image(frames[0]); f.initialize(corners[0])
    rng = np.random.default_rng(3)
    prev = np.zeros((n, 8)); prev_ar = np.zeros((n, 8))
    for t in (1, 2):
        nrm = rng.normal(size=(1, n, 8)); uni = rng.uniform(size=(1, n))
        f.set_image(frames[t]); f.update(nrm, uni)
        st, w, cw, mx, resampled = f.particles()
        # expected perturbed states (before resampling)
        z = nrm[0] * SIGMA_PLAIN
        exp = np.empty((n, 8)); exp_ar = np.zeros((n, 8))
        for i in range(n):
            if upd == "additive":
                exp[i] = prev[i] + (prev_ar[i] if dyn == "auto_regression1" else 0) + z[i]
                exp_ar[i] = 0.5 * (exp[i] - prev[i])
            else:
                Wb, Wz = _warp(prev[i]), _warp(z[i])
                W = Wb @ (_warp(prev_ar[i]) if dyn == "auto_regression1" else np.eye(3)) @ Wz
                W = W / W[2, 2]
                exp[i] = _state(W)
                A = np.linalg.inv(Wb) @ W
                exp_ar[i] = 0.5 * _state(A / A[2, 2])
        assert resampled
        idx = np.searchsorted(cw, uni[0], side="left")            # cw is normalised after resampling
        assert np.allclose(st, exp[idx], rtol=1e-10, atol=1e-13)
        assert np.isclose(cw[-1], 1.0) and np.all(np.diff(cw) >= 0)
        assert w[idx[mx]] == w[idx].max() and mx == np.flatnonzero(w[idx] == w[idx].max()).max()
        assert np.allclose(f.state(), st[mx])
        prev = st.copy()
        prev_ar = exp_ar[idx] if dyn == "auto_regression1" else np.zeros((n, 8))


def test_oracle_pf_tracks_the_sequence(pf_inputs):
    from mtf_b200 import synth
    frames, warps, corners = pf_inputs
    pp = _pf_params(n_particles=400, sigma=SIGMA_CORNER, corner_based_sampling=1, mean_type="ssm", adaptive_resampling_thresh=0.0)
    f = O.OraclePF(O.make_params("ssd", "homography", "fclk"), pp)
    f.set_image(frames[0]); f.initialize(corners[1])
    rng = np.random.default_rng(5)
    for t in (1, 2, 3):
        f.set_image(frames[t]); f.update(rng.normal(size=(1, 400, 10)), rng.uniform(size=(1, 400)))
    truth = synth.warp_corners(warps[3], corners[1:2])[0]
    assert np.abs(f.corners() - truth).max() < 1.5          # a 400-particle filter: within a pixel or so of the ground truth


# ------------------------------------------------------------------------------------------------ GPU
CASES = [
    dict(dynamic_model="auto_regression1", update_type="compositional", corner_based_sampling=1, mean_type="none",
         adaptive_resampling_thresh=0.2),                                           # the shipped configuration (modules.cfg)
    dict(dynamic_model="random_walk", update_type="additive", corner_based_sampling=0, mean_type="ssm",
         adaptive_resampling_thresh=0.0),
    dict(dynamic_model="random_walk", update_type="compositional", corner_based_sampling=0, mean_type="corners",
         adaptive_resampling_thresh=0.0, resampling_type="linear_multinomial"),
    dict(dynamic_model="auto_regression1", update_type="additive", corner_based_sampling=1, mean_type="ssm",
         adaptive_resampling_thresh=0.0, likelihood_func="gaussian", reset_to_mean=1),
    dict(dynamic_model="auto_regression1", update_type="compositional", corner_based_sampling=0, mean_type="none",
         resampling_type="none", likelihood_func="reciprocal"),
    dict(dynamic_model="random_walk", update_type="compositional", corner_based_sampling=1, mean_type="none",
         adaptive_resampling_thresh=0.0, max_iters=3, epsilon=1e-6),
]


@pytest.mark.gpu
@pytest.mark.parametrize("am", ["ssd", "ncc"])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_pf_tracker_matches_oracle_given_the_random_stream(pf_inputs, case, am):
    from mtf_b200 import api
    frames, _, corners = pf_inputs
    kw = dict(CASES[case])
    P, n = 3, 256
    sigma = SIGMA_CORNER if kw["corner_based_sampling"] else SIGMA_PLAIN
    g = api.PFTracker(api.make_params(am, "homography", "pf", n_patches=P), n_particles=n, sigma=sigma, **kw)
    g.initialize(corners[:P], frames[0])
    orc = []
    for k in range(P):
        o = O.OraclePF(O.make_params(am, "homography", "fclk"), g.pf_params)
        o.set_image(frames[0]); o.initialize(corners[k]); orc.append(o)
    R, iters = g.n_normals, g.pf_params.max_iters
    rng = np.random.default_rng(100 + case)
    for t in (1, 2, 3):
        nrm = rng.normal(size=(iters, P, n, R)); uni = rng.uniform(size=(iters, P, n))
        g.set_random_stream(nrm, uni)
        g.update(frames[t])
        st, w, cw, mx = g.particles()
        got = g.getRegion()
        for k in range(P):
            orc[k].set_image(frames[t]); orc[k].update(nrm[:, k], uni[:, k])
            st2, w2, cw2, mx2, _ = orc[k].particles()
            assert mx[k] == mx2, (case, t, k)
            # corner based sampling: the device takes the 4-point homography in closed form (centred coordinates), the oracle
            # from the reference's DLT null vector in raw pixel coordinates -- they agree to ~1e-9 px on the perturbation
            # (the DLT side's conditioning); everything else is the same arithmetic
            tol = dict(rtol=1e-7, atol=1e-8) if kw["corner_based_sampling"] else dict(rtol=1e-9, atol=1e-12)
            assert np.allclose(st[k], st2, **tol), (case, t, k, np.abs(st[k] - st2).max())
            assert np.allclose(w[k], w2, rtol=1e-6 if kw["corner_based_sampling"] else 1e-9, atol=1e-300), (case, t, k)
            assert np.allclose(got[k], orc[k].corners(), rtol=0, atol=1e-6 if kw["corner_based_sampling"] else 1e-8), (case, t, k)


@pytest.mark.gpu
def test_pf_tracker_f32_precision_and_device_generator(pf_inputs):
    """the F32 evaluation kernel under the tracker, deviates from the device generator (Philox4x32-10 + Box-Muller), recorded
    and replayed through the oracle; streams depend on (seed, object index in the whole job), not on how objects are sharded"""
    from mtf_b200 import api
    frames, _, corners = pf_inputs
    P, n = 4, 512
    kw = dict(n_particles=n, sigma=SIGMA_CORNER, corner_based_sampling=1, mean_type="ssm", adaptive_resampling_thresh=0.0,
              record_randoms=1, seed=1234)
    g = api.PFTracker(api.make_params("ssd", "homography", "pf", n_patches=P, precision="f32"), **kw)
    g.initialize(corners[:P], frames[0])
    h = api.PFTracker(api.make_params("ssd", "homography", "pf", n_patches=2, precision="f32"), object_offset=2, **kw)
    h.initialize(corners[2:4], frames[0])
    o = O.OraclePF(O.make_params("ssd", "homography", "fclk"), g.pf_params)
    o.set_image(frames[0]); o.initialize(corners[1])
    for t in (1, 2, 3):
        g.update(frames[t]); h.update(frames[t])
        nrm, uni = g.random_stream()
        nrm_h, uni_h = h.random_stream()
        assert np.array_equal(nrm[:, 2:4], nrm_h) and np.array_equal(uni[:, 2:4], uni_h)
        assert np.allclose(g.getRegion()[2:4], h.getRegion(), atol=1e-12)
        assert abs(nrm.mean()) < 0.02 and abs(nrm.std() - 1) < 0.02 and 0 < uni.min() and uni.max() < 1
        assert abs(np.corrcoef(nrm[0, 0, :, 0], nrm[0, 0, :, 1])[0, 1]) < 0.15
        o.set_image(frames[t]); o.update(nrm[:, 1], uni[:, 1])
        st, w, cw, mx = g.particles()
        st2, w2, _, _, _ = o.particles()
        # fp32 pixel arithmetic: weights to 1e-4 relative; a uniform deviate that falls within that of a cumulative weight
        # picks the neighbouring particle (both are valid draws), so resampled states are compared as a set by their mean
        assert np.allclose(w[1], w2, rtol=2e-4)
        assert np.abs(g.getRegion()[1] - o.corners()).max() < 0.05
    g2 = api.PFTracker(api.make_params("ssd", "homography", "pf", n_patches=P, precision="f32"), **dict(kw, seed=99))
    g2.initialize(corners[:P], frames[0]); g2.update(frames[1])
    assert not np.array_equal(g2.random_stream()[0][:, :, :, 0], nrm[:, :, :, 0])


@pytest.mark.gpu
def test_pf_tracker_call_order_and_unsupported(pf_inputs):
    from mtf_b200 import api
    frames, _, corners = pf_inputs
    g = api.BatchTracker(api.make_params("ssd", "homography", "pf", n_patches=1))
    g.initialize(corners[:1], frames[0])
    with pytest.raises(api.MTFError) as e:
        g.update(frames[1])
    assert e.value.status == 3
    with pytest.raises(api.MTFError) as e:
        api.PFTracker(api.make_params("ssd", "affine", "pf", n_patches=1), sigma=0.1)
    assert e.value.status == 2
    with pytest.raises(api.MTFError) as e:
        api.PFTracker(api.make_params("ssd", "homography", "pf", n_patches=1), sigma=0.1, resampling_type="residual")
    assert e.value.status == 2
    # setRegion re-seeds the particles at the new region (NT/PF.cc:596-600)
    t = api.PFTracker(api.make_params("ssd", "homography", "pf", n_patches=2), n_particles=64, sigma=SIGMA_CORNER)
    t.initialize(corners[:2], frames[0]); t.update(frames[1])
    t.setRegion(corners[2:4])
    st, w, _, _ = t.particles()
    assert np.all(st == 0) and np.allclose(w, 1 / 64) and np.allclose(t.getRegion(), corners[2:4])


@pytest.mark.gpu
def test_pf_tracker_at_config5_size(pf_inputs):
    """BASELINE config 5's per-object size -- 10 000 particles (weights in 80 KB of shared memory, the sequential prefix sum,
    10 000 binary searches) -- against the oracle's loop given the recorded device-generated stream, plus the size-independent
    properties of a resampling step"""
    from mtf_b200 import api
    frames, _, corners = pf_inputs
    P, n = 2, 10000
    g = api.PFTracker(api.make_params("ssd", "homography", "pf", n_patches=P), n_particles=n, sigma=SIGMA_CORNER,
                      corner_based_sampling=1, mean_type="ssm", adaptive_resampling_thresh=0.0, record_randoms=1, seed=5)
    g.initialize(corners[:P], frames[0])
    o = O.OraclePF(O.make_params("ssd", "homography", "fclk"), g.pf_params)
    o.set_image(frames[0]); o.initialize(corners[0])
    for t in (1, 2):
        g.update(frames[t])
        nrm, uni = g.random_stream()
        st, w, cw, mx = g.particles()
        # properties: normalised cumulative weights are a distribution function; every resampled state is one of the weights'
        # owners; the reported region is the mean state's
        assert np.all(np.diff(cw, axis=1) >= 0) and np.allclose(cw[:, -1], 1.0, rtol=0, atol=1e-15)
        assert np.all(w > 0) and np.all(np.isfinite(st))
        o.set_image(frames[t]); o.update(nrm[:, 0], uni[:, 0])
        st2, w2, cw2, mx2, resampled = o.particles()
        assert resampled and mx[0] == mx2
        assert np.allclose(w[0], w2, rtol=1e-6, atol=1e-300)
        assert np.allclose(cw[0], cw2, rtol=1e-6, atol=1e-12)
        assert np.allclose(st[0], st2, rtol=1e-7, atol=1e-8)
        assert np.allclose(g.getRegion()[0], o.corners(), rtol=0, atol=1e-6)
        assert np.allclose(g.state()[0], st[0].mean(axis=0), rtol=1e-10, atol=1e-13)
