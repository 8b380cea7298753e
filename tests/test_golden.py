"""Committed golden vectors (tests/golden/*.npz, minted by tests/golden/make_golden.py from the oracle).
CPU: the oracle still reproduces them bit for bit.  GPU: the CUDA path, through the C-ABI, against them."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle_lib as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {"ssd_hom_fclk": ("ssd", "homography", "fclk", {}), "ssd_hom_esm": ("ssd", "homography", "esm", {}),
         "ssd_aff_iclk": ("ssd", "affine", "iclk", {}), "ssd_hom_fclk_lm": ("ssd", "homography", "fclk", {"leven_marq": 1}),
         "ncc_aff_esm": ("ncc", "affine", "esm", {}), "ncc_hom_fclk": ("ncc", "homography", "fclk", {}),
         "mi_hom_iclk": ("mi", "homography", "iclk", {"hess_type": 0})}


def _load(name):
    return np.load(os.path.join(GOLDEN, "lk_%s.npz" % name), allow_pickle=True)


def test_all_fixtures_present():
    have = {os.path.basename(p)[3:-4] for p in glob.glob(os.path.join(GOLDEN, "lk_*.npz"))}
    assert have == set(CASES)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    am, ssm, sm, extra = CASES[name]
    d = _load(name)
    res = int(d["res"])
    frames = [np.ascontiguousarray(f) for f in d["frames"]]
    for gm in (0, 1):
        for i, c in enumerate(d["corners"]):
            o = O.OracleTracker(O.make_params(am, ssm, sm, resx=res, resy=res, grad_mode=gm, max_iters=12, **extra))
            o.set_image(frames[0]); o.initialize(c)
            assert np.array_equal(o.init_warp(), d["gm%d_init_warp" % gm][i])
            assert np.array_equal(o.init_pix_vals(), d["gm%d_init_pix_vals" % gm][i])
            logs = []
            for fr in frames[1:]:
                o.set_image(fr); o.update(); logs += o.log()
            assert np.array_equal(np.array([e["f"] for e in logs]), np.asarray(d["gm%d_f" % gm][i], dtype=float))
            assert np.array_equal(np.array([e["corners"] for e in logs]), np.asarray(d["gm%d_iter_corners" % gm][i], dtype=float))
            assert np.array_equal(o.corners(), d["gm%d_final_corners" % gm][i])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_golden(name):
    from mtf_b200 import api
    am, ssm, sm, extra = CASES[name]
    d = _load(name)
    res = int(d["res"])
    frames = [np.ascontiguousarray(f) for f in d["frames"]]
    cs = d["corners"]
    tr = api.BatchTracker(api.make_params(am, ssm, sm, n_patches=len(cs), resx=res, resy=res, max_iters=12, **extra))
    tr.enable_iter_log(12)
    tr.initialize(cs, frames[0])
    assert np.array_equal(tr.init_warp(), np.stack(d["gm1_init_warp"]))
    assert np.array_equal(tr.init_pix_vals(), np.stack(d["gm1_init_pix_vals"]))
    assert np.array_equal(tr.init_pts(), np.stack(d["gm1_init_pts"]))
    logs = [[] for _ in cs]
    n_it = []
    for fr in frames[1:]:
        tr.update(fr)
        for i, l in enumerate(tr.iter_log()):
            logs[i] += l
        n_it.append(tr.n_iters().copy())
    n_it = np.array(n_it).T
    loose = am == "mi"
    for i in range(len(cs)):
        assert np.array_equal(n_it[i], d["gm1_n_iters"][i])
        gc = np.asarray(d["gm1_iter_corners"][i], dtype=float)
        gf = np.asarray(d["gm1_f"][i], dtype=float)
        assert len(logs[i]) == len(gf)
        assert np.abs(np.array([e["corners"] for e in logs[i]]) - gc).max() <= (1e-4 if loose else 1e-5)
        assert np.allclose([e["f"] for e in logs[i]], gf, rtol=1e-6, atol=1e-9)
        assert [e["rejected"] for e in logs[i]] == [bool(x) for x in d["gm1_rejected"][i]]
        # against the reference's finite-difference gradients: same fixed point within the quotient's noise
        assert np.abs(tr.getRegion()[i] - d["gm0_final_corners"][i]).max() <= (5e-2 if loose else 5e-3)


@pytest.mark.gpu
def test_cuda_pf_matches_golden():
    from mtf_b200 import api
    d = np.load(os.path.join(GOLDEN, "pf_ssd_hom.npz"))
    frames = [np.ascontiguousarray(f) for f in d["frames"]]
    tr = api.BatchTracker(api.make_params("ssd", "homography", "pf", n_patches=len(d["corners"]), resx=int(d["res"]),
                                          resy=int(d["res"]), likelihood_alpha=float(d["alpha"])))
    tr.initialize(d["corners"], frames[0])
    tr.setImage(frames[1])
    lik, sim = tr.pf_evaluate(d["states"])
    assert np.allclose(sim, d["similarity"], rtol=1e-12, atol=0)
    assert np.allclose(lik, d["likelihood"], rtol=1e-11, atol=0)
