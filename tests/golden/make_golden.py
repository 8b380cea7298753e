#!/usr/bin/env python
"""Mints the golden vectors in this directory from the CPU oracle (oracle/mtf_oracle.cpp).

The reference ships no tests and no golden vectors for this path and cannot be built in this image (SURVEY.md
section 4 and 8c), so these fixtures pin the ORACLE (regression) and give the CUDA path a fixed target that does not
depend on the oracle being rebuilt: small seeded cases, inputs and outputs stored together.

    python tests/golden/make_golden.py        # rewrites tests/golden/lk_*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from mtf_b200 import synth  # noqa: E402
from oracle import oracle_lib as O  # noqa: E402

CASES = [
    # name, am, ssm, sm, res, extra oracle params
    ("ssd_hom_fclk", "ssd", "homography", "fclk", 24, {}),
    ("ssd_hom_esm", "ssd", "homography", "esm", 24, {}),
    ("ssd_aff_iclk", "ssd", "affine", "iclk", 24, {}),
    ("ssd_hom_fclk_lm", "ssd", "homography", "fclk", 24, {"leven_marq": 1}),
    ("ncc_aff_esm", "ncc", "affine", "esm", 16, {}),
    ("ncc_hom_fclk", "ncc", "homography", "fclk", 24, {}),
    ("mi_hom_iclk", "mi", "homography", "iclk", 32, {"hess_type": 0}),
]


def main():
    frames, _ = synth.make_sequence(3, 128, 128, seed=77, walk_seed=78, sigma=0.6)
    frames = [np.ascontiguousarray(f) for f in frames]
    rng = np.random.default_rng(79)
    for name, am, ssm, sm, res, extra in CASES:
        side = res - 1.0 if name.endswith("esm") else res + 0.37
        cs = synth.make_patches(4, side, 128, 128, seed=80, margin=12.0)
        cs[2:] += rng.uniform(-1.5, 1.5, size=cs[2:].shape)          # two general quadrilaterals
        out = {"frames": np.stack(frames), "corners": cs, "res": res}
        for gm in (0, 1):
            per_patch = []
            for c in cs:
                o = O.OracleTracker(O.make_params(am, ssm, sm, resx=res, resy=res, grad_mode=gm, max_iters=12, **extra))
                o.set_image(frames[0]); o.initialize(c)
                rec = {"init_warp": o.init_warp(), "init_pix_vals": o.init_pix_vals(), "init_pts": o.init_pts()}
                logs = []
                for fr in frames[1:]:
                    o.set_image(fr); o.update()
                    logs.append(o.log())
                rec["n_iters"] = np.array([len(l) for l in logs])
                rec["f"] = np.array([e["f"] for l in logs for e in l])
                rec["jacobian"] = np.array([e["jacobian"] for l in logs for e in l])
                rec["hessian"] = np.array([e["hessian"] for l in logs for e in l])
                rec["iter_corners"] = np.array([e["corners"] for l in logs for e in l])
                rec["rejected"] = np.array([e["rejected"] for l in logs for e in l])
                rec["final_corners"] = o.corners()
                per_patch.append(rec)
            for k in per_patch[0]:
                out["gm%d_%s" % (gm, k)] = np.array([r[k] for r in per_patch], dtype=object if k in
                                                    ("f", "jacobian", "hessian", "iter_corners", "rejected") else None)
        np.savez_compressed(os.path.join(HERE, "lk_%s.npz" % name), **out)
        print("wrote lk_%s.npz" % name)
    # particle evaluation
    cs = synth.make_patches(2, 24.37, 128, 128, seed=81, margin=12.0)
    states = rng.normal(size=(2, 40, 8)) * np.array([1e-2, 1e-2, 1.0, 1e-2, 1e-2, 1.0, 1e-5, 1e-5])
    lik, sim = [], []
    for i, c in enumerate(cs):
        o = O.OracleTracker(O.make_params("ssd", "homography", "fclk", resx=24, resy=24, likelihood_alpha=20.0))
        o.set_image(frames[0]); o.initialize(c); o.set_image(frames[1])
        a, b = o.pf_evaluate(states[i]); lik.append(a); sim.append(b)
    np.savez_compressed(os.path.join(HERE, "pf_ssd_hom.npz"), frames=np.stack(frames), corners=cs, states=states,
                        likelihood=np.array(lik), similarity=np.array(sim), res=24, alpha=20.0)
    print("wrote pf_ssd_hom.npz")


if __name__ == "__main__":
    main()
