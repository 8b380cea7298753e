#!/usr/bin/env python
"""Device-resident timings of the BASELINE configs that are parity-test cases rather than bench lines (3, 4, 5),
through the same C ABI, CUDA events on the library's stream.  Informational: python profiles/other_configs.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mtf_b200 import api, synth  # noqa: E402


def timed(fn, stream, reps):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
    frames, _ = synth.make_sequence(3, 1024, 1024)
    d_frames = [torch.from_numpy(f).to(dev) for f in frames]
    out = []

    def lk(name, am, ssm, sm, P, res, side, iters, **kw):
        cs = synth.make_patches(P, side, 1024, 1024, margin=10.0 + side / 2)
        tr = api.BatchTracker(api.make_params(am, ssm, sm, n_patches=P, resx=res, resy=res, max_iters=iters, epsilon=0.0, **kw))
        tr.set_stream(stream.cuda_stream)
        tr.initialize(cs, d_frames[0])
        state = {"i": 0}

        def step():
            state["i"] += 1
            tr.setImage(d_frames[1 + state["i"] % 2]); tr.update()
        ms = timed(step, stream, 10)
        out.append({"config": name, "ms_per_frame": ms, "patch_passes_per_s": P * iters / (ms * 1e-3), "patches": P,
                    "pixels_per_patch": res * res, "passes_per_frame": iters, "finite": bool(np.isfinite(tr.getRegion()).all())})

    lk("2: FCLK+SSD+Homography 1024 x 50x50", "ssd", "homography", "fclk", 1024, 50, 49.0, 30)
    lk("2': same with hom_normalized_init=1 (Config/modules.cfg)", "ssd", "homography", "fclk", 1024, 50, 49.0, 30, hom_normalized_init=1)
    lk("1': ESM+SSD+Homography 1024 x 50x50", "ssd", "homography", "esm", 1024, 50, 49.0, 30)
    # the fp32-arithmetic precision (SSD): same configurations
    lk("2 (F32): FCLK+SSD+Homography 1024 x 50x50", "ssd", "homography", "fclk", 1024, 50, 49.0, 30, precision="f32")
    lk("2' (F32): same with hom_normalized_init=1", "ssd", "homography", "fclk", 1024, 50, 49.0, 30, hom_normalized_init=1, precision="f32")
    lk("1' (F32): ESM+SSD+Homography 1024 x 50x50 (SumOfSelf: reference-basis QR)", "ssd", "homography", "esm", 1024, 50, 49.0, 30, precision="f32")
    lk("(F32): ESM+SSD+Homography, CurrentSelf Hessian (local-basis solve)", "ssd", "homography", "esm", 1024, 50, 49.0, 30, precision="f32", hess_type=1)
    lk("(F32): ICLK+SSD+Homography 1024 x 50x50", "ssd", "homography", "iclk", 1024, 50, 49.0, 30, precision="f32")
    lk("(F32): FCLK+SSD+Affine 1024 x 50x50", "ssd", "affine", "fclk", 1024, 50, 49.0, 30, precision="f32")
    lk("(F64): FCLK+SSD+Affine 1024 x 50x50", "ssd", "affine", "fclk", 1024, 50, 49.0, 30)
    lk("(F64): ICLK+SSD+Homography 1024 x 50x50", "ssd", "homography", "iclk", 1024, 50, 49.0, 30)
    lk("(F32): FCLK+SSD+Homography 1024 cells 25x25", "ssd", "homography", "fclk", 1024, 25, 25.0, 30, precision="f32")
    lk("3: ESM+NCC+Affine 1024 cells 10x10", "ncc", "affine", "esm", 1024, 10, 10.0, 30)
    lk("3: ESM+NCC+Affine 1024 cells 25x25", "ncc", "affine", "esm", 1024, 25, 25.0, 30)
    lk("4: ICLK+MI+Homography 1024 x 100x100 (one GPU's share of 8192 on 8)", "mi", "homography", "iclk", 1024, 100, 99.0, 30,
       hess_type=0)
    # 5: PF, 64 objects x 10000 particles (one GPU's eighth = 8 objects)
    P, n = 8, 10000
    cs = synth.make_patches(P, 49.0, 1024, 1024)
    tr = api.BatchTracker(api.make_params("ssd", "homography", "pf", n_patches=P))
    tr32 = api.BatchTracker(api.make_params("ssd", "homography", "pf", n_patches=P, precision="f32"))      # F32 particle evaluation
    tr32.set_stream(stream.cuda_stream)
    tr.set_stream(stream.cuda_stream)
    tr.initialize(cs, d_frames[0]); tr.setImage(d_frames[1])
    rng = np.random.default_rng(0)
    states = torch.from_numpy(rng.normal(size=(P, n, 8)) * np.array([1e-2, 1e-2, 1.0, 1e-2, 1e-2, 1.0, 1e-5, 1e-5])).to(dev)
    lik = torch.empty((P, n), dtype=torch.float64, device=dev); sim = torch.empty_like(lik)
    L = api.load_library()

    def pf():
        tr._check(L.mtfb_pf_evaluate_device(tr._h, states.data_ptr(), n, lik.data_ptr(), sim.data_ptr()))
    ms = timed(pf, stream, 5)
    out.append({"config": "5: PF+SSD+Homography 8 objects x 10000 particles x 50x50 (one GPU's share of 64 objects on 8)",
                "ms_per_frame": ms, "particles_per_s": P * n / (ms * 1e-3), "finite": bool(torch.isfinite(lik).all())})
    def pf32():
        tr32._check(L.mtfb_pf_evaluate_device(tr32._h, states.data_ptr(), n, lik.data_ptr(), sim.data_ptr()))
    tr32.initialize(cs, d_frames[0]); tr32.setImage(d_frames[1])
    ms = timed(pf32, stream, 5)
    out.append({"config": "5 (F32): PF+SSD+Homography 8 objects x 10000 particles x 50x50",
                "ms_per_frame": ms, "particles_per_s": P * n / (ms * 1e-3), "finite": bool(torch.isfinite(lik).all())})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
