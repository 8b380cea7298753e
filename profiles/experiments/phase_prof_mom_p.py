"""phase cycle counters of the moment kernel against the batch size (contention on an SM), experiment build -DMTFB_PROF=1"""
import ctypes as C, os, sys
sys.path.insert(0, '.')
os.environ.setdefault("MTFB_LIB", os.path.abspath("scratch/variants/libprof.so"))
import numpy as np
from mtf_b200 import api, synth
sys.argv = ["bench"]
import bench
frames, corners, order = bench.workload()
for solve in ("reference", "local"):
    for P in (148, 1024):
        tr = api.BatchTracker(api.make_params("ssd", "homography", "fclk", n_patches=P, max_iters=30, epsilon=0.0,
                                              threads_per_patch=128, precision="f32", f32_solve=solve))
        tr.initialize(corners[:P], frames[0])
        for i in range(3):
            tr.update(frames[order[i]])
        tr.synchronize()
        api.load_library().mtfb_prof_reset(tr._h)
        for i in range(3, 7):
            tr.update(frames[order[i]])
        tr.synchronize()
        out = (C.c_longlong * 16)()
        api.load_library().mtfb_prof_read(tr._h, out, 16)
        print("   slowest CTA of all frames: %.0f cycles" % out[3])
        v = np.array(out[:16], dtype=np.float64) / (P * 30 * 4)
        print("mom %s P=%d cycles per patch-pass: pixel loop %.0f, reduce %.0f, solve %.0f, apply %.0f, pass constants %.0f; per patch-frame: prologue %.0f, "
              "all passes %.0f, epilogue %.0f; solve: columns %.0f, QR %.0f, back-substitution %.0f" % (solve, P, v[8], v[9], v[10], v[11], v[12], v[13] * 30, v[14] * 30, v[15] * 30, v[4], v[5], v[6]))
