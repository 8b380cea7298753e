"""phase cycle counters of the moment kernel (experiment build -DMTFB_PROF=1, profiles/build_mom_variant.sh prof "-DMTFB_PROF=1")"""
import ctypes as C, os, sys
sys.path.insert(0, '.')
os.environ.setdefault("MTFB_LIB", os.path.abspath("scratch/variants/libprof.so"))
import numpy as np
from mtf_b200 import api, synth
sys.argv = ["bench"]
import bench
frames, corners, order = bench.workload()
for solve in ("reference", "local"):
    for T in (128,):
        tr = api.BatchTracker(api.make_params("ssd", "homography", "fclk", n_patches=1024, max_iters=30, epsilon=0.0,
                                              threads_per_patch=T, precision="f32", f32_solve=solve))
        tr.initialize(corners, frames[0])
        for i in range(4):
            tr.update(frames[order[i]])
        tr.synchronize()
        out = (C.c_longlong * 16)()
        api.load_library().mtfb_prof_read(tr._h, out, 16)
        v = np.array(out[:16], dtype=np.float64) / (1024 * 30 * 4)
        print(os.path.basename(os.environ["MTFB_LIB"]), "mom %s T=%d cycles per patch-pass: pixel loop %.0f, reduce %.0f, basis map %.0f, serial step / "
              "apply %.0f (setup %.0f, QR %.0f, back-subst %.0f), pass constants %.0f" % (solve, T, v[8], v[9], v[10], v[11], v[4], v[5], v[6], v[12]))
