"""per-patch cycles and fp64-deferred pixel evaluations of the moment kernel (experiment build -DMTFB_PROF=2)"""
import ctypes as C, os, sys
sys.path.insert(0, '.')
os.environ.setdefault("MTFB_LIB", os.path.abspath("scratch/variants/libprof2.so"))
import numpy as np
from mtf_b200 import api, synth
sys.argv = ["bench"]
import bench
frames, corners, order = bench.workload()
for solve in ("local", "reference"):
    P = 1024
    tr = api.BatchTracker(api.make_params("ssd", "homography", "fclk", n_patches=P, max_iters=30, epsilon=0.0,
                                          threads_per_patch=128, precision="f32", f32_solve=solve))
    tr.initialize(corners[:P], frames[0])
    for i in range(6):
        tr.update(frames[order[i]])
        v = tr.getSimilarity() if hasattr(tr, "getSimilarity") else None
        if v is None:
            v = np.empty(P); api.load_library().mtfb_get_similarity(tr._h, v.ctypes.data_as(C.POINTER(C.c_double)))
        cyc = np.floor(v); slow = np.rint((v - cyc) * 1048576)
        o = np.argsort(cyc)
        print("%s frame %d: cycles min %.0f median %.0f p90 %.0f p99 %.0f max %.0f | deferred pixel evaluations per patch-frame (of 75000): median %.0f p90 %.0f p99 %.0f max %.0f | corr(cycles, deferred) %.3f"
              % (solve, i, cyc.min(), np.median(cyc), np.percentile(cyc, 90), np.percentile(cyc, 99), cyc.max(), np.median(slow),
                 np.percentile(slow, 90), np.percentile(slow, 99), slow.max(), np.corrcoef(cyc, slow)[0, 1]))
        print("   slowest 8 patches:", [(int(k), int(cyc[k]), int(slow[k])) for k in o[-8:]])
        sm = tr.n_iters() if hasattr(tr, "n_iters") else None
        if sm is None:
            sm = np.empty(P, dtype=np.int32); api.load_library().mtfb_get_n_iters(tr._h, sm.ctypes.data_as(C.POINTER(C.c_int)))
        if i == 4:
            per_sm = {}
            for k in range(P):
                per_sm.setdefault(int(sm[k]), []).append((int(cyc[k]), k))
            cnt = np.array([len(v) for v in per_sm.values()])
            print("   CTAs per SM: min %d max %d, SMs %d" % (cnt.min(), cnt.max(), len(per_sm)))
            for smid in sorted(per_sm)[:6] + sorted(per_sm)[-3:]:
                print("   SM %3d:" % smid, sorted(per_sm[smid]))
            mx = np.array([max(c for c, _ in v) for v in per_sm.values()]); mn = np.array([min(c for c, _ in v) for v in per_sm.values()])
            print("   per-SM slowest CTA: min %.0f median %.0f max %.0f; per-SM fastest CTA: min %.0f median %.0f max %.0f"
                  % (mx.min(), np.median(mx), mx.max(), mn.min(), np.median(mn), mn.max()))
