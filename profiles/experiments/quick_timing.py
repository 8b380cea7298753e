import sys, os
sys.path.insert(0, '.')
import numpy as np, torch
from mtf_b200 import api, synth
sys.argv = ["bench"]
import bench
frames, corners0, order = bench.workload()
dev = torch.device("cuda", 0)
d_frames = [torch.from_numpy(f).to(dev) for f in frames]
cfgs = [tuple(x.split(":")) for x in os.environ.get("CFGS", "f32:64:1024").split(",")]
for prec, T, P in cfgs:
    T = int(T); P = int(P)
    corners = synth.make_patches(P, 49.0, 1024, 1024, seed=42)
    tr = api.BatchTracker(api.make_params("ssd", "homography", "fclk", n_patches=P, max_iters=30, epsilon=0.0,
                                          threads_per_patch=T, precision=prec))  # T = 0: library default
    stream = torch.cuda.Stream(dev); tr.set_stream(stream.cuda_stream)
    tr.initialize(corners, d_frames[0]); tr.synchronize()
    for i in range(3):
        tr.setImage(d_frames[order[i]]); tr.update()
    tr.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record(stream)
    for i in range(n):
        tr.setImage(d_frames[order[(3 + i) % len(order)]]); tr.update()
    e1.record(stream); e1.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("%s %s T=%d P=%d: %.3f ms/frame  %.2f M iters/s" % (os.path.basename(os.environ.get("MTFB_LIB", "main")), prec, T, P, ms, P * 30 / ms / 1e3), flush=True)
    tr.close()
