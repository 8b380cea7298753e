import sys, os
sys.path.insert(0, '.')
import numpy as np, torch
from mtf_b200 import api, synth
sys.argv = ["bench"]
import bench
frames, corners, order = bench.workload()
order = list(range(1, 8)) + list(range(6, -1, -1))   # the old order, through frame 0
dev = torch.device("cuda", 0)
d_frames = [torch.from_numpy(f).to(dev) for f in frames]
for prec in ("f32", "f64"):
    tr = api.BatchTracker(api.make_params("ssd", "homography", "fclk", n_patches=1024, max_iters=30, epsilon=0.0, precision=prec))
    stream = torch.cuda.Stream(dev); tr.set_stream(stream.cuda_stream)
    tr.initialize(corners, d_frames[0]); tr.synchronize()
    ts = []
    for i in range(48):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        tr.setImage(d_frames[order[i % len(order)]]); tr.update()
        e1.record(stream); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(prec, "frame index order:", [order[i % len(order)] for i in range(16)])
    print(" ".join("%.3f" % t for t in ts))
    st = tr.state()
    print("max |state|:", np.abs(st).max(axis=0))
