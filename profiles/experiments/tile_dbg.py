import sys; sys.path.insert(0,'.')
import numpy as np
from mtf_b200 import api, synth
frames,_ = synth.make_sequence(2,384,384)
cs = synth.make_patches(6, 52.3, 384, 384, seed=17)
g = api.BatchTracker(api.make_params("ssd","homography","fclk", n_patches=len(cs), threads_per_patch=32, occupancy=0))
g.initialize(cs, frames[0]); g.update(frames[1]); print(g.n_iters())
