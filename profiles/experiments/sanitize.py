import sys
sys.path.insert(0, '.')
import numpy as np
from mtf_b200 import api, synth
frames, _ = synth.make_sequence(3, 256, 256, seed=5, walk_seed=6, sigma=2.5)
for T in (64, 32, 128):
    for ssm, sm in (("homography", "fclk"), ("affine", "esm"), ("homography", "iclk")):
        cs = synth.make_patches(6, 49.0, 256, 256, seed=3, margin=20.0)
        cs[0] += np.array([[-60.0], [0.0]])          # hangs over the border
        tr = api.BatchTracker(api.make_params("ssd", ssm, sm, n_patches=len(cs), precision="f32", threads_per_patch=T, max_iters=6))
        tr.initialize(cs, frames[0])
        for f in frames[1:]:
            tr.update(f)
        print(T, ssm, sm, np.isfinite(tr.getRegion()).all(), flush=True)
        tr.close()
tr = api.BatchTracker(api.make_params("ssd", "homography", "fclk", n_patches=2))
tr.setRawImage(np.clip(frames[0], 0, 255).astype(np.uint8)); print("preproc", tr.image(256, 256).mean())
