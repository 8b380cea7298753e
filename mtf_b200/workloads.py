"""The BASELINE.json configurations as concrete, seeded workloads, shared by bench.py, the parity tests
on the benchmark's own inputs (tests/test_gpu_bench_workload.py) and profiles/.  Host side, NumPy only.

  config 2  FCLK + SSD + Homography, 1024 patches 50 x 50 on 1024 x 1024 frames (the headline)
  config 3  ESM + NCC + Affine, GridTracker 32 x 32 cells (10 x 10 default / 25 x 25 modules.cfg), re-initialised every frame
  config 4  ICLK + MI + Homography, 100 x 100 patches, 8192 of them over the GPUs (2048 x 2048 frames)
  config 5  PF + SSD + Homography, 64 objects x 10 000 particles
"""
import os

import numpy as np

from . import synth

IMG = 1024
N_FRAMES = 8


def sequence(n_frames=N_FRAMES, size=IMG, cache=True):
    """frames 0..n_frames-1 of the synthetic sequence (seed 1234, walk seed 5678, sigma 1 px) and the ground-truth warps;
    cached in /tmp (synthesis takes a few seconds per process)"""
    path = "/tmp/mtfb_bench_seq_%d_%d.npz" % (n_frames, size)
    if cache:
        try:
            z = np.load(path)
            return list(z["frames"]), list(z["warps"])
        except Exception:
            pass
    frames, warps = synth.make_sequence(n_frames, size, size, seed=1234, walk_seed=5678, sigma=1.0)
    if cache:
        try:
            tmp = path + ".%d.npz" % os.getpid()
            np.savez(tmp, frames=np.stack(frames), warps=np.stack(warps))
            os.replace(tmp, path)
        except Exception:
            pass
    return frames, warps


def frame_order(n_frames=N_FRAMES):
    """ping-pong over frames 1 .. n_frames-1: consecutive frames stay one random-walk step apart for any number of steps;
    frame 0 only initialises ("frames 1..T")"""
    return list(range(1, n_frames)) + list(range(n_frames - 2, 1, -1))


# ---------------------------------------------------------------------------------------------- config 2
CONFIG2 = dict(am="ssd", ssm="homography", sm="fclk", res=50, side=49.0, n_patches=1024, max_iters=30,
               hess_type=1, hom_normalized_init=0)


def config2_patches(n_patches=1024, seed_offset=0, size=IMG):
    return synth.make_patches(n_patches, 49.0, size, size, seed=42 + seed_offset)


# ---------------------------------------------------------------------------------------------- config 3
def grid_cells(grid=32, cell=None, size=IMG, margin=24.0):
    """GridTracker's cell layout (SM/src/GridTracker.cc:395-441 initTrackers: grid_res x grid_res cells tiling the tracked
    region): here the region is the frame minus a margin, `cell` the cell side in px (None: the cells tile the region)"""
    lo, hi = margin, size - 1 - margin
    step = (hi - lo) / grid
    side = step if cell is None else cell
    out = np.empty((grid * grid, 2, 4))
    for r in range(grid):
        for c in range(grid):
            cx, cy = lo + (c + 0.5) * step, lo + (r + 0.5) * step
            x0, y0 = cx - side / 2, cy - side / 2
            out[r * grid + c, 0] = [x0, x0 + side, x0 + side, x0]
            out[r * grid + c, 1] = [y0, y0, y0 + side, y0 + side]
    return out


CONFIG3 = dict(am="ncc", ssm="affine", sm="esm", grid=32, max_iters=30, hess_type=2, jac_type=1)


# ---------------------------------------------------------------------------------------------- config 4
CONFIG4 = dict(am="mi", ssm="homography", sm="iclk", res=100, side=99.0, n_patches=8192, max_iters=30, hess_type=0,
               mi_n_bins=8, mi_pre_seed=10.0, mi_pou=0, size=2048)


def config4_patches(n_patches=8192, size=2048):
    """8192 boxes of 99 px on a 2048 x 2048 frame: a 91 x 91 lattice, neighbours overlap (as GridTracker cells and
    particle clouds do)"""
    return synth.make_patches(n_patches, 99.0, size, size, seed=43, margin=20.0)


# ---------------------------------------------------------------------------------------------- config 5
CONFIG5 = dict(am="ssd", ssm="homography", sm="pf", res=50, side=49.0, n_objects=64, n_particles=10000)

PF_SIGMA_HOM = np.array([1e-2, 1e-2, 1.0, 1e-2, 1e-2, 1.0, 1e-5, 1e-5])


def config5_objects(n_objects=64, size=IMG):
    return synth.make_patches(n_objects, 49.0, size, size, seed=44)


def config5_states(n_objects, n_particles, seed=0):
    """particle states from a seeded host generator (the reference seeds Boost.Random from random_device,
    ProjectiveBase.cc:192-197: trajectories are not reproducible by design; parity is per particle given the state)"""
    rng = np.random.default_rng(seed)
    return rng.normal(size=(n_objects, n_particles, 8)) * PF_SIGMA_HOM
