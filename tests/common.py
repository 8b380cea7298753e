"""Shared helpers of the test suite."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def patches(n, side, h, w, seed=42):
    from mtf_b200 import synth
    return synth.make_patches(n, side, h, w, seed=seed, margin=40.0)


def quad_patches(n, h, w, seed=7, side=50.0, jitter=4.0):
    """general quadrilaterals (not axis-aligned rectangles): the DLT then has a non-trivial third row,
    which exercises the un-normalised init_pts_hm quirk of Homography.cc:68"""
    from mtf_b200 import synth
    c = synth.make_patches(n, side + 0.37, h, w, seed=seed, margin=40.0 + jitter)
    rng = np.random.default_rng(seed)
    return c + rng.uniform(-jitter, jitter, size=c.shape)


# ------------------------------------------------------------------ host build of lk_math.cuh (CPU tests)
_hm = None


def host_math():
    global _hm
    if _hm is not None:
        return _hm
    src = os.path.join(HERE, "host_math", "host_math.cpp")
    hdr = os.path.join(ROOT, "mtf_b200", "csrc", "lk_math.cuh")
    out = os.path.join(HERE, "host_math", "_build", "libhost_math.so")
    if not os.path.exists(out) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(out):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", "-std=c++17",
                               "-o", out, src])
    L = C.CDLL(out)
    dp, fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
    L.hm_sample_grad.argtypes = [fp, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_double, dp, dp, dp]
    L.hm_sample.argtypes = [fp, C.c_int, C.c_int, C.c_int, dp, C.c_int, dp]
    L.hm_stage.argtypes = [C.c_int, fp, C.c_int, C.c_int, C.c_int, dp, C.c_int, dp, C.c_int, dp, dp, C.c_double,
                           dp, dp, dp, dp, dp]
    L.hm_compose.argtypes = [C.c_int, dp, dp, dp]
    L.hm_invert_state.argtypes = [C.c_int, dp, dp]
    L.hm_warp_corners.argtypes = [C.c_int, dp, dp, dp]
    _hm = L
    return L


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def warp_from_state(ssm, s):
    if ssm == "homography":
        return np.array([[1 + s[0], s[1], s[2]], [s[3], 1 + s[4], s[5]], [s[6], s[7], 1.0]])
    return np.array([[1 + s[2], s[3], s[0]], [s[4], 1 + s[5], s[1]], [0, 0, 1.0]])
