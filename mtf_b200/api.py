"""ctypes binding of libmtf_b200.so (include/mtf_b200.h) and a host-side mirror of the reference's
tracker interface (include/mtf/TrackerBase.h:9-70): setImage / initialize / update / setRegion / getRegion.

The C++ shim a reference maintainer would compile into MTF is include/mtf_b200_tracker.h; this module is
the same thing for Python callers (tests, bench.py).  There is no CPU path: if the shared library is missing
or no B200 is visible, construction raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MTFB_LIB") or os.path.join(_HERE, "libmtf_b200.so")   # MTFB_LIB: experiment builds

AM = {"ssd": 0, "ncc": 1, "mi": 2}
SSM = {"homography": 0, "affine": 1, "translation": 2, "8": 0, "6": 1, "2": 2}
SM = {"esm": 0, "fclk": 1, "iclk": 2, "pf": 3, "falk": 4, "ialk": 5}
ESM_HESS = {"initial_self": 0, "current_self": 1, "sum_of_self": 2, "original": 3, "sum_of_std": 4, "std": 5}
ESM_JAC = {"original": 0, "diff_of_jacs": 1}
LK_HESS = {"initial_self": 0, "current_self": 1, "std": 2}
PRECISION = {"f64": 0, "f32": 1}
F32_SOLVE = {"reference": 0, "local": 1}

STATUS_NAMES = {1: "InvalidArgument", 2: "FunctonNotImplemented", 3: "LogicError", 4: "InvalidTrackerState",
                5: "CudaError", 6: "OutOfMemory"}


class MTFError(RuntimeError):
    """Counterpart of mtf::utils::Exception (Utilities/include/mtf/Utilities/excpUtils.h:8-55)."""

    def __init__(self, status, message):
        super().__init__("%s: %s" % (STATUS_NAMES.get(status, status), message))
        self.status = status
        self.type = STATUS_NAMES.get(status, str(status))


class Params(C.Structure):
    _fields_ = [("am", C.c_int), ("ssm", C.c_int), ("sm", C.c_int),
                ("resx", C.c_int), ("resy", C.c_int), ("n_patches", C.c_int),
                ("max_iters", C.c_int), ("epsilon", C.c_double),
                ("hess_type", C.c_int), ("jac_type", C.c_int),
                ("chained_warp", C.c_int), ("leven_marq", C.c_int),
                ("lm_delta_init", C.c_double), ("lm_delta_update", C.c_double),
                ("nt_semantics", C.c_int), ("grad_eps", C.c_double),
                ("hom_normalized_init", C.c_int), ("mi_n_bins", C.c_int),
                ("mi_pre_seed", C.c_double), ("mi_pou", C.c_int),
                ("likelihood_alpha", C.c_double), ("device", C.c_int), ("threads_per_patch", C.c_int),
                ("occupancy", C.c_int), ("precision", C.c_int), ("f32_solve", C.c_int)]


PF_DYNAMIC = {"random_walk": 0, "auto_regression1": 1}
PF_UPDATE = {"additive": 0, "compositional": 1}
PF_LIKELIHOOD = {"am": 0, "gaussian": 1, "reciprocal": 2}
PF_RESAMPLING = {"none": 0, "binary_multinomial": 1, "linear_multinomial": 2, "residual": 3}
PF_MEAN = {"none": 0, "ssm": 1, "corners": 2}


class PFParams(C.Structure):
    """mtfb_pf_params = the fields of PFParams (SM/include/mtf/SM/PFParams.h) this library implements"""
    _fields_ = [("n_particles", C.c_int), ("max_iters", C.c_int), ("epsilon", C.c_double),
                ("dynamic_model", C.c_int), ("update_type", C.c_int), ("likelihood_func", C.c_int),
                ("resampling_type", C.c_int), ("mean_type", C.c_int), ("reset_to_mean", C.c_int),
                ("adaptive_resampling_thresh", C.c_double), ("measurement_sigma", C.c_double), ("ar_coeff", C.c_double),
                ("ssm_sigma", C.c_double * 8), ("ssm_mean", C.c_double * 8), ("corner_based_sampling", C.c_int),
                ("seed", C.c_ulonglong), ("object_offset", C.c_int), ("record_randoms", C.c_int)]


class EstParams(C.Structure):
    """mtfb_est_params = SSMEstimatorParams (SSM/src/SSMEstimatorParams.cc) + the cvRNG seed"""
    _fields_ = [("method", C.c_int), ("ransac_reproj_thresh", C.c_double), ("n_model_pts", C.c_int), ("refine", C.c_int),
                ("max_iters", C.c_int), ("max_subset_attempts", C.c_int), ("confidence", C.c_double),
                ("lm_max_iters", C.c_int), ("seed", C.c_ulonglong)]


EST_METHOD = {"ransac": 0, "lmeds": 1, "least_squares": 2}


class IterLog(C.Structure):
    _fields_ = [("f", C.c_double), ("jacobian", C.c_double * 8), ("hessian", C.c_double * 64),
                ("state_update", C.c_double * 8), ("corners", C.c_double * 8),
                ("update_norm", C.c_double), ("rejected", C.c_int), ("valid", C.c_int)]


EXPORTS = [
    "mtfb_last_error", "mtfb_version", "mtfb_launch_count", "mtfb_default_params", "mtfb_create", "mtfb_destroy",
    "mtfb_set_stream", "mtfb_synchronize", "mtfb_set_image", "mtfb_set_image_device", "mtfb_set_image_u8",
    "mtfb_set_image_async", "mtfb_set_image_u8_async",
    "mtfb_get_image", "mtfb_initialize",
    "mtfb_set_region", "mtfb_update", "mtfb_iterate_once", "mtfb_enable_iter_log", "mtfb_get_iter_log",
    "mtfb_pf_evaluate", "mtfb_pf_evaluate_device", "mtfb_get_corners", "mtfb_get_state", "mtfb_get_n_iters",
    "mtfb_get_similarity", "mtfb_get_patch_status", "mtfb_get_init_warp", "mtfb_get_init_pts",
    "mtfb_get_init_pix_vals", "mtfb_get_curr_stage", "mtfb_get_curr_stage_f32", "mtfb_device_results",
    "mtfb_state_size", "mtfb_debug_colpiv_qr_solve",
    "mtfb_pf_default_params", "mtfb_pf_configure", "mtfb_pf_set_random_stream", "mtfb_pf_get_random_stream", "mtfb_pf_get_particles",
    "mtfb_est_default_params", "mtfb_estimate_warp_from_pts", "mtfb_estimate_warp_from_corners_device", "mtfb_grid_enable", "mtfb_grid_estimate", "mtfb_grid_commit", "mtfb_grid_advance", "mtfb_grid_initialize",
    "mtfb_grid_get_pts",
    "mtfb_peer_export", "mtfb_peer_attach", "mtfb_peer_gather", "mtfb_peer_gathered", "mtfb_get_gathered_region",
]

PEER_HANDLE_BYTES = 64
_lib = None


def load_library(path=LIB_PATH):
    """dlopen the C-ABI library and declare its prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise MTFError(5, "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C mtf_b200/csrc); there is no CPU fallback" % path)
    L = C.CDLL(path)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    L.mtfb_last_error.restype = C.c_char_p
    L.mtfb_version.restype = C.c_char_p
    L.mtfb_launch_count.argtypes = [vp]; L.mtfb_launch_count.restype = C.c_long
    L.mtfb_default_params.argtypes = [C.POINTER(Params)]; L.mtfb_default_params.restype = None
    L.mtfb_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.mtfb_destroy.argtypes = [vp]
    L.mtfb_set_stream.argtypes = [vp, vp]
    L.mtfb_synchronize.argtypes = [vp]
    L.mtfb_set_image.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
    L.mtfb_set_image_device.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
    L.mtfb_set_image_u8.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
    L.mtfb_set_image_async.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
    L.mtfb_set_image_u8_async.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
    L.mtfb_get_image.argtypes = [vp, vp]
    L.mtfb_initialize.argtypes = [vp, dp]
    L.mtfb_set_region.argtypes = [vp, dp]
    L.mtfb_update.argtypes = [vp]
    L.mtfb_iterate_once.argtypes = [vp, dp, dp, dp, dp]
    L.mtfb_enable_iter_log.argtypes = [vp, C.c_int]
    L.mtfb_get_iter_log.argtypes = [vp, C.POINTER(IterLog)]
    L.mtfb_pf_evaluate.argtypes = [vp, dp, C.c_int, dp, dp]
    L.mtfb_pf_evaluate_device.argtypes = [vp, vp, C.c_int, vp, vp]
    for name in ("mtfb_get_corners", "mtfb_get_state", "mtfb_get_similarity", "mtfb_get_init_warp",
                 "mtfb_get_init_pts", "mtfb_get_init_pix_vals"):
        getattr(L, name).argtypes = [vp, dp]
    L.mtfb_get_n_iters.argtypes = [vp, ip]
    L.mtfb_get_patch_status.argtypes = [vp, ip]
    L.mtfb_get_curr_stage.argtypes = [vp, dp, dp, dp, dp]
    L.mtfb_get_curr_stage_f32.argtypes = [vp, vp, vp, vp, vp, vp]
    L.mtfb_device_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.mtfb_state_size.argtypes = [vp]
    L.mtfb_debug_colpiv_qr_solve.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, ip, ip]
    L.mtfb_pf_default_params.argtypes = [C.POINTER(PFParams)]; L.mtfb_pf_default_params.restype = None
    L.mtfb_pf_configure.argtypes = [vp, C.POINTER(PFParams)]
    L.mtfb_pf_set_random_stream.argtypes = [vp, vp, vp]
    L.mtfb_pf_get_random_stream.argtypes = [vp, vp, vp]
    L.mtfb_pf_get_particles.argtypes = [vp, vp, vp, vp, vp]
    L.mtfb_est_default_params.argtypes = [C.POINTER(EstParams)]; L.mtfb_est_default_params.restype = None
    L.mtfb_estimate_warp_from_pts.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.POINTER(EstParams), vp, vp, vp, vp]
    L.mtfb_estimate_warp_from_corners_device.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.POINTER(EstParams), vp, vp, vp, vp]
    L.mtfb_grid_enable.argtypes = [vp]
    L.mtfb_grid_estimate.argtypes = [vp, C.c_int, C.POINTER(EstParams), vp, vp, vp, vp]
    L.mtfb_grid_commit.argtypes = [vp]
    L.mtfb_grid_initialize.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, vp]
    L.mtfb_grid_advance.argtypes = [vp, C.c_int, C.POINTER(EstParams), C.c_int, C.c_int, C.c_double, C.c_double, vp, vp, vp, vp, vp]
    L.mtfb_grid_get_pts.argtypes = [vp, vp, vp]
    L.mtfb_peer_export.argtypes = [vp, C.c_int, vp]
    L.mtfb_peer_attach.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    L.mtfb_peer_gather.argtypes = [vp]
    L.mtfb_peer_gathered.argtypes = [vp, C.POINTER(vp), ip]
    L.mtfb_get_gathered_region.argtypes = [vp, dp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("mtfb_last_error", "mtfb_version", "mtfb_launch_count", "mtfb_default_params", "mtfb_pf_default_params",
                        "mtfb_est_default_params"):
            fn.restype = C.c_int
    _lib = L
    return L


def default_params(**kw):
    p = Params()
    load_library().mtfb_default_params(C.byref(p))
    return set_params(p, **kw)


def set_params(p, **kw):
    for k, v in kw.items():
        if k == "am":
            v = AM[v] if isinstance(v, str) else v
        elif k == "ssm":
            v = SSM[v] if isinstance(v, str) else v
        elif k == "sm":
            v = SM[v] if isinstance(v, str) else v
        elif k == "hess_type" and isinstance(v, str):
            v = (ESM_HESS if p.sm == SM["esm"] else LK_HESS)[v]
        elif k == "jac_type" and isinstance(v, str):
            v = ESM_JAC[v]
        elif k == "precision" and isinstance(v, str):
            v = PRECISION[v]
        elif k == "f32_solve" and isinstance(v, str):
            v = F32_SOLVE[v]
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def make_params(am="ssd", ssm="homography", sm="fclk", **kw):
    """Parameters of one (SM, AM, SSM) combination with the per-SM default Hessian of the reference
    (ESMParams.cc:7 SumOfSelf, FCLKParams.cc:6 CurrentSelf, ICLKParams.cc:6 InitialSelf, FALKParams.cc:5 / IALKParams.cc:6
    InitialSelf)."""
    p = default_params(am=am, ssm=ssm, sm=sm)
    if "hess_type" not in kw:
        p.hess_type = {"esm": 2, "fclk": 1, "iclk": 0, "pf": 0, "falk": 0, "ialk": 0}[sm if isinstance(sm, str) else
                                                                  {v: k for k, v in SM.items()}[sm]]
    return set_params(p, **kw)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def debug_colpiv_qr_solve(A, b, fast=False, device=0):
    """the device's warp-level column-pivoted QR on host-supplied systems: A (n_sys, n, n), b (n_sys, n)
    -> x (n_sys, n), nonzero_pivots (n_sys,), perm (n_sys, n)"""
    L = load_library()
    A = np.asarray(A, dtype=np.float64)
    n_sys, n = A.shape[0], A.shape[-1]
    Ac = np.ascontiguousarray(A.transpose(0, 2, 1))          # column-major per system
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.empty((n_sys, n)); nz = np.empty(n_sys, dtype=np.int32); perm = np.empty((n_sys, n), dtype=np.int32)
    st = L.mtfb_debug_colpiv_qr_solve(device, n, int(fast), n_sys, _dp(Ac), _dp(b), _dp(x),
                                      nz.ctypes.data_as(C.POINTER(C.c_int)), perm.ctypes.data_as(C.POINTER(C.c_int)))
    if st != 0:
        raise MTFError(st, L.mtfb_last_error().decode())
    return x, nz, perm


class BatchTracker:
    """P independent patch trackers sharing one image: the batched counterpart of the reference's
    `vector<TrackerBase*>` fan-out (SM/src/GridTracker.cc:247-264).  Method names follow TrackerBase."""

    def __init__(self, params):
        self._L = load_library()
        self._h = C.c_void_p()
        self.params = params
        self._check(self._L.mtfb_create(C.byref(params), C.byref(self._h)))
        self.S = self._L.mtfb_state_size(self._h)
        self.N = params.resx * params.resy
        self.P = params.n_patches
        self._img = None

    # ---------------------------------------------------------------- plumbing
    def _check(self, status):
        if status != 0:
            raise MTFError(status, self._L.mtfb_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.mtfb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        self._check(self._L.mtfb_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._check(self._L.mtfb_synchronize(self._h))

    @property
    def launch_count(self):
        return self._L.mtfb_launch_count(self._h)

    # ---------------------------------------------------------------- TrackerBase
    def setImage(self, img):
        """img: float32 h x w NumPy array (CV_32FC1, inputType() of the reference trackers), or a CUDA
        float32 torch tensor (kept by reference, like the cv::Mat header ImageBase::setCurrImg keeps)."""
        if hasattr(img, "data_ptr"):      # torch tensor
            if not img.is_cuda or str(img.dtype) != "torch.float32" or img.dim() != 2 or img.stride(1) != 1:
                raise MTFError(1, "setImage: CUDA float32 2-D tensor with unit column stride required")
            if img.device.index is not None and img.device.index != self.params.device:
                raise MTFError(1, "setImage: the tensor lives on cuda:%d, the context on cuda:%d" % (img.device.index, self.params.device))
            self._img = img
            self._check(self._L.mtfb_set_image_device(self._h, C.c_void_p(img.data_ptr()), img.shape[0],
                                                      img.shape[1], img.stride(0)))
            return
        if img.dtype != np.float32 or img.ndim != 2 or img.strides[1] != 4:
            raise MTFError(1, "setImage: float32 2-D array with contiguous rows required")
        self._img = img
        self._check(self._L.mtfb_set_image(self._h, C.c_void_p(img.ctypes.data), img.shape[0], img.shape[1],
                                           img.strides[0] // 4))

    def setRawImage(self, img, kernel_size=5, sigma=3.0):
        """raw uint8 frame, h x w (gray) or h x w x 3 (BGR): the pre-processing MTF runs in front of setImage
        (utils::GaussianSmoothing, preprocUtils.cc:108-127) is done on the device behind the upload"""
        if img.dtype != np.uint8 or img.ndim not in (2, 3) or (img.ndim == 3 and (img.shape[2] != 3 or img.strides[2] != 1)) \
                or img.strides[1] != (1 if img.ndim == 2 else 3):
            raise MTFError(1, "setRawImage: uint8 h x w or h x w x 3 array with contiguous rows required")
        self._img = img
        self._check(self._L.mtfb_set_image_u8(self._h, C.c_void_p(img.ctypes.data), img.shape[0], img.shape[1],
                                              img.strides[0], 1 if img.ndim == 2 else 3, kernel_size, float(sigma)))

    def set_raw_image_pinned(self, ptr, h, w, row_stride, channels, kernel_size=5, sigma=3.0):
        self._check(self._L.mtfb_set_image_u8(self._h, C.c_void_p(ptr), h, w, row_stride, channels, kernel_size, float(sigma)))

    def image(self, h, w):
        """the float frame the trackers currently sample"""
        out = np.empty((h, w), dtype=np.float32)
        self._check(self._L.mtfb_get_image(self._h, C.c_void_p(out.ctypes.data)))
        return out

    def prefetch_image_pinned(self, ptr, h, w, row_stride):
        """mtfb_set_image_async: upload on the copy stream while the previous frame is tracked; the next update() samples it"""
        self._check(self._L.mtfb_set_image_async(self._h, C.c_void_p(ptr), h, w, row_stride))

    def prefetch_raw_image_pinned(self, ptr, h, w, row_stride, channels, kernel_size=5, sigma=3.0):
        self._check(self._L.mtfb_set_image_u8_async(self._h, C.c_void_p(ptr), h, w, row_stride, channels, kernel_size, float(sigma)))

    def set_image_pinned(self, ptr, h, w, row_stride):
        """host pointer variant (pinned buffers owned by the caller)"""
        self._check(self._L.mtfb_set_image(self._h, C.c_void_p(ptr), h, w, row_stride))

    def initialize(self, corners, img=None):
        if img is not None:
            self.setImage(img)
        c = np.ascontiguousarray(corners, dtype=np.float64).reshape(self.P, 8)
        self._check(self._L.mtfb_initialize(self._h, _dp(c)))

    def setRegion(self, corners, img=None):
        if img is not None:
            self.setImage(img)
        c = np.ascontiguousarray(corners, dtype=np.float64).reshape(self.P, 8)
        self._check(self._L.mtfb_set_region(self._h, _dp(c)))

    def update(self, img=None):
        if img is not None:
            self.setImage(img)
        self._check(self._L.mtfb_update(self._h))

    def getRegion(self):
        out = np.empty((self.P, 2, 4))
        self._check(self._L.mtfb_get_corners(self._h, _dp(out)))
        return out

    # ---------------------------------------------------------------- accessors
    def _get(self, name, shape, dtype=np.float64):
        out = np.empty(shape, dtype=dtype)
        ptr = out.ctypes.data_as(C.POINTER(C.c_double if dtype == np.float64 else C.c_int))
        self._check(getattr(self._L, name)(self._h, ptr))
        return out

    def state(self):
        return self._get("mtfb_get_state", (self.P, self.S))

    def n_iters(self):
        return self._get("mtfb_get_n_iters", (self.P,), np.int32)

    def similarity(self):
        return self._get("mtfb_get_similarity", (self.P,))

    def patch_status(self):
        return self._get("mtfb_get_patch_status", (self.P,), np.int32)

    def init_warp(self):
        return self._get("mtfb_get_init_warp", (self.P, 3, 3))

    def init_pts(self):
        return self._get("mtfb_get_init_pts", (self.P, self.N, 2))

    def init_pix_vals(self):
        return self._get("mtfb_get_init_pix_vals", (self.P, self.N))

    def curr_stage(self, pts=True, pix_vals=True, pix_grad=True, pix_jac=True):
        """(pts P x N x 2, It P x N, dIt_dx P x N x 2, dIt_dp P x N x S) at the current state."""
        a = np.empty((self.P, self.N, 2)) if pts else None
        b = np.empty((self.P, self.N)) if pix_vals else None
        g = np.empty((self.P, 2, self.N)) if pix_grad else None
        j = np.empty((self.P, self.S, self.N)) if pix_jac else None
        null = C.POINTER(C.c_double)()
        self._check(self._L.mtfb_get_curr_stage(self._h, _dp(a) if pts else null, _dp(b) if pix_vals else null,
                                                _dp(g) if pix_grad else null, _dp(j) if pix_jac else null))
        return (a, b, None if g is None else g.transpose(0, 2, 1), None if j is None else j.transpose(0, 2, 1))

    def curr_stage_f32(self):
        """precision='f32' contexts: (idx P x N x 2 int32 = (lx, ly) or -1, It P x N f32, dIt_dx P x N x 2 f32,
        dIt_dp P x N x S f64 in the reference's basis, fast_err P x N f32) at the current state."""
        idx = np.empty((self.P, self.N, 2), dtype=np.int32)
        val = np.empty((self.P, self.N), dtype=np.float32)
        g = np.empty((self.P, 2, self.N), dtype=np.float32)
        j = np.empty((self.P, self.S, self.N))
        err = np.empty((self.P, self.N), dtype=np.float32)
        self._check(self._L.mtfb_get_curr_stage_f32(self._h, idx.ctypes.data, val.ctypes.data, g.ctypes.data,
                                                    j.ctypes.data, err.ctypes.data))
        return idx, val, g.transpose(0, 2, 1), j.transpose(0, 2, 1), err

    def iterate_once(self):
        """one Gauss-Newton pass: (J P x S, H P x S x S, f P, dp P x S)"""
        J = np.empty((self.P, self.S)); H = np.empty((self.P, self.S, self.S)); f = np.empty(self.P)
        dp = np.empty((self.P, self.S))
        self._check(self._L.mtfb_iterate_once(self._h, _dp(J), _dp(H), _dp(f), _dp(dp)))
        return J, H.transpose(0, 2, 1), f, dp      # column-major S x S -> [i][j]

    def enable_iter_log(self, n_slots):
        self._log_slots = n_slots
        self._check(self._L.mtfb_enable_iter_log(self._h, n_slots))

    def iter_log(self):
        n = self._log_slots
        arr = (IterLog * (self.P * n))()
        self._check(self._L.mtfb_get_iter_log(self._h, arr))
        S = self.S
        out = []
        for p in range(self.P):
            rows = []
            for i in range(n):
                e = arr[p * n + i]
                if not e.valid:
                    break
                rows.append(dict(f=e.f, jacobian=np.array(e.jacobian[:S]),
                                 hessian=np.array(e.hessian[:S * S]).reshape(S, S).T,
                                 state_update=np.array(e.state_update[:S]),
                                 corners=np.array(e.corners[:]).reshape(2, 4),
                                 update_norm=e.update_norm, rejected=bool(e.rejected)))
            out.append(rows)
        return out

    def pf_evaluate(self, states):
        states = np.ascontiguousarray(states, dtype=np.float64)
        n = states.shape[1]
        assert states.shape == (self.P, n, self.S)
        lik = np.empty((self.P, n)); sim = np.empty((self.P, n))
        self._check(self._L.mtfb_pf_evaluate(self._h, _dp(states), n, _dp(lik), _dp(sim)))
        return lik, sim

    # ---- robust warp estimation from point pairs (SSM::estimateWarpFromPts; mtf_b200/csrc/grid_estimator.cu)
    def _est_result(self, ssm, n, call):
        S = 8 if ssm == 0 else 6
        su = np.zeros(S); mask = np.zeros(n, dtype=np.uint8); warp = np.zeros(9); info = np.zeros(4, dtype=np.int32)
        self._check(call(su.ctypes.data, mask.ctypes.data, warp.ctypes.data, info.ctypes.data))
        return {"ok": bool(info[0]), "state_update": su, "mask": mask, "warp": warp.reshape(3, 3), "drawn": int(info[1]),
                "n_inliers": int(info[2]), "lm_evals": int(info[3])}

    def estimate_warp_from_pts(self, ssm, in_pts, out_pts, est_params):
        """ssm.estimateWarpFromPts(state_update, mask, in_pts, out_pts, est_params) for Homography / Affine"""
        ssm = SSM[ssm] if isinstance(ssm, str) else int(ssm)
        a = np.ascontiguousarray(in_pts, dtype=np.float32).reshape(-1, 2)
        b = np.ascontiguousarray(out_pts, dtype=np.float32).reshape(-1, 2)
        if a.shape != b.shape:
            raise ValueError("in_pts and out_pts differ in size")
        n = a.shape[0]
        return self._est_result(ssm, n, lambda su, mk, wp, inf: self._L.mtfb_estimate_warp_from_pts(
            self._h, ssm, a.ctypes.data, b.ctypes.data, n, C.byref(est_params), su, mk, wp, inf))

    def estimate_warp_from_corners_device(self, ssm, d_in_corners, d_out_corners, n, est_params):
        """the same from two device arrays (raw pointers) of n x 8 corners: the points are their centroids"""
        ssm = SSM[ssm] if isinstance(ssm, str) else int(ssm)
        return self._est_result(ssm, n, lambda su, mk, wp, inf: self._L.mtfb_estimate_warp_from_corners_device(
            self._h, ssm, d_in_corners, d_out_corners, n, C.byref(est_params), su, mk, wp, inf))

    def grid_enable(self):
        self._check(self._L.mtfb_grid_enable(self._h))

    def grid_estimate(self, ssm, est_params):
        """the same between the centroids at the last initialize / setRegion / grid_commit and the current ones, on the device"""
        ssm = SSM[ssm] if isinstance(ssm, str) else int(ssm)
        return self._est_result(ssm, self.P, lambda su, mk, wp, inf: self._L.mtfb_grid_estimate(
            self._h, ssm, C.byref(est_params), su, mk, wp, inf))

    def grid_initialize(self, grid_size_x, grid_size_y, patch_size_x, patch_size_y, region):
        """GridTracker::initialize for the shipped grid on the device: cell layout for `region` + every cell initialised there"""
        reg = np.ascontiguousarray(region, dtype=np.float64).reshape(8)
        self._check(self._L.mtfb_grid_initialize(self._h, int(grid_size_x), int(grid_size_y), float(patch_size_x), float(patch_size_y),
                                                 reg.ctypes.data))

    def grid_advance(self, ssm, est_params, grid_size_x, grid_size_y, patch_size_x, patch_size_y, region):
        """grid_estimate + the region's update + the cell layout + the cells' re-initialisation, all on the device
        (mtfb_grid_advance); -> (estimate dict, new region (2, 4))"""
        ssm = SSM[ssm] if isinstance(ssm, str) else int(ssm)
        reg = np.array(region, dtype=np.float64).reshape(8).copy()
        est = self._est_result(ssm, self.P, lambda su, mk, wp, inf: self._L.mtfb_grid_advance(
            self._h, ssm, C.byref(est_params), int(grid_size_x), int(grid_size_y), float(patch_size_x), float(patch_size_y),
            reg.ctypes.data, su, mk, wp, inf))
        return est, reg.reshape(2, 4)

    def grid_commit(self):
        self._check(self._L.mtfb_grid_commit(self._h))

    def grid_pts(self):
        a = np.empty((self.P, 2), dtype=np.float32); b = np.empty((self.P, 2), dtype=np.float32)
        self._check(self._L.mtfb_grid_get_pts(self._h, a.ctypes.data, b.ctypes.data))
        return a, b

    # ---------------------------------------------------------------- multi-GPU exchange over NVLink peer memory
    def peer_export(self, n_total):
        """allocate this rank's gathered array for a job of n_total patches; returns its 64-byte CUDA IPC handle"""
        h = np.zeros(PEER_HANDLE_BYTES, dtype=np.uint8)
        self._check(self._L.mtfb_peer_export(self._h, int(n_total), h.ctypes.data))
        self._peer_n_total = int(n_total)
        return h

    def peer_attach(self, rank, world, row0, handles):
        """handles: (world, 64) uint8, the ranks' peer_export() results in rank order"""
        h = np.ascontiguousarray(np.asarray(handles, dtype=np.uint8).reshape(world, PEER_HANDLE_BYTES))
        self._check(self._L.mtfb_peer_attach(self._h, int(rank), int(world), int(row0), h.ctypes.data))

    def peer_gather(self):
        """after update(): signal the peers and wait for theirs (stream-ordered)"""
        self._check(self._L.mtfb_peer_gather(self._h))

    def peer_gathered_ptr(self):
        a, n = C.c_void_p(), C.c_int()
        self._check(self._L.mtfb_peer_gathered(self._h, C.byref(a), C.byref(n)))
        return a.value, n.value

    def getGatheredRegion(self):
        out = np.empty((self._peer_n_total, 2, 4))
        self._check(self._L.mtfb_get_gathered_region(self._h, _dp(out)))
        return out

    def device_results(self):
        """raw device pointers (corners P x 8 f64, state P x S f64, n_iters P i32)"""
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self._L.mtfb_device_results(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value


def make_est_params(method="ransac", seed=0, **kw):
    """mtfb_est_params: SSMEstimatorParams' defaults (SSMEstimatorParams.cc:5-13) with overrides"""
    p = EstParams()
    load_library().mtfb_est_default_params(C.byref(p))
    p.method = EST_METHOD[method] if isinstance(method, str) else int(method)
    p.seed = int(seed)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def make_pf_params(n_particles=500, sigma=None, mean=None, **kw):
    """mtfb_pf_params: the shipped configuration (Config/modules.cfg:152-177) with the sampler's sigma / mean set"""
    p = PFParams()
    load_library().mtfb_pf_default_params(C.byref(p))
    p.n_particles = int(n_particles)
    names = {"dynamic_model": PF_DYNAMIC, "update_type": PF_UPDATE, "likelihood_func": PF_LIKELIHOOD,
             "resampling_type": PF_RESAMPLING, "mean_type": PF_MEAN}
    if sigma is not None:
        sg = np.broadcast_to(np.asarray(sigma, dtype=np.float64), (8,)) if np.ndim(sigma) == 0 else np.asarray(sigma, dtype=np.float64)
        for i in range(min(8, len(sg))):
            p.ssm_sigma[i] = float(sg[i])
    if mean is not None:
        mn = np.broadcast_to(np.asarray(mean, dtype=np.float64), (8,)) if np.ndim(mean) == 0 else np.asarray(mean, dtype=np.float64)
        for i in range(min(8, len(mn))):
            p.ssm_mean[i] = float(mn[i])
    for k, v in kw.items():
        if k in names and isinstance(v, str):
            v = names[k][v]
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


class PFTracker(BatchTracker):
    """P particle-filter trackers (nt::PF, SM/src/NT/PF.cc) sharing one image: initialize / update / setRegion / getRegion as
    TrackerBase; every stage of PF::update runs on the device (mtf_b200/csrc/pf_tracker.cu)."""

    def __init__(self, params, n_particles=500, sigma=None, mean=None, **pf_kw):
        super().__init__(params)
        self.pf_params = make_pf_params(n_particles, sigma, mean, **pf_kw)
        self._check(self._L.mtfb_pf_configure(self._h, C.byref(self.pf_params)))
        self.n_particles = self.pf_params.n_particles
        self.n_normals = 10 if (self.pf_params.corner_based_sampling and self.S == 8) else self.S

    def set_random_stream(self, normals=None, uniforms=None):
        """host-supplied deviates for the next update(): normals (max_iters, P, n, R), uniforms (max_iters, P, n)"""
        it = self.pf_params.max_iters
        a = None if normals is None else np.ascontiguousarray(normals, dtype=np.float64).reshape(it, self.P, self.n_particles, self.n_normals)
        b = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64).reshape(it, self.P, self.n_particles)
        self._check(self._L.mtfb_pf_set_random_stream(self._h, None if a is None else a.ctypes.data, None if b is None else b.ctypes.data))

    def random_stream(self):
        it = self.pf_params.max_iters
        a = np.empty((it, self.P, self.n_particles, self.n_normals)); b = np.empty((it, self.P, self.n_particles))
        self._check(self._L.mtfb_pf_get_random_stream(self._h, a.ctypes.data, b.ctypes.data))
        return a, b

    def particles(self):
        """(states (P, n, S), weights (P, n), cum_weights (P, n), max_wt_id (P,)) after the last update"""
        st = np.empty((self.P, self.n_particles, self.S)); w = np.empty((self.P, self.n_particles)); cw = np.empty_like(w)
        mx = np.empty(self.P, dtype=np.int32)
        self._check(self._L.mtfb_pf_get_particles(self._h, st.ctypes.data, w.ctypes.data, cw.ctypes.data, mx.ctypes.data))
        return st, w, cw, mx
