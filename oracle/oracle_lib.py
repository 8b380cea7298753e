"""ctypes binding of the CPU oracle (oracle/mtf_oracle.h).  Test infrastructure only."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = _HERE
# MTF_ORACLE_GENERIC=1: the build without -march=native (CPU-baseline row of bench.py --impl reference)
_LIB_PATH = os.path.join(_ORACLE_DIR, "_build", "libmtf_oracle_generic.so" if os.environ.get("MTF_ORACLE_GENERIC") == "1"
                         else "libmtf_oracle.so")

AM = {"ssd": 0, "ncc": 1, "mi": 2}
SSM = {"homography": 0, "affine": 1, "translation": 2}
SM = {"esm": 0, "fclk": 1, "iclk": 2, "falk": 4, "ialk": 5}


class OrcParams(C.Structure):
    _fields_ = [("am", C.c_int), ("ssm", C.c_int), ("sm", C.c_int),
                ("resx", C.c_int), ("resy", C.c_int), ("max_iters", C.c_int),
                ("epsilon", C.c_double), ("hess_type", C.c_int), ("jac_type", C.c_int),
                ("chained_warp", C.c_int), ("leven_marq", C.c_int),
                ("lm_delta_init", C.c_double), ("lm_delta_update", C.c_double),
                ("nt_semantics", C.c_int), ("grad_eps", C.c_double),
                ("hom_normalized_init", C.c_int), ("mi_n_bins", C.c_int),
                ("mi_pre_seed", C.c_double), ("mi_pou", C.c_int),
                ("likelihood_alpha", C.c_double), ("grad_mode", C.c_int), ("fast_sums", C.c_int)]


class OrcPFParams(C.Structure):
    _fields_ = [("n_particles", C.c_int), ("max_iters", C.c_int), ("epsilon", C.c_double),
                ("dynamic_model", C.c_int), ("update_type", C.c_int), ("likelihood_func", C.c_int),
                ("resampling_type", C.c_int), ("mean_type", C.c_int), ("reset_to_mean", C.c_int),
                ("adaptive_resampling_thresh", C.c_double), ("measurement_sigma", C.c_double), ("ar_coeff", C.c_double),
                ("ssm_sigma", C.c_double * 8), ("ssm_mean", C.c_double * 8), ("corner_based_sampling", C.c_int)]


class OrcEstParams(C.Structure):
    """orc_est_params: SSMEstimatorParams (SSM/src/SSMEstimatorParams.cc) + the cvRNG seed"""
    _fields_ = [("method", C.c_int), ("ransac_reproj_thresh", C.c_double), ("n_model_pts", C.c_int), ("refine", C.c_int),
                ("max_iters", C.c_int), ("max_subset_attempts", C.c_int), ("confidence", C.c_double),
                ("lm_max_iters", C.c_int), ("seed", C.c_ulonglong)]


EST_METHOD = {"ransac": 0, "lmeds": 1, "least_squares": 2}


class OrcIterLog(C.Structure):
    _fields_ = [("f", C.c_double), ("jacobian", C.c_double * 8), ("hessian", C.c_double * 64),
                ("state_update", C.c_double * 8), ("corners", C.c_double * 8),
                ("update_norm", C.c_double), ("rejected", C.c_int)]


def build():
    """(Re)build the oracle shared library if its sources are newer than the binary."""
    src = [os.path.join(_ORACLE_DIR, f) for f in ("mtf_oracle.cpp", "mtf_oracle_est.cpp", "mtf_oracle.h", "Makefile")]
    if os.path.exists(_LIB_PATH) and all(os.path.getmtime(s) <= os.path.getmtime(_LIB_PATH) for s in src):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _ORACLE_DIR], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.orc_default_params.argtypes = [C.POINTER(OrcParams)]
    L.orc_create.argtypes = [C.POINTER(OrcParams)]; L.orc_create.restype = C.c_void_p
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_set_image.argtypes = [C.c_void_p, fp, C.c_int, C.c_int]
    L.orc_initialize.argtypes = [C.c_void_p, dp]; L.orc_initialize.restype = C.c_int
    L.orc_update.argtypes = [C.c_void_p]; L.orc_update.restype = C.c_int
    L.orc_set_region.argtypes = [C.c_void_p, dp]; L.orc_set_region.restype = C.c_int
    L.orc_n_iters.argtypes = [C.c_void_p]; L.orc_n_iters.restype = C.c_int
    L.orc_n_log.argtypes = [C.c_void_p]; L.orc_n_log.restype = C.c_int
    L.orc_log.argtypes = [C.c_void_p, C.c_int]; L.orc_log.restype = C.POINTER(OrcIterLog)
    L.orc_state_size.argtypes = [C.c_void_p]; L.orc_state_size.restype = C.c_int
    L.orc_get_similarity.argtypes = [C.c_void_p]; L.orc_get_similarity.restype = C.c_double
    for name in ("orc_get_corners", "orc_get_state", "orc_get_pts", "orc_get_init_pts", "orc_get_init_pix_vals",
                 "orc_get_curr_pix_vals", "orc_get_curr_pix_grad", "orc_get_curr_pix_jacobian",
                 "orc_get_init_pix_jacobian", "orc_get_init_warp", "orc_get_stage_times"):
        getattr(L, name).argtypes = [C.c_void_p, dp]
    L.orc_pf_evaluate.argtypes = [C.c_void_p, dp, C.c_int, dp, dp]
    L.orc_set_state.argtypes = [C.c_void_p, dp]
    L.orc_gaussian_kernel5.argtypes = [C.c_double, fp]
    L.orc_preproc_gauss5.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, fp]
    L.orc_pix_val.argtypes = [fp, C.c_int, C.c_int, C.c_double, C.c_double]; L.orc_pix_val.restype = C.c_double
    L.orc_get_pix_vals.argtypes = [fp, C.c_int, C.c_int, dp, C.c_int, C.c_double, C.c_double, dp]
    L.orc_get_img_grad.argtypes = [fp, C.c_int, C.c_int, dp, C.c_int, C.c_double, C.c_double, dp]
    L.orc_get_img_grad_analytic.argtypes = [fp, C.c_int, C.c_int, dp, C.c_int, C.c_double, C.c_double, dp]
    L.orc_homography_dlt.argtypes = [dp, dp, dp]
    L.orc_affine_ndlt.argtypes = [dp, dp, dp]
    L.orc_colpiv_qr_solve.argtypes = [dp, dp, C.c_int, dp]
    L.orc_colpiv_qr.argtypes = [dp, C.c_int, C.c_int, dp, ip, dp]; L.orc_colpiv_qr.restype = C.c_int
    L.orc_norm_unit_square_pts.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, dp, dp]
    L.orc_batch_track.argtypes = [C.POINTER(OrcParams), C.POINTER(fp), C.c_int, C.c_int, C.c_int, dp, C.c_int,
                                  C.c_int, dp, ip, dp]
    L.orc_batch_track.restype = C.c_long
    L.orc_batch_track_reset.argtypes = L.orc_batch_track.argtypes
    L.orc_batch_track_reset.restype = C.c_long
    L.orc_batch_pf_evaluate.argtypes = [C.POINTER(OrcParams), fp, fp, C.c_int, C.c_int, dp, C.c_int, dp, C.c_int, C.c_int, dp, dp]
    L.orc_batch_pf_evaluate.restype = C.c_long
    L.orc_pf_create.argtypes = [C.POINTER(OrcParams), C.POINTER(OrcPFParams)]; L.orc_pf_create.restype = C.c_void_p
    L.orc_pf_destroy.argtypes = [C.c_void_p]
    L.orc_pf_set_image.argtypes = [C.c_void_p, fp, C.c_int, C.c_int]
    L.orc_pf_initialize.argtypes = [C.c_void_p, dp]
    L.orc_pf_update.argtypes = [C.c_void_p, dp, dp]
    L.orc_pf_get_corners.argtypes = [C.c_void_p, dp]
    L.orc_pf_get_state.argtypes = [C.c_void_p, dp]
    L.orc_pf_n_normals.argtypes = [C.c_void_p]; L.orc_pf_n_normals.restype = C.c_int
    L.orc_pf_get_particles.argtypes = [C.c_void_p, dp, dp, dp, ip]; L.orc_pf_get_particles.restype = C.c_int
    L.orc_est_default_params.argtypes = [C.POINTER(OrcEstParams)]
    L.orc_estimate_warp.argtypes = [C.c_int, fp, fp, C.c_int, C.POINTER(OrcEstParams), dp, C.POINTER(C.c_ubyte), dp, ip]
    L.orc_estimate_warp.restype = C.c_int
    L.orc_est_subsets.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_int, ip]; L.orc_est_subsets.restype = C.c_int
    L.orc_cv_rand_int.argtypes = [C.POINTER(C.c_ulonglong)]; L.orc_cv_rand_int.restype = C.c_uint
    L.orc_sym_eigen.argtypes = [dp, C.c_int, dp, dp]
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def make_params(am="ssd", ssm="homography", sm="fclk", **kw):
    p = OrcParams()
    lib().orc_default_params(C.byref(p))
    p.am, p.ssm, p.sm = AM[am], SSM[ssm], SM[sm]
    if sm == "esm" and "hess_type" not in kw:
        p.hess_type = 2  # SumOfSelf (ESMParams.cc:7)
    if sm in ("iclk", "falk", "ialk") and "hess_type" not in kw:
        p.hess_type = 0  # InitialSelf (ICLKParams.cc:6, FALKParams.cc:5, IALKParams.cc:6)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


class OracleTracker:
    """One reference-style tracker (SM + AM + SSM) evaluated by the CPU oracle."""

    def __init__(self, params):
        self.p = params
        self.h = lib().orc_create(C.byref(params))
        self.S = lib().orc_state_size(self.h)
        self.N = params.resx * params.resy
        self._img = None

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    def set_image(self, img):
        img = np.ascontiguousarray(img, dtype=np.float32)
        self._img = img  # the oracle keeps the pointer, like ImageBase::setCurrImg
        lib().orc_set_image(self.h, _fp(img), img.shape[0], img.shape[1])

    def initialize(self, corners):
        c = np.ascontiguousarray(corners, dtype=np.float64).reshape(8)
        return lib().orc_initialize(self.h, _dp(c))

    def set_region(self, corners):
        c = np.ascontiguousarray(corners, dtype=np.float64).reshape(8)
        return lib().orc_set_region(self.h, _dp(c))

    def update(self):
        return lib().orc_update(self.h)

    def _get(self, name, n):
        out = np.empty(n, dtype=np.float64)
        getattr(lib(), name)(self.h, _dp(out))
        return out

    @property
    def n_iters(self):
        return lib().orc_n_iters(self.h)

    @property
    def similarity(self):
        return lib().orc_get_similarity(self.h)

    def corners(self):
        return self._get("orc_get_corners", 8).reshape(2, 4)

    def state(self):
        return self._get("orc_get_state", self.S)

    def set_state(self, state):
        """ssm.setState: warp the template points with `state` (no appearance update)"""
        s = np.ascontiguousarray(state, dtype=np.float64)
        assert s.size == self.S
        lib().orc_set_state(self.h, _dp(s))

    def pts(self):
        return self._get("orc_get_pts", 2 * self.N).reshape(self.N, 2)

    def init_pts(self):
        return self._get("orc_get_init_pts", 2 * self.N).reshape(self.N, 2)

    def init_pix_vals(self):
        return self._get("orc_get_init_pix_vals", self.N)

    def curr_pix_vals(self):
        return self._get("orc_get_curr_pix_vals", self.N)

    def curr_pix_grad(self):
        return self._get("orc_get_curr_pix_grad", 2 * self.N).reshape(2, self.N).T

    def curr_pix_jacobian(self):
        return self._get("orc_get_curr_pix_jacobian", self.N * self.S).reshape(self.S, self.N).T

    def init_pix_jacobian(self):
        return self._get("orc_get_init_pix_jacobian", self.N * self.S).reshape(self.S, self.N).T

    def init_warp(self):
        return self._get("orc_get_init_warp", 9).reshape(3, 3)

    def stage_times(self):
        return self._get("orc_get_stage_times", 9)

    def log(self):
        out = []
        S = self.S
        for i in range(lib().orc_n_log(self.h)):
            e = lib().orc_log(self.h, i).contents
            out.append(dict(f=e.f, jacobian=np.array(e.jacobian[:S]),
                            hessian=np.array(e.hessian[:S * S]).reshape(S, S).T,
                            state_update=np.array(e.state_update[:S]),
                            corners=np.array(e.corners[:]).reshape(2, 4),
                            update_norm=e.update_norm, rejected=bool(e.rejected)))
        return out

    def pf_evaluate(self, states):
        states = np.ascontiguousarray(states, dtype=np.float64)
        n = states.shape[0]
        lik = np.empty(n); sim = np.empty(n)
        lib().orc_pf_evaluate(self.h, _dp(states), n, _dp(lik), _dp(sim))
        return lik, sim


def pix_vals(img, pts, norm_mult=1.0, norm_add=0.0):
    img = np.ascontiguousarray(img, dtype=np.float32)
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    out = np.empty(pts.shape[0])
    lib().orc_get_pix_vals(_fp(img), img.shape[0], img.shape[1], _dp(pts), pts.shape[0], norm_mult, norm_add, _dp(out))
    return out


def img_grad(img, pts, grad_eps=1e-8, mult=1.0):
    img = np.ascontiguousarray(img, dtype=np.float32)
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    out = np.empty(2 * pts.shape[0])
    lib().orc_get_img_grad(_fp(img), img.shape[0], img.shape[1], _dp(pts), pts.shape[0], grad_eps, mult, _dp(out))
    return out.reshape(2, -1).T


def img_grad_analytic(img, pts, grad_eps=1e-8, mult=1.0):
    img = np.ascontiguousarray(img, dtype=np.float32)
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    out = np.empty(2 * pts.shape[0])
    lib().orc_get_img_grad_analytic(_fp(img), img.shape[0], img.shape[1], _dp(pts), pts.shape[0], grad_eps, mult, _dp(out))
    return out.reshape(2, -1).T


def gaussian_kernel5(sigma=3.0):
    k = np.empty(5, dtype=np.float32)
    lib().orc_gaussian_kernel5(float(sigma), _fp(k))
    return k


def preproc_gauss5(img_u8, sigma=3.0):
    """uint8 h x w (gray) or h x w x 3 (BGR) -> float32 h x w: convertTo + BGR2GRAY + GaussianBlur(5, sigma)"""
    a = np.ascontiguousarray(img_u8, dtype=np.uint8)
    ch = 1 if a.ndim == 2 else a.shape[2]
    h, w = a.shape[:2]
    out = np.empty((h, w), dtype=np.float32)
    lib().orc_preproc_gauss5(a.ctypes.data, h, w, a.strides[0], ch, float(sigma), _fp(out))
    return out


def homography_dlt(in_corners, out_corners):
    a = np.ascontiguousarray(in_corners, dtype=np.float64).reshape(8)
    b = np.ascontiguousarray(out_corners, dtype=np.float64).reshape(8)
    H = np.empty(9)
    lib().orc_homography_dlt(_dp(a), _dp(b), _dp(H))
    return H.reshape(3, 3)


def affine_ndlt(in_corners, out_corners):
    """utils::computeAffineNDLT (warpUtils.cc:378-386)"""
    a = np.ascontiguousarray(in_corners, dtype=np.float64).reshape(8)
    b = np.ascontiguousarray(out_corners, dtype=np.float64).reshape(8)
    H = np.empty(9)
    lib().orc_affine_ndlt(_dp(a), _dp(b), _dp(H))
    return H.reshape(3, 3)


def colpiv_qr_solve(A, b):
    A = np.asfortranarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.empty(A.shape[0])
    lib().orc_colpiv_qr_solve(A.ctypes.data_as(C.POINTER(C.c_double)), _dp(b), A.shape[0], _dp(x))
    return x


def colpiv_qr(A):
    """(R upper-triangular rows x cols, perm, householder coefficients, nonzero_pivots) of Eigen's ColPivHouseholderQR"""
    A = np.asfortranarray(A, dtype=np.float64)
    rows, cols = A.shape
    qr = np.empty((rows, cols), order="F"); perm = np.empty(cols, dtype=np.int32); tau = np.empty(min(rows, cols))
    nz = lib().orc_colpiv_qr(A.ctypes.data_as(C.POINTER(C.c_double)), rows, cols, qr.ctypes.data_as(C.POINTER(C.c_double)),
                             perm.ctypes.data_as(C.POINTER(C.c_int)), _dp(tau))
    return np.triu(qr), perm, tau, nz


def norm_unit_square_pts(resx, resy, min_x=-0.5, min_y=-0.5, max_x=0.5, max_y=0.5):
    pts = np.empty(2 * resx * resy); c = np.empty(8)
    lib().orc_norm_unit_square_pts(resx, resy, min_x, min_y, max_x, max_y, _dp(pts), _dp(c))
    return pts.reshape(-1, 2), c.reshape(2, 4)


def batch_track(params, frames, corners, n_threads=0, reset_each_frame=False):
    """CPU baseline driver: P independent trackers, OpenMP over patches (GridTracker.cc:253-256); reset_each_frame: every
    tracker is re-initialised at its initial corners after each frame (GridTracker's grid_reset_at_each_frame)."""
    frames = [np.ascontiguousarray(f, dtype=np.float32) for f in frames]
    h, w = frames[0].shape
    arr = (C.POINTER(C.c_float) * len(frames))(*[_fp(f) for f in frames])
    corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(-1, 8)
    P = corners.shape[0]
    final = np.empty((P, 8)); iters = np.zeros(P, dtype=np.int32); secs = C.c_double(0)
    fn = lib().orc_batch_track_reset if reset_each_frame else lib().orc_batch_track
    total = fn(C.byref(params), arr, len(frames), h, w, _dp(corners), P, n_threads,
               _dp(final), iters.ctypes.data_as(C.POINTER(C.c_int)), C.byref(secs))
    return total, secs.value, final.reshape(P, 2, 4), iters


def batch_pf_evaluate(params, frame0, frame1, corners, states, n_threads=0):
    """CPU baseline driver of the PF particle loop: (evaluations, seconds, likelihood (n_objects, n_particles))"""
    f0 = np.ascontiguousarray(frame0, dtype=np.float32); f1 = np.ascontiguousarray(frame1, dtype=np.float32)
    h, w = f0.shape
    corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(-1, 8)
    states = np.ascontiguousarray(states, dtype=np.float64)
    n_obj, n_part = states.shape[0], states.shape[1]
    lik = np.empty((n_obj, n_part)); secs = C.c_double(0)
    n = lib().orc_batch_pf_evaluate(C.byref(params), _fp(f0), _fp(f1), h, w, _dp(corners), n_obj, _dp(states), n_part, n_threads,
                                    _dp(lik), C.byref(secs))
    return n, secs.value, lik


class OraclePF:
    """nt::PF restated (oracle/mtf_oracle.cpp orc_pf) for one object; the caller supplies the random deviates"""

    def __init__(self, params, pf_params):
        """pf_params: any structure with the OrcPFParams field names (e.g. mtf_b200.api.PFParams)"""
        self._L = lib()
        q = OrcPFParams()
        for name, _ in OrcPFParams._fields_:
            v = getattr(pf_params, name)
            if name in ("ssm_sigma", "ssm_mean"):
                for i in range(8):
                    getattr(q, name)[i] = v[i]
            else:
                setattr(q, name, v)
        self.pf_params = q
        self._h = self._L.orc_pf_create(C.byref(params), C.byref(q))
        self.S = 8 if params.ssm == 0 else 6
        self.n = q.n_particles
        self.n_normals = self._L.orc_pf_n_normals(self._h)
        self._img = None

    def __del__(self):
        try:
            self._L.orc_pf_destroy(self._h)
        except Exception:
            pass

    def set_image(self, img):
        self._img = np.ascontiguousarray(img, dtype=np.float32)
        self._L.orc_pf_set_image(self._h, _fp(self._img), self._img.shape[0], self._img.shape[1])

    def initialize(self, corners):
        c = np.ascontiguousarray(corners, dtype=np.float64).reshape(8)
        return self._L.orc_pf_initialize(self._h, _dp(c))

    def update(self, normals, uniforms):
        a = np.ascontiguousarray(normals, dtype=np.float64).reshape(self.pf_params.max_iters, self.n, self.n_normals)
        b = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(self.pf_params.max_iters, self.n)
        return self._L.orc_pf_update(self._h, _dp(a), _dp(b))

    def corners(self):
        out = np.empty(8); self._L.orc_pf_get_corners(self._h, _dp(out)); return out.reshape(2, 4)

    def state(self):
        out = np.empty(self.S); self._L.orc_pf_get_state(self._h, _dp(out)); return out

    def particles(self):
        st = np.empty((self.n, self.S)); w = np.empty(self.n); cw = np.empty(self.n); r = C.c_int(0)
        mx = self._L.orc_pf_get_particles(self._h, _dp(st), _dp(w), _dp(cw), C.byref(r))
        return st, w, cw, mx, bool(r.value)


def make_est_params(method="ransac", seed=0, **kw):
    """SSMEstimatorParams defaults (SSMEstimatorParams.cc:5-13) with overrides"""
    q = OrcEstParams()
    lib().orc_est_default_params(C.byref(q))
    q.method = EST_METHOD[method] if isinstance(method, str) else int(method)
    q.seed = seed
    for k, v in kw.items():
        if not hasattr(q, k):
            raise KeyError(k)
        setattr(q, k, v)
    return q


def estimate_warp(ssm, in_pts, out_pts, est_params):
    """SSM::estimateWarpFromPts for Homography / Affine.  in_pts, out_pts: n x 2 (cv::Point2f).  est_params: any structure
    with the OrcEstParams field names.  Returns dict(ok, warp 3x3, mask, state_update, drawn, n_inliers, lm_evals)."""
    L = lib()
    q = OrcEstParams()
    for name, _ in OrcEstParams._fields_:
        setattr(q, name, getattr(est_params, name))
    a = np.ascontiguousarray(in_pts, dtype=np.float32).reshape(-1, 2)
    b = np.ascontiguousarray(out_pts, dtype=np.float32).reshape(-1, 2)
    n = a.shape[0]
    S = 8 if SSM.get(ssm, ssm) == 0 else 6
    warp = np.zeros(9); mask = np.zeros(n, dtype=np.uint8); su = np.zeros(S); info = np.zeros(4, dtype=np.int32)
    rc = L.orc_estimate_warp(SSM.get(ssm, ssm), _fp(a), _fp(b), n, C.byref(q), _dp(warp),
                             mask.ctypes.data_as(C.POINTER(C.c_ubyte)), _dp(su), info.ctypes.data_as(C.POINTER(C.c_int)))
    if rc < 0:
        raise ValueError("orc_estimate_warp: invalid arguments")
    return {"ok": bool(rc), "warp": warp.reshape(3, 3), "mask": mask, "state_update": su, "drawn": int(info[1]),
            "n_inliers": int(info[2]), "lm_evals": int(info[3])}


def est_subsets(in_pts, out_pts, model_points, max_attempts, seed, n_subsets):
    a = np.ascontiguousarray(in_pts, dtype=np.float32).reshape(-1, 2)
    b = np.ascontiguousarray(out_pts, dtype=np.float32).reshape(-1, 2)
    idx = np.zeros((n_subsets, model_points), dtype=np.int32)
    k = lib().orc_est_subsets(_fp(a), _fp(b), a.shape[0], model_points, max_attempts, seed, n_subsets,
                              idx.ctypes.data_as(C.POINTER(C.c_int)))
    return k, idx


def cv_rand_ints(seed, n):
    st = C.c_ulonglong(seed if seed else 0xFFFFFFFFFFFFFFFF)
    return np.array([lib().orc_cv_rand_int(C.byref(st)) for _ in range(n)], dtype=np.uint32)


def sym_eigen(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    V = np.zeros((n, n)); w = np.zeros(n)
    lib().orc_sym_eigen(_dp(A), n, _dp(V), _dp(w))
    return w, V


class OracleGrid:
    """GridTracker<SSM> (SM/src/GridTracker.cc:232-285, 345-392) restated on the oracle's own pieces: one OracleTracker per
    cell, orc_estimate_warp for ssm.estimateWarpFromPts, orc_homography_dlt / orc_norm_unit_square_pts for ssm.setCorners.
    Forward-backward error estimation (fb_err_thresh > 0: GridTracker.cc:292-343) included; no pyramids.  est_params: any
    structure with the OrcEstParams fields; frame t uses seed + t."""

    def __init__(self, cell_params, grid_size_x, grid_size_y, patch_size_x, patch_size_y, reset_at_each_frame=1, dyn_patch_size=0,
                 patch_centroid_inside=True, ssm="homography", est_params=None, seed=1, fb_err_thresh=0, fb_reinit=0):
        self.fb_err_thresh, self.fb_reinit = float(fb_err_thresh), bool(fb_reinit)
        self.enable_fb = self.fb_err_thresh > 0                                # GridTracker.cc:186-189
        self.curr_img = self.prev_img = None
        self.fb_err_mask = None
        self.gx, self.gy = grid_size_x, grid_size_y
        self.n = grid_size_x * grid_size_y
        self.psx, self.psy = float(patch_size_x), float(patch_size_y)
        self.reset, self.reinit = int(reset_at_each_frame), int(reset_at_each_frame) == 1
        self.dyn, self.inside = int(dyn_patch_size), bool(patch_centroid_inside)
        self.resx, self.resy = (self.gx + 1, self.gy + 1) if (self.dyn or self.inside) else (self.gx, self.gy)
        self.ssm = ssm
        self.est_params = est_params if est_params is not None else make_est_params()
        self.seed, self.frame = seed, 0
        self.trackers = [OracleTracker(cell_params) for _ in range(self.n)]
        self.prev_pts = np.zeros((self.n, 2), dtype=np.float32)
        self.curr_pts = np.zeros((self.n, 2), dtype=np.float32)
        self.last = None

    @staticmethod
    def _centroid(c):
        c = np.asarray(c, dtype=np.float64).reshape(2, 4)
        return np.array([np.float32((c[0, 0] + c[0, 1] + c[0, 2] + c[0, 3]) / 4.0), np.float32((c[1, 0] + c[1, 1] + c[1, 2] + c[1, 3]) / 4.0)],
                        dtype=np.float32)

    def _set_corners(self, corners):
        self.corners = np.array(corners, dtype=np.float64).reshape(2, 4)
        bp, bc = norm_unit_square_pts(self.resx, self.resy)
        H = homography_dlt(bc, self.corners)
        q = H @ np.vstack([bp.T, np.ones(bp.shape[0])])
        self.pts = q[:2] / q[2]

    def _reset(self, reinit):
        for t in range(self.n):
            r, c = divmod(t, self.gx)
            pc = None
            if self.resx == self.gx + 1:
                w = self.gx + 1
                pc = self.pts[:, [r * w + c, r * w + c + 1, (r + 1) * w + c + 1, (r + 1) * w + c]]
            if not self.dyn:
                cen = self.pts[:, t].copy()
                if self.inside:
                    cen = np.array([(pc[0, 0] + pc[0, 1] + pc[0, 2] + pc[0, 3]) / 4.0, (pc[1, 0] + pc[1, 1] + pc[1, 2] + pc[1, 3]) / 4.0])
                x0, y0 = cen[0] - self.psx / 2.0, cen[1] - self.psy / 2.0
                pc = np.array([[x0, x0 + self.psx, x0 + self.psx, x0], [y0, y0, y0 + self.psy, y0 + self.psy]])
            if reinit:
                self.trackers[t].initialize(pc)
            else:
                self.trackers[t].set_region(pc)
            self.prev_pts[t] = self._centroid(self.trackers[t].corners())

    def set_image(self, img):
        self.curr_img = img
        for t in self.trackers:
            t.set_image(img)

    def initialize(self, corners):
        self._set_corners(corners)
        self._reset(True)
        if self.enable_fb:
            self.prev_img = np.array(self.curr_img, copy=True)                 # GridTracker.cc:241-243

    def _backward_estimation(self, q):
        """GridTracker::backwardEstimation (GridTracker.cc:292-343)"""
        fb_prev = np.zeros((self.n, 2), dtype=np.float32)
        for t in range(self.n):
            tr = self.trackers[t]
            loc = tr.corners().copy()
            if self.fb_reinit:
                tr.initialize(loc)
            tr.set_image(self.prev_img)
            tr.update()
            fb_prev[t] = self._centroid(tr.corners())
            tr.set_image(self.curr_img)
            tr.set_region(loc)
        # cv::Point2f coordinates: the difference is a float operation, its square a double one (GridTracker.cc:310-313)
        d = (fb_prev - self.prev_pts).astype(np.float64)
        mask = ~((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) > self.fb_err_thresh)
        if int(mask.sum()) < q.n_model_pts:
            for t in range(self.n):
                if mask[t]:
                    continue
                mask[t] = True
                if int(mask.sum()) == q.n_model_pts:
                    break
        # (the reference appends the late additions at the END of the masked lists: same order here)
        first = ~((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) > self.fb_err_thresh)
        order = list(np.flatnonzero(first)) + [t for t in range(self.n) if mask[t] and not first[t]]
        est = estimate_warp(self.ssm, self.prev_pts[order], self.curr_pts[order], q)
        pix_mask = np.zeros(self.n, dtype=np.uint8)
        # pix_mask[tracker] = fb_err_mask ? pix_mask_est[est_pt_id++] : 0, walking the trackers in order (GridTracker.cc:336-342)
        k = 0
        for t in range(self.n):
            if mask[t]:
                pix_mask[t] = est["mask"][k]; k += 1
        self.fb_err_mask = mask
        est = dict(est); est["mask_est"] = est["mask"]; est["mask"] = pix_mask; est["order"] = np.array(order)
        return est

    def update(self):
        for t in range(self.n):
            self.trackers[t].update()
            self.curr_pts[t] = self._centroid(self.trackers[t].corners())
        self.frame += 1
        q = OrcEstParams()
        for name, _ in OrcEstParams._fields_:
            setattr(q, name, getattr(self.est_params, name))
        q.seed = self.seed + self.frame
        if self.enable_fb:
            est = self._backward_estimation(q)
            self.prev_img = np.array(self.curr_img, copy=True)
        else:
            est = estimate_warp(self.ssm, self.prev_pts, self.curr_pts, q)
        self.last = est
        s = est["state_update"]
        if self.ssm == "homography":
            W = np.array([[1 + s[0], s[1], s[2]], [s[3], 1 + s[4], s[5]], [s[6], s[7], 1.0]])
        else:
            W = np.array([[1 + s[2], s[3], s[0]], [s[4], 1 + s[5], s[1]], [0, 0, 1.0]])
        qh = W @ np.vstack([self.corners, np.ones(4)])
        self._set_corners(qh[:2] / qh[2])
        if self.reset:
            self._reset(self.reinit)
        else:
            self.prev_pts[:] = self.curr_pts
        return self.corners.copy()
