/*
 * mtf_b200.h -- C ABI of the B200-native Lucas-Kanade hot path (libmtf_b200.so).
 *
 * Drop-in boundary for abhineet123/MTF's per-iteration LK path.  The reference has no FFI for this
 * path (its plug-in surface is three C++ abstract classes, include/mtf/TrackerBase.h:9-70,
 * AM/include/mtf/AM/AppearanceModel.h:63-396, SSM/include/mtf/SSM/StateSpaceModel.h:49-408), so the
 * entry points below are what a `TrackerBase` subclass added to mtf::getTracker
 * (include/mtf/mtf.h:929) binds -- see include/mtf_b200_tracker.h for that subclass and
 * INTEGRATION.md for the factory branch.  Each function names the reference member it replaces.
 *
 * Plain C, POD only: no Eigen / OpenCV / torch types.  Every function returns an mtfb_status;
 * mtfb_last_error() gives the message for the calling thread.  A context is single-threaded and
 * non-re-entrant like a reference tracker instance; distinct contexts may be used concurrently.
 *
 * One context = one BATCH of P independent patch trackers (the GridTracker.cc:247-264 /
 * PF.cc:198-289 fan-out) sharing one image, one (SM, AM, SSM) combination and one parameter set.
 *
 * Layouts (all row-major C arrays, fp64 unless stated):
 *   corners   P x 2 x 4   x_UL,x_UR,x_LR,x_LL,y_UL,y_UR,y_LR,y_LL per patch -- the 2x4 CV_64FC1
 *                          cv::Mat of TrackerBase::initialize (SM/include/mtf/SM/SearchMethod.h:19)
 *   state     P x S        S = 8 (Homography.cc:94-107 order), 6 (Affine.cc:117-131 order) or 2 (Translation: tx, ty)
 *   pts       P x N x 2    x0,y0,x1,y1,... = Eigen column-major PtsT (2 x N)
 *   pix_vals  P x N
 *   pix_grad  P x 2 x N    all Ix then all Iy = column-major PixGradT (N x 2)
 *   pix_jac   P x S x N    S contiguous N-columns = column-major MatrixXd (N x S)
 *   hessian   S x S        column-major
 */
#ifndef MTF_B200_H
#define MTF_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum mtfb_status {
	MTFB_OK = 0,
	MTFB_ERR_INVALID_ARG = 1,     /* mtf::utils::InvalidArgument      excpUtils.h:38-45 */
	MTFB_ERR_NOT_SUPPORTED = 2,   /* mtf::utils::FunctonNotImplemented excpUtils.h:29-36 */
	MTFB_ERR_LOGIC = 3,           /* mtf::utils::LogicError           excpUtils.h:47-54 (call order) */
	MTFB_ERR_INVALID_STATE = 4,   /* mtf::utils::InvalidTrackerState  excpUtils.h:20-27 (NaN / singular H in >= 1 patch) */
	MTFB_ERR_CUDA = 5,            /* CUDA runtime error, no device, wrong architecture */
	MTFB_ERR_NO_MEMORY = 6
} mtfb_status;

/* AM/AM.cmake:5 names; SSM/SSM.cmake:1 names; SM/SM.cmake names */
enum { MTFB_AM_SSD = 0, MTFB_AM_NCC = 1, MTFB_AM_MI = 2 };
enum { MTFB_SSM_HOMOGRAPHY = 0, MTFB_SSM_AFFINE = 1,
       MTFB_SSM_TRANSLATION = 2 };   /* SSM/src/Translation.cc (GridTracker's default cell model): SSD, F64, ESM / FCLK / ICLK / FALK / IALK */
enum { MTFB_SM_ESM = 0, MTFB_SM_FCLK = 1, MTFB_SM_ICLK = 2, MTFB_SM_PF = 3,
       MTFB_SM_FALK = 4, MTFB_SM_IALK = 5 };   /* the additive searches nt::FALK / nt::IALK (SM/src/NT/FALK.cc, IALK.cc): SSD, F64 */
/* ESMParams::HessType / JacType (SM/include/mtf/SM/ESMParams.h:13-17) */
enum { MTFB_ESM_HESS_INITIAL_SELF = 0, MTFB_ESM_HESS_CURRENT_SELF = 1, MTFB_ESM_HESS_SUM_OF_SELF = 2,
       MTFB_ESM_HESS_ORIGINAL = 3, MTFB_ESM_HESS_SUM_OF_STD = 4, MTFB_ESM_HESS_STD = 5 };
enum { MTFB_ESM_JAC_ORIGINAL = 0, MTFB_ESM_JAC_DIFF_OF_JACS = 1 };
/* FCLKParams::HessType / ICLKParams::HessType / FALKParams / IALKParams::HessType (FCLKParams.h:8, ICLKParams.h:9) */
enum { MTFB_LK_HESS_INITIAL_SELF = 0, MTFB_LK_HESS_CURRENT_SELF = 1, MTFB_LK_HESS_STD = 2 };
/* arithmetic of the per-pixel part of the update kernel (mtfb_params::precision).
 *   F64: every operation in fp64, in the reference's order: warped points, sampling indices, pixel values bit-identical
 *        to the reference's Eigen path; Jacobian / Hessian to summation order.
 *   F32: fp32 per-pixel arithmetic in patch-local coordinates with BIT-EXACT SAMPLING INDICES (pixels whose fp32
 *        coordinate is within the fp32 error bound of a cell boundary are re-evaluated in fp64), fp64 reduction and
 *        solve; Jacobian / Hessian / corners agree with F64 to fp32 tolerance.  SSD: ESM / FCLK /
 *        ICLK with the chained warp, and PF particle evaluation.  NCC: ESM (DiffOfJacs Jacobian) / FCLK with the Initial /
 *        Current / SumOf Self Hessians, chained warp -- one fp32 sweep per pass instead of the reference's three
 *        (AM/src/NCC.cc:124-280, 337-389; DESIGN.md section 4 "NCC"). */
enum { MTFB_PRECISION_F64 = 0, MTFB_PRECISION_F32 = 1 };
/* how the F32 precision solves H dp = -J^T when the Hessian is the pass's own (FCLK / ESM CurrentSelf, no LM)
 * (mtfb_params::f32_solve; F64 contexts always run the reference's solve):
 *   REFERENCE: H and J^T are mapped from the kernel's centred accumulation basis to the reference's parameters in fp64 and
 *        solved by the same column-pivoted Householder QR with Eigen's rank rule (`colPivHouseholderQr().solve`,
 *        SM/src/NT/FCLK.cc:298, NT/ESM.cc:266) -- including the truncated steps that rule takes when raw pixel coordinates
 *        (hom_normalized_init = 0) make H numerically rank deficient.
 *   LOCAL: an 8 x 8 Gauss-Jordan solve in the centred basis (condition ~1e3 instead of ~1e16), mapped back: the full
 *        Gauss-Newton step -H^-1 J^T in exact arithmetic, without the rank decisions.  Faster; converges to the same fixed
 *        point; per-pass updates differ from the reference's wherever its rank rule fires. */
enum { MTFB_F32_SOLVE_REFERENCE = 0, MTFB_F32_SOLVE_LOCAL = 1 };
/* per-patch status bits reported by mtfb_get_patch_status */
enum { MTFB_PATCH_OK = 0, MTFB_PATCH_NAN = 1, MTFB_PATCH_SINGULAR = 2, MTFB_PATCH_OUT_OF_IMAGE = 4 };

/*
 * Parameters = the fields of {ESM,FCLK,ICLK,PF}Params + the AM/SSM parameters the path reads
 * (SM/src/{ESM,FCLK,ICLK}Params.cc:4-17, AM/include/mtf/AM/MI.h MIParams, AM/include/mtf/AM/ImageBase.h:7-9,
 * SSM/include/mtf/SSM/Homography.h HomographyParams).  mtfb_default_params() fills the shipped
 * configuration (Config/mtf.cfg:14,24 + Config/modules.cfg).
 */
typedef struct mtfb_params {
	int am, ssm, sm;
	int resx, resy;              /* sampling resolution: N = resx * resy                          */
	int n_patches;               /* P: independent trackers in the batch                          */
	int max_iters;
	double epsilon;              /* stop when || prev_corners - curr_corners ||^2 < epsilon       */
	int hess_type, jac_type;
	int chained_warp;            /* {esm,fc,ic}_chained_warp (nt:: search methods only; 0 = the
	                                getWarpedImgGrad path, evaluated literally)                    */
	int leven_marq;
	double lm_delta_init, lm_delta_update;
	int nt_semantics;            /* 1: nt::SM control flow (SM/src/NT), 0: templated twins (SM/src) */
	double grad_eps;             /* ImageBase.h:7-9; used for the out-of-image edge emulation only:
	                                the product computes the eps -> 0 limit of the reference's finite
	                                difference analytically (DESIGN.md "gradient semantics")      */
	int hom_normalized_init;     /* the SSM's normalized_init: hom_normalized_init (0: factory default, parameters.h:261;
	                                1: Config/modules.cfg) for the Homography; aff_normalized_init for the Affine SSM
	                                (Affine.cc:65-74: start from utils::computeAffineNDLT; F64 precision)              */
	int mi_n_bins;               /* MIParams n_bins (parameters.h:344)                            */
	double mi_pre_seed;
	int mi_pou;
	double likelihood_alpha;     /* AMParams::likelihood_alpha, PF path                           */
	int device;                  /* CUDA device ordinal                                           */
	int threads_per_patch;       /* 0 = chosen from n_patches; otherwise 32, 64, 128 or 256        */
	int occupancy;               /* register budget of the update kernel: 0 / 1 / 2 = about 8 / 12 /
	                                16 resident warps per SM (tuning knob, results do not change)  */
	int precision;               /* MTFB_PRECISION_F64 (default) or MTFB_PRECISION_F32                */
	int f32_solve;               /* MTFB_F32_SOLVE_REFERENCE (default) or MTFB_F32_SOLVE_LOCAL; F32 only */
} mtfb_params;

/* one Gauss-Newton pass as the reference's record_event() trail would show it (NT/FCLK.cc:190-321);
 * written by mtfb_update() when an iteration log is attached */
typedef struct mtfb_iter_log {
	double f;                    /* am->getSimilarity() after updateSimilarity                    */
	double jacobian[8];          /* df_dp                                                         */
	double hessian[64];          /* d2f_dp2, column-major S x S in the first S*S entries          */
	double state_update[8];      /* ssm_update                                                    */
	double corners[8];           /* ssm->getCorners() after the update                            */
	double update_norm;          /* || prev_corners - curr_corners ||^2                           */
	int rejected;                /* Levenberg-Marquardt rejected the previous step in this pass   */
	int valid;                   /* 1 if this pass was executed                                   */
} mtfb_iter_log;

typedef struct mtfb_ctx mtfb_ctx;

const char *mtfb_last_error(void);
const char *mtfb_version(void);
/* number of kernels this library has launched on the calling context so far (bench.py gpu_launches) */
long mtfb_launch_count(const mtfb_ctx *ctx);

void mtfb_default_params(mtfb_params *p);
/* replaces: new nt::FCLK(AM(getAM(..)), SSM(getSSM(..)), params)  include/mtf/mtf.h:1282-1300 */
mtfb_status mtfb_create(const mtfb_params *p, mtfb_ctx **out);
mtfb_status mtfb_destroy(mtfb_ctx *ctx);
/* run all work of this context on a caller-owned cudaStream_t, taken literally (NULL = CUDA's legacy default
 * stream).  A new context runs on a private non-blocking stream of its own. */
mtfb_status mtfb_set_stream(mtfb_ctx *ctx, void *cuda_stream);
mtfb_status mtfb_synchronize(mtfb_ctx *ctx);

/* replaces TrackerBase::setImage -> ImageBase::setCurrImg (AM/src/ImageBase.cc:38-60).  The reference keeps
 * a pointer to a buffer the application overwrites in place every frame (TrackerBase.h:21-26), so this
 * must be called before every initialize()/update(): it uploads the h x w float frame (row_stride in
 * elements) to the device asynchronously on the context stream. */
mtfb_status mtfb_set_image(mtfb_ctx *ctx, const float *host_img, int h, int w, int row_stride);
/* same, frame already resident on the device (pitch in elements); the pointer is used, not copied: it must belong to the
 * context's device, and the buffer must stay valid and unmodified until the work queued on it has finished (mtfb_synchronize,
 * or any getter); writes to it must be ordered after that on the context stream (mtfb_set_stream) */
mtfb_status mtfb_set_image_device(mtfb_ctx *ctx, const float *dev_img, int h, int w, int pitch);

/* replaces the pre-processing in front of setImage: utils::GaussianSmoothing (PreProcBase::processFrame for CV_32FC1,
 * Utilities/src/preprocUtils.cc:108-127; pre_proc_type "gauss", gauss_kernel_size 5, gauss_sigma_x 3 of
 * Config/parameters.h:229-235) applied to the RAW frame on the device: uint8, channels = 1 (gray) or 3 (BGR, converted with
 * cv::cvtColor's float weights), row_stride in BYTES.  Uploads h x w x channels bytes instead of 4 h w, then behaves like
 * mtfb_set_image with the smoothed float frame.  kernel_size must be 5; sigma > 0 (sigma_y = sigma_x as MTF passes it). */
mtfb_status mtfb_set_image_u8(mtfb_ctx *ctx, const unsigned char *host_img, int h, int w, int row_stride, int channels,
	int kernel_size, double sigma);
/* PREFETCH variants: the upload (and, for raw frames, the pre-processing) runs on a private copy stream into one of two
 * device buffers while the context stream is still tracking the previous frame; the next mtfb_initialize / mtfb_set_region /
 * mtfb_update waits for it (event, no host synchronisation) and samples it.  One frame may be in flight.  The host buffer must
 * be pinned for the copy to overlap and must stay untouched until that next call has been made and the context synchronised
 * (mtfb_synchronize, or any getter).  The reference has no counterpart: it re-reads the caller's buffer in place on every
 * update() (TrackerBase.h:21-26) -- an application that wants the overlap hands frame t+1 over right after update(t):
 *     mtfb_set_image_async(ctx, frame[0]); mtfb_initialize(ctx, corners);
 *     for t: mtfb_set_image_async(ctx, frame[t]); mtfb_update(ctx); (read the results of t - 1 meanwhile) */
mtfb_status mtfb_set_image_async(mtfb_ctx *ctx, const float *host_img, int h, int w, int row_stride);
mtfb_status mtfb_set_image_u8_async(mtfb_ctx *ctx, const unsigned char *host_img, int h, int w, int row_stride, int channels,
	int kernel_size, double sigma);
/* the frame the trackers currently sample (after mtfb_set_image / mtfb_set_image_u8), h x w floats, contiguous; host pointer */
mtfb_status mtfb_get_image(mtfb_ctx *ctx, float *out);

/* replaces SearchMethod::initialize(corners) (SM/src/NT/FCLK.cc:102-169, NT/ESM.cc:110-146,
 * NT/ICLK.cc:71-127): ssm.setCorners (4-point DLT, Utilities/src/warpUtils.cc:171-223) + am.initializePixVals
 * + the SM-specific template gradients / Jacobians / Hessians, for all P patches, on the device. */
mtfb_status mtfb_initialize(mtfb_ctx *ctx, const double *corners /* P x 8, host */);
/* replaces SearchMethod::setRegion (NT/FCLK.cc:360-376, NT/ESM.cc:148-167, NT/ICLK.cc:132-160 with update_ssm = 0): new
 * corners, identity warp; ESM and FCLK-InitialSelf also rebuild init_pix_jacobian / init_self_hessian (SSD only) */
mtfb_status mtfb_set_region(mtfb_ctx *ctx, const double *corners /* P x 8, host */);
/* replaces SearchMethod::update() (NT/FCLK.cc:171-358, NT/ESM.cc:170-297, NT/ICLK.cc:160-299): the whole
 * <= max_iters loop for every patch in ONE launch.  Asynchronous; results are read with the getters. */
mtfb_status mtfb_update(mtfb_ctx *ctx);
/* a single Gauss-Newton pass for every patch (max_iters = 1 for this call): stage parity and hosts that
 * want to interleave their own logic between iterations.  Outputs may be NULL; host pointers. */
mtfb_status mtfb_iterate_once(mtfb_ctx *ctx, double *jacobian /* P x S */, double *hessian /* P x S x S */,
	double *similarity /* P */, double *state_update /* P x S */);
/* attach (n_slots > 0) or detach (0) a device-side iteration log of n_slots passes per patch */
mtfb_status mtfb_enable_iter_log(mtfb_ctx *ctx, int n_slots);
mtfb_status mtfb_get_iter_log(mtfb_ctx *ctx, mtfb_iter_log *out /* P x n_slots, host */);

/* replaces the PF particle loop NT/PF.cc:303-320: for object o (one initialised patch = one template) and
 * each of its n_particles states: ssm.setState -> am.updatePixVals -> am.updateSimilarity(false) ->
 * am.getLikelihood() (SSD.h:41-43, NCC.cc:50-53, MI.cc:384-387).  states: P x n_particles x S; outputs P x n_particles
 * (NULL allowed). Host pointers. */
mtfb_status mtfb_pf_evaluate(mtfb_ctx *ctx, const double *states, int n_particles,
	double *likelihood, double *similarity);
/* device-pointer variant (no copies): all three buffers live on ctx's device */
mtfb_status mtfb_pf_evaluate_device(mtfb_ctx *ctx, const double *d_states, int n_particles,
	double *d_likelihood, double *d_similarity);

/* ---- the particle filter search method around mtfb_pf_evaluate: nt::PF (SM/src/NT/PF.cc:15-600, SM/include/mtf/SM/PFParams.h) ----
 * A context created with sm = MTFB_SM_PF becomes a PF tracker once mtfb_pf_configure() has been called (before
 * mtfb_initialize): mtfb_initialize() then also runs PF::initializeParticles (NT/PF.cc:185-198), mtfb_set_region() is
 * PF::setRegion (NT/PF.cc:596-600) and mtfb_update() is PF::update (NT/PF.cc:207-446) for every object of the batch: dynamic
 * model -> particle evaluation -> weights -> resampling -> mean state, all on the device (pf_tracker.cu).
 * Implemented: Homography (generatePerturbation with and without hom_corner_based_sampling, Homography.cc:899-915); ONE sampler
 * distribution (n_distr = 1); RandomWalk / AutoRegression1 x Additive / Compositional; likelihood AM / Gaussian / Reciprocal
 * (the last two for SSD and NCC); resampling None / BinaryMultinomial / LinearMultinomial, adaptive or not; mean None / SSM /
 * Corners; reset_to_mean.  Not implemented (MTFB_ERR_NOT_SUPPORTED): Residual resampling, several distributions with
 * update_distr_wts, pix_sigma, jacobian_as_sigma, enable_learning, the Affine samplers (Affine.cc:464-560). */
enum { MTFB_PF_RANDOM_WALK = 0, MTFB_PF_AUTO_REGRESSION1 = 1 };                                  /* PFParams::DynamicModel */
enum { MTFB_PF_UPDATE_ADDITIVE = 0, MTFB_PF_UPDATE_COMPOSITIONAL = 1 };                          /* PFParams::UpdateType */
enum { MTFB_PF_LIKELIHOOD_AM = 0, MTFB_PF_LIKELIHOOD_GAUSSIAN = 1, MTFB_PF_LIKELIHOOD_RECIPROCAL = 2 };  /* LikelihoodFunc */
enum { MTFB_PF_RESAMPLE_NONE = 0, MTFB_PF_RESAMPLE_BINARY_MULTINOMIAL = 1, MTFB_PF_RESAMPLE_LINEAR_MULTINOMIAL = 2,
       MTFB_PF_RESAMPLE_RESIDUAL = 3 };                                                          /* ResamplingType */
enum { MTFB_PF_MEAN_NONE = 0, MTFB_PF_MEAN_SSM = 1, MTFB_PF_MEAN_CORNERS = 2 };                  /* MeanType */
typedef struct mtfb_pf_params {
	int n_particles;
	int max_iters;               /* PF iterations per frame (pf_max_iters) */
	double epsilon;              /* stop iterating when || prev_corners - curr_corners ||^2 < epsilon */
	int dynamic_model, update_type, likelihood_func, resampling_type, mean_type;
	int reset_to_mean;
	double adaptive_resampling_thresh;   /* in (0, 1]: resample only when n_eff <= thresh * n_particles; else always */
	double measurement_sigma;
	double ar_coeff;             /* `a` of the AutoRegression1 models (ProjectiveBase.h:72-75 default 0.5) */
	double ssm_sigma[8], ssm_mean[8];    /* the sampler's normal distributions, one per state entry; with corner based sampling
	                                        entry 0 is the common translation's and entry 1 the corner offsets' (Homography.cc:901-906) */
	int corner_based_sampling;   /* HomographyParams::corner_based_sampling (parameters.h:262 default 1) */
	unsigned long long seed;     /* device generator (Philox4x32-10 + Box-Muller); the reference seeds from random_device */
	int object_offset;           /* index of this context's first object in a job sharded over several contexts / GPUs */
	int record_randoms;          /* keep the deviates of the last update for mtfb_pf_get_random_stream */
} mtfb_pf_params;
/* Config/modules.cfg:152-177 (500 particles, AutoRegression1, Compositional, AM likelihood, BinaryMultinomial with adaptive
 * threshold 0.2, mean None, one iteration); sigma / mean zero: the caller sets the sampler */
void mtfb_pf_default_params(mtfb_pf_params *p);
/* replaces: new nt::PF(am, ssm, params) + ssm->initializeSampler (NT/PF.cc:15-134, 152) */
mtfb_status mtfb_pf_configure(mtfb_ctx *ctx, const mtfb_pf_params *p);
/* random stream for the NEXT mtfb_update() from the host instead of the device generator (parity tests; replaying a recorded
 * run): normals max_iters x P x n_particles x R standard normal deviates (R = 8, or 10 with corner based sampling: tx, ty, then
 * (dx, dy) of the four corners), uniforms max_iters x P x n_particles in (0, 1).  Either may be NULL. */
mtfb_status mtfb_pf_set_random_stream(mtfb_ctx *ctx, const double *normals, const double *uniforms);
/* the deviates the last mtfb_update() used (record_randoms = 1), same layouts; uniforms of objects that did not resample are 0 */
mtfb_status mtfb_pf_get_random_stream(mtfb_ctx *ctx, double *normals, double *uniforms);
/* particle_states[curr_set_id], particle_wts, particle_cum_wts, max_wt_id after the last update; any output may be NULL */
mtfb_status mtfb_pf_get_particles(mtfb_ctx *ctx, double *states /* P x n x S */, double *weights /* P x n */,
	double *cum_weights /* P x n */, int *max_wt_id /* P */);

/* ---- robust warp estimation from point pairs: the step GridTracker::update runs after its cells
 * Replaces SSM::estimateWarpFromPts of Homography (SSM/src/Homography.cc:885-897 -> estimateHomography,
 * HomographyEstimator.cc:168-228) and Affine (Affine.cc:359-369 -> estimateAffine, AffineEstimator.cc:126-185) with the RANSAC /
 * LMedS drivers and the Levenberg-Marquardt refinement of SSMEstimator.cc:73-518, as ONE kernel launch.  Fields and defaults
 * = SSMEstimatorParams (SSMEstimatorParams.cc:5-13, .h:11-32) without use_boost_rng (the cvRNG stream only); `seed` replaces the
 * reference's boost::random_device seed (SSMEstimator.cc:22-24) so that a run can be reproduced: the subsets drawn are those of
 * cvRNG(seed) / cvRandInt.  Not implemented: Translation's estimator (Translation.cc:187-238), n_model_pts > 8. */
enum { MTFB_EST_RANSAC = 0, MTFB_EST_LMEDS = 1, MTFB_EST_LEAST_SQUARES = 2 };   /* SSMEstimatorParams::EstType */
typedef struct mtfb_est_params {
	int method;
	double ransac_reproj_thresh;   /* <= 0 -> 3 (SSMEstimatorParams.cc:55-57) */
	int n_model_pts, refine, max_iters, max_subset_attempts;
	double confidence;
	int lm_max_iters;
	unsigned long long seed;       /* 0 -> cvRNG's default state */
} mtfb_est_params;
void mtfb_est_default_params(mtfb_est_params *p);
/* in_pts, out_pts: n x 2 floats (std::vector<cv::Point2f>), host pointers.  Outputs (host, any may be NULL): state_update (8
 * doubles; 6 used for Affine), mask (n bytes: the inliers), warp (3 x 3 row-major; zeros when the estimation fails, as the
 * reference returns), info = { result, hypotheses drawn, inliers, LM evaluations }.  ssm: MTFB_SSM_HOMOGRAPHY or _AFFINE
 * (GridTracker's own SSM, independent of the cells' SSM). */
mtfb_status mtfb_estimate_warp_from_pts(mtfb_ctx *ctx, int ssm, const float *in_pts, const float *out_pts, int n,
	const mtfb_est_params *ep, double *state_update, unsigned char *mask, double *warp, int *info);
/* the same from two DEVICE arrays of n quadrilaterals (n x 8 doubles, x0..x3 y0..y3): the points are their centroids
 * (utils::getCentroid, miscUtils.h:473-480).  What a multi-GPU host runs on the all-gathered corners of every rank's cells. */
mtfb_status mtfb_estimate_warp_from_corners_device(mtfb_ctx *ctx, int ssm, const double *d_in_corners, const double *d_out_corners,
	int n, const mtfb_est_params *ep, double *state_update, unsigned char *mask, double *warp, int *info);
/* the same on the context's own cells without a host copy of the points: in_pts = the centroids (utils::getCentroid,
 * miscUtils.h:473-480) of the regions at the last mtfb_initialize / mtfb_set_region / mtfb_grid_commit (prev_pts,
 * GridTracker.cc:389), out_pts = the centroids of the current regions (curr_pts, :257).  mtfb_grid_enable first (it allocates the
 * point buffers; from then on initialize / set_region record prev_pts with one more small launch). */
mtfb_status mtfb_grid_enable(mtfb_ctx *ctx);
mtfb_status mtfb_grid_estimate(mtfb_ctx *ctx, int ssm, const mtfb_est_params *ep, double *state_update, unsigned char *mask,
	double *warp, int *info);
/* prev_pts = curr_pts: GridTracker::update with reset_at_each_frame = 0 (GridTracker.cc:276-279) */
mtfb_status mtfb_grid_commit(mtfb_ctx *ctx);
/* GridTracker::update after its cells, WITHOUT a host hop (SM/src/GridTracker.cc:265-280 + resetTrackers :345-392 for
 * reset_at_each_frame = 1, patch_centroid_inside = 1, dyn_patch_size = 0 -- the shipped grid): mtfb_grid_estimate, then on the
 * device region <- ssm.applyWarpToCorners(region, state_update), the (grid_size + 1)^2 grid of the region
 * (utils::getPtsFromCorners), every cell's patch_size box around the centroid of its four grid points, every cell
 * re-initialised there on the current frame, prev_pts <- the new centroids.  region: 2 x 4 corners of the grid's own SSM,
 * in / out (host).  One stream synchronisation, at the end. */
/* GridTracker::initialize / setRegion with reset_at_each_frame = 1 (GridTracker.cc:232-245, 282-287): the same layout for a given
 * region, every cell initialised there on the current frame (the region is not changed: with an identity update the layout
 * kernel's applyWarpToCorners is the identity up to the division by one) */
mtfb_status mtfb_grid_initialize(mtfb_ctx *ctx, int grid_size_x, int grid_size_y, double patch_size_x, double patch_size_y,
	const double *region /* 8 */);
mtfb_status mtfb_grid_advance(mtfb_ctx *ctx, int ssm, const mtfb_est_params *ep, int grid_size_x, int grid_size_y,
	double patch_size_x, double patch_size_y, double *region /* 8, in / out */, double *state_update, unsigned char *mask /* P */,
	double *warp /* 9 */, int *info /* 4 */);
/* prev_pts / curr_pts as the last mtfb_grid_estimate saw them (P x 2 floats each, host; either may be NULL) */
mtfb_status mtfb_grid_get_pts(mtfb_ctx *ctx, float *prev_pts, float *curr_pts);

/* getters = the accessors of SURVEY.md 8(a18): ssm->getCorners/getState/getPts, am->getSimilarity,
 * am->getInitPixVals/getCurrPixVals/getCurrPixGrad ...; they synchronise the context stream.
 * Host pointers. */
mtfb_status mtfb_get_corners(mtfb_ctx *ctx, double *out /* P x 8 */);
mtfb_status mtfb_get_state(mtfb_ctx *ctx, double *out /* P x S */);
mtfb_status mtfb_get_n_iters(mtfb_ctx *ctx, int *out /* P: loop passes of the last update() */);
mtfb_status mtfb_get_similarity(mtfb_ctx *ctx, double *out /* P */);
mtfb_status mtfb_get_patch_status(mtfb_ctx *ctx, int *out /* P */);
mtfb_status mtfb_get_init_warp(mtfb_ctx *ctx, double *out /* P x 9, DLT warp row-major */);
mtfb_status mtfb_get_init_pts(mtfb_ctx *ctx, double *out /* P x N x 2 */);
mtfb_status mtfb_get_init_pix_vals(mtfb_ctx *ctx, double *out /* P x N */);
/* debug taps: evaluate pts / It / dIt_dx / dIt_dp at the CURRENT state with the path's own device
 * functions (one extra launch); any output may be NULL */
mtfb_status mtfb_get_curr_stage(mtfb_ctx *ctx, double *pts /* P x N x 2 */, double *pix_vals /* P x N */,
	double *pix_grad /* P x 2 x N */, double *pix_jac /* P x S x N */);
/* precision = MTFB_PRECISION_F32 only: the same taps from the fp32 front end.  idx: the sampling indices (lx, ly)
 * = ((int)x, (int)y) of imgUtils.h:99-100, -1 where the point is outside the image; fast_err: |fp32 - fp64| warped
 * coordinate (px) for pixels that took the fp32 path, -1 for those re-evaluated in fp64.  Any output may be NULL. */
mtfb_status mtfb_get_curr_stage_f32(mtfb_ctx *ctx, int *idx /* P x N x 2 */, float *pix_vals /* P x N */,
	float *pix_grad /* P x 2 x N */, double *pix_jac /* P x S x N */, float *fast_err /* P x N */);
/* device pointers of the result arrays, valid until destroy (corners P x 8, state P x S, n_iters P int):
 * what a multi-GPU host all-gathers without a host round trip */
mtfb_status mtfb_device_results(mtfb_ctx *ctx, double **d_corners, double **d_state, int **d_n_iters);
int mtfb_state_size(const mtfb_ctx *ctx);

/* ---- Multi-GPU: the per-frame exchange of the batch's corners over NVLink peer memory (one process per GPU).
 * The reference's fan-out is a shared-memory loop over trackers[i]->update() whose results every later stage reads
 * (SM/src/GridTracker.cc:247-274: the cells' regions feed ssm.estimateWarpFromPts); split over GPUs, each rank tracks a
 * contiguous range of the job's n_total patches and needs all n_total x 8 corners after every frame.  No collective library:
 * every rank exports an array (CUDA IPC), maps the others', and the update kernels store each patch's final corners
 * straight into the arrays of all ranks; mtfb_peer_gather then publishes / awaits one sequence flag per rank.
 *   mtfb_peer_export   allocate this rank's gathered array (n_total x 8 fp64, double-buffered) and return its IPC handle;
 *                      the host exchanges the handles (any transport: torch.distributed, MPI, a file)
 *   mtfb_peer_attach   rank / world (<= 8) / this rank's first row / the handles of all ranks (its own entry is ignored)
 *   mtfb_peer_gather   after mtfb_update (or initialize / set_region: their corners are pushed by one small kernel):
 *                      stream-ordered; when it has run, the local gathered array is complete.  Every rank calls it once per
 *                      frame; mtfb_update refuses to start the next frame before.
 *   mtfb_peer_gathered device pointer of the current gathered array (valid until the next but one frame)
 *   mtfb_get_gathered_region  the same, copied to the host (synchronises; reports a rank that never signalled) */
#define MTFB_PEER_HANDLE_BYTES 64
typedef struct mtfb_peer_handle { unsigned char bytes[MTFB_PEER_HANDLE_BYTES]; } mtfb_peer_handle;
mtfb_status mtfb_peer_export(mtfb_ctx *ctx, int n_total, mtfb_peer_handle *out);
mtfb_status mtfb_peer_attach(mtfb_ctx *ctx, int rank, int world, int row0, const mtfb_peer_handle *handles /* world */);
mtfb_status mtfb_peer_gather(mtfb_ctx *ctx);
mtfb_status mtfb_peer_gathered(mtfb_ctx *ctx, const double **d_corners /* n_total x 8 */, int *n_total);
mtfb_status mtfb_get_gathered_region(mtfb_ctx *ctx, double *out /* n_total x 8 */);

/* test entry point (no reference counterpart as a function: the solve inside nt::FCLK / ESM / ICLK::update,
 * `H.colPivHouseholderQr().solve(J^T)`, SM/src/NT/FCLK.cc:298, NT/ESM.cc:266, NT/ICLK.cc:228): runs the device's warp-level
 * column-pivoted Householder QR on n_sys caller-supplied n x n systems (n = 6 or 8; A column-major, host pointers).
 * fast = 0: the literal restatement of Eigen 3.3's computeInPlace (norm down-dating); 1: the tuned per-pass variant;
 * 2: the variant of the F32 precision (square root / reciprocals from Newton-refined fp32 seeds, column-wise back-substitution).
 * Outputs: x (n_sys x n), nonzero_pivots (n_sys, may be NULL), perm (n_sys x n: original column at position i, may be NULL). */
mtfb_status mtfb_debug_colpiv_qr_solve(int device, int n, int fast, int n_sys, const double *A, const double *b, double *x,
	int *nonzero_pivots, int *perm);

#ifdef __cplusplus
}
#endif
#endif
