/*
 * mtf_oracle.h -- C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a dependency-free fp64 restatement of the
 * reference's (abhineet123/MTF) Lucas-Kanade hot path, used exclusively as the
 * checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Nothing under mtf_b200/ may include, link or call it.
 *
 * PARITY UNPINNED: the reference ships no tests / golden vectors for this path
 * and cannot be compiled in this image (no Eigen, OpenCV C++ or Boost headers),
 * so the oracle is pinned only by (i) line-by-line fidelity to the reference
 * sources cited at every function and (ii) an independent NumPy restatement
 * (tests/np_ref.py) plus the reference's own Diagnostics identities.
 *
 * Layouts follow the reference's Eigen (column-major) types:
 *   pts      2 x N   -> x0,y0,x1,y1,...
 *   corners  2 x 4   -> passed ROW-major here like the cv::Mat the reference
 *                        takes: x_UL,x_UR,x_LR,x_LL,y_UL,y_UR,y_LR,y_LL
 *   dI_dx    N x 2   -> all Ix then all Iy
 *   dI_dp    N x S   -> S contiguous N-columns
 *   H        S x S   -> column-major (symmetric up to rounding)
 */
#ifndef MTF_ORACLE_H
#define MTF_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_AM_SSD = 0, ORC_AM_NCC = 1, ORC_AM_MI = 2 };
enum { ORC_SSM_HOMOGRAPHY = 0, ORC_SSM_AFFINE = 1, ORC_SSM_TRANSLATION = 2 };
enum { ORC_SM_ESM = 0, ORC_SM_FCLK = 1, ORC_SM_ICLK = 2, ORC_SM_FALK = 4, ORC_SM_IALK = 5 };
/* Hessian type codes: the reference has one enum per SM; values as in
 * SM/include/mtf/SM/{ESM,FCLK,ICLK}Params.h */
enum { ORC_ESM_HESS_INITIAL_SELF = 0, ORC_ESM_HESS_CURRENT_SELF = 1, ORC_ESM_HESS_SUM_OF_SELF = 2,
       ORC_ESM_HESS_ORIGINAL = 3, ORC_ESM_HESS_SUM_OF_STD = 4, ORC_ESM_HESS_STD = 5 };
enum { ORC_ESM_JAC_ORIGINAL = 0, ORC_ESM_JAC_DIFF_OF_JACS = 1 };
enum { ORC_LK_HESS_INITIAL_SELF = 0, ORC_LK_HESS_CURRENT_SELF = 1, ORC_LK_HESS_STD = 2 };

typedef struct orc_params {
	int am, ssm, sm;
	int resx, resy;
	int max_iters;
	double epsilon;
	int hess_type, jac_type;
	int chained_warp;
	int leven_marq;
	double lm_delta_init, lm_delta_update;
	int nt_semantics;        /* 1: nt:: search methods, 0: templated twins */
	double grad_eps;
	int hom_normalized_init;
	int mi_n_bins;
	double mi_pre_seed;
	int mi_pou;
	double likelihood_alpha;
	int grad_mode;           /* 0: reference finite difference (imgUtils.cc:233-254); 1: its eps -> 0 limit,
	                            evaluated analytically (not a reference mode; see mtf_oracle.cpp) */
	int fast_sums;           /* 1: vectorised dot products in the Jacobian / Hessian products (CPU-baseline timing: Eigen's
	                            product kernels are vectorised too); 2: additionally the Hessian J^T J as a cache-blocked SYRK on
	                            the upper triangle (less work than Eigen's full product: an upper bound on the reference's
	                            speed); 0: sequential sums (parity tests) */
} orc_params;

typedef struct orc_tracker orc_tracker;

/* per-iteration record kept by update() for stage parity */
typedef struct orc_iter_log {
	double f;
	double jacobian[8];
	double hessian[64];
	double state_update[8];
	double corners[8];
	double update_norm;
	int rejected;            /* LM rejected the previous step in this pass */
} orc_iter_log;

void orc_default_params(orc_params *p);
orc_tracker *orc_create(const orc_params *p);
void orc_destroy(orc_tracker *t);
/* img: h x w row-major float, contiguous; pointer is kept, not copied
 * (as ImageBase::setCurrImg does) */
void orc_set_image(orc_tracker *t, const float *img, int h, int w);
int orc_initialize(orc_tracker *t, const double *corners);
int orc_update(orc_tracker *t);
int orc_set_region(orc_tracker *t, const double *corners);
int orc_n_iters(const orc_tracker *t);        /* loop passes executed in last update() */
int orc_n_log(const orc_tracker *t);          /* entries in the iteration log */
const orc_iter_log *orc_log(const orc_tracker *t, int i);
void orc_get_corners(const orc_tracker *t, double *out8);
void orc_get_state(const orc_tracker *t, double *outS);
int orc_state_size(const orc_tracker *t);
/* ssm.setState: warp the template points with the given state (stage tests at an arbitrary state) */
void orc_set_state(orc_tracker *t, const double *state);
void orc_get_pts(const orc_tracker *t, double *out2N);
void orc_get_init_pts(const orc_tracker *t, double *out2N);
void orc_get_init_pix_vals(const orc_tracker *t, double *outN);
void orc_get_curr_pix_vals(const orc_tracker *t, double *outN);
void orc_get_curr_pix_grad(const orc_tracker *t, double *outN2);
void orc_get_curr_pix_jacobian(const orc_tracker *t, double *outNS);
void orc_get_init_pix_jacobian(const orc_tracker *t, double *outNS);
double orc_get_similarity(const orc_tracker *t);
void orc_get_init_warp(const orc_tracker *t, double *out9); /* DLT warp, row-major */
/* PF inner loop: setState -> updatePixVals -> updateSimilarity(false) -> getLikelihood */
void orc_pf_evaluate(orc_tracker *t, const double *states, int n_particles,
	double *likelihood, double *similarity);
/* per-stage wall clock of the last update(), seconds, reference record_event labels order:
 * 0 updatePixVals 1 updateSimilarity 2 grads 3 updatePixGrad 4 pix jacobian 5 jacobian
 * 6 hessian 7 solve 8 compositionalUpdate */
void orc_get_stage_times(const orc_tracker *t, double *out9);

/* leaf utilities, exposed for stage tests */
double orc_pix_val(const float *img, int h, int w, double x, double y);
void orc_get_pix_vals(const float *img, int h, int w, const double *pts, int n,
	double norm_mult, double norm_add, double *out);
void orc_get_img_grad(const float *img, int h, int w, const double *pts, int n,
	double grad_eps, double pix_mult, double *outN2);
void orc_get_img_grad_analytic(const float *img, int h, int w, const double *pts, int n,
	double grad_eps, double pix_mult, double *outN2);
void orc_homography_dlt(const double *in_corners, const double *out_corners, double *H9);
/* utils::computeAffineNDLT (warpUtils.cc:378-386), row-major 3 x 3 */
void orc_affine_ndlt(const double *in_corners, const double *out_corners, double *H9);
void orc_colpiv_qr_solve(const double *A, const double *b, int n, double *x);
int orc_colpiv_qr(const double *A, int rows, int cols, double *qr_out, int *perm_out, double *hcoeffs_out);
void orc_norm_unit_square_pts(int resx, int resy, double min_x, double min_y,
	double max_x, double max_y, double *pts2N, double *corners8);

/* pre-processing (PreProcBase::processFrame with GaussianSmoothing, CV_32FC1 output): uint8 gray (channels 1) or BGR
 * (channels 3) frame, stride in bytes -> h x w float.  orc_gaussian_kernel5 = cv::getGaussianKernel(5, sigma, CV_32F) */
void orc_gaussian_kernel5(double sigma, float *k5);
void orc_preproc_gauss5(const unsigned char *img, int h, int w, int stride, int channels, double sigma, float *out);

/* batch drivers used as the CPU baseline: P independent trackers over the same image,
 * OpenMP over patches as GridTracker.cc:253-256 does.  Returns total LK iterations. */
long orc_batch_track(const orc_params *p, const float *const *frames, int n_frames, int h, int w,
	const double *corners /* P x 8 */, int n_patches, int n_threads,
	double *final_corners /* P x 8 */, int *iters_per_patch, double *seconds);

/* the same with GridTracker's per-frame reset of every cell at its initial corners (grid_reset_at_each_frame = 1);
 * final_corners = the corners after the last frame's update, before its reset */
long orc_batch_track_reset(const orc_params *p, const float *const *frames, int n_frames, int h, int w,
	const double *corners /* P x 8 */, int n_patches, int n_threads,
	double *final_corners /* P x 8 */, int *iters_per_patch, double *seconds);
/* PF particle evaluation of n_objects templates (initialised on frame0) x n_particles states on frame1; returns the
 * number of particle evaluations */
long orc_batch_pf_evaluate(const orc_params *p, const float *frame0, const float *frame1, int h, int w,
	const double *corners /* n_objects x 8 */, int n_objects, const double *states /* n_objects x n_particles x S */,
	int n_particles, int n_threads, double *likelihood /* n_objects x n_particles */, double *seconds);

/* nt::PF (SM/src/NT/PF.cc) with one sampler distribution; the caller supplies the random deviates:
 * normals max_iters x n_particles x R (R = orc_pf_n_normals: S, or 10 with corner based sampling), uniforms max_iters x n_particles */
typedef struct orc_pf_params {
	int n_particles, max_iters;
	double epsilon;
	int dynamic_model, update_type, likelihood_func, resampling_type, mean_type, reset_to_mean;
	double adaptive_resampling_thresh, measurement_sigma, ar_coeff;
	double ssm_sigma[8], ssm_mean[8];
	int corner_based_sampling;
} orc_pf_params;
typedef struct orc_pf orc_pf;
orc_pf *orc_pf_create(const orc_params *tp, const orc_pf_params *pp);
void orc_pf_destroy(orc_pf *f);
void orc_pf_set_image(orc_pf *f, const float *img, int h, int w);
int orc_pf_initialize(orc_pf *f, const double *corners);
int orc_pf_update(orc_pf *f, const double *normals, const double *uniforms);
void orc_pf_get_corners(const orc_pf *f, double *out8);
void orc_pf_get_state(const orc_pf *f, double *outS);
int orc_pf_n_normals(const orc_pf *f);
/* particle_states[curr_set_id], particle_wts, particle_cum_wts (any may be NULL); returns max_wt_id */
int orc_pf_get_particles(const orc_pf *f, double *states, double *wts, double *cum, int *resampled);

/* robust warp estimation from point pairs (mtf_oracle_est.cpp): SSM::estimateWarpFromPts of Homography / Affine, i.e.
 * estimateHomography / estimateAffine (SSM/src/{SSMEstimator,HomographyEstimator,AffineEstimator}.cc).  method follows
 * SSMEstimatorParams::EstType (SSMEstimatorParams.h:11); seed feeds cvRNG (the reference seeds from random_device). */
enum { ORC_EST_RANSAC = 0, ORC_EST_LMEDS = 1, ORC_EST_LEAST_SQUARES = 2 };
typedef struct orc_est_params {
	int method;
	double ransac_reproj_thresh;
	int n_model_pts, refine, max_iters, max_subset_attempts;
	double confidence;
	int lm_max_iters;
	unsigned long long seed;
} orc_est_params;
void orc_est_default_params(orc_est_params *p);
int orc_estimate_warp(int ssm, const float *in_pts, const float *out_pts, int n, const orc_est_params *ep,
	double *warp9, unsigned char *mask, double *state_update, int *info4);
int orc_est_subsets(const float *in_pts, const float *out_pts, int n, int model_points, int max_attempts, unsigned long long seed,
	int n_subsets, int *idx);
unsigned orc_cv_rand_int(unsigned long long *state);
void orc_sym_eigen(const double *A, int n, double *V_rows, double *w_desc);

#ifdef __cplusplus
}
#endif
#endif
