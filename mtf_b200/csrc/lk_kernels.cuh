// lk_kernels.cuh -- declarations shared by the kernel translation unit and the C-ABI host code.
#pragma once
#include <cuda_runtime.h>
#include "../../include/mtf_b200.h"
#include "lk_math.cuh"

namespace mtfb {

// Everything one launch needs, passed by value.  All pointers are device pointers; per-patch arrays are
// indexed [patch][...] (SoA per patch so that consecutive threads touch consecutive addresses).
// Multi-GPU exchange over NVLink peer memory (peer_gather.cu): where the ranks of the job keep their (n_total x 8) arrays of
// gathered corners -- this rank's own and, mapped through CUDA IPC, the others' -- and this rank's first row in them.
// The update kernels store a patch's final corners straight into all of them (n = 0: single GPU, nothing extra).
enum { MTFB_MAX_PEERS = 8 };
struct PeerOut { double *dst[MTFB_MAX_PEERS]; int n, row0; };

struct DevBatch {
	int P, N, resx, resy;
	Image img;
	const double *xv, *yv;       // normalised sampling grid (LinSpaced values), resx / resy entries
	float gx_lo, gx_step, gy_lo, gy_step;   // the same grid in fp32 as low + i * step (precision = MTFB_PRECISION_F32)
	const double *norm_corners;  // 8: corners of the normalised grid (x0..x3, y0..y3)
	double *dlt;                 // P x 9   DLT warp of setCorners, row-major
	double *warp;                // P x 9   curr_warp, row-major
	double *state;               // P x S   curr_state
	double *corners;             // P x 8   curr_corners
	double *init_corners;        // P x 8
	double *I0;                  // P x N   template pixel values (am.I0)
	double *G0;                  // P x 2 x N  template gradient: chained with the init warp, or (chained = 0) the
	                             //            warped-image gradient of initialize() time
	double *G0raw;               // P x 2 x N  am.getInitPixGrad(): the template gradient before any chaining (kept for
	                             //            setRegion of ESM / FCLK-InitialSelf, SSD), else null
	float *I0f, *G0f;            // fp32 copies of I0 / G0 (precision = MTFB_PRECISION_F32), else null
	int I0f_stride;              // elements per patch in I0f (N rounded up to 4: 16-byte aligned rows for the bulk copy)
	double *Hinit;               // P x 64  init_self_hessian, column-major S x S
	double *It_scratch;          // P x N   MI with templates too large for shared memory: the current pixel values of a
	                             //         pass, written by the histogram sweep and read back by the gradient sweep; NCC with
	                             //         ESM / FCLK: the pixel values of the last pass (what setRegion's Hessian reads); else null
	double *am_scal;             // P x 8   per-template scalars of the AM (NCC: I0_mean, c)
	double *ncc_tab;             // P x 64  NCC: template sums behind cmptInitHessian (sum D0 | sum D0 D0^T | sum I0cc D0)
	double *f;                   // P       similarity
	int *n_iters;                // P
	int *status;                 // P
	long long *n_iters_prof;     // experiment builds (MTFB_PROF): phase cycle counters, else null
	mtfb_iter_log *log;          // P x log_slots or null
	int log_slots;
	// parameters
	int max_iters, hess_type, jac_type, leven_marq, nt_semantics;
	int additive_sm;             // the search method is nt::FALK / nt::IALK: their template Jacobian is ssm.cmptPixJacobian at the start state
	int chained;                 // {esm,fc,ic}_chained_warp
	int norm_init;               // hom_normalized_init
	int f32_local_solve;         // mtfb_params::f32_solve == MTFB_F32_SOLVE_LOCAL
	double epsilon, lm_delta_init, lm_delta_update, grad_eps;
	double pix_mult, pix_add;    // am pix_norm_mult / pix_norm_add (1, 0 except MI)
	double grad_mult;            // pix_mult / (2 grad_eps)  (imgUtils.cc:238)
	PeerOut peers;               // gathered-corner arrays of all ranks (fused all-gather), n = 0 unless mtfb_peer_attach
};

struct StageTaps { double *pts, *pix_vals, *pix_grad, *pix_jac; };
// taps of the fp32 front end (lk_ssd_f32.cu): idx P x N x 2 (lx, ly), fast_err P x N (|fp32 - fp64| warped coordinate
// of the pixels that took the fp32 path, -1 for those that took the fp64 path)
struct StageTapsF32 { int *idx; float *pix_vals, *pix_grad, *fast_err; double *pix_jac; };

// launchers; threads = threads per patch (32 / 64 / 128 / 256), occ = register-budget knob of the SSD update kernel
// lk_ssd.cu
cudaError_t launch_init_ssd(int ssm, int threads, const DevBatch &b, const double *d_corners, cudaStream_t st);
cudaError_t launch_update_ssd(int ssm, int sm, int threads, int occ, const DevBatch &b, cudaStream_t st);
cudaError_t launch_set_region(int ssm, const DevBatch &b, const double *d_corners, cudaStream_t st);
// setRegion of the search methods that keep template Jacobians (NT/ESM.cc:150-168, NT/FCLK.cc:360-376): SSD
cudaError_t launch_reinit_ssd(int ssm, int threads, const DevBatch &b, const double *d_corners, cudaStream_t st);
// ... MI (lk_mi.cu)
cudaError_t launch_reinit_mi(int ssm, int threads, const DevBatch &b, const double *d_corners, int n_bins, double pre_seed,
	double *mi_tab, cudaStream_t st);
// ... NCC (lk_ncc.cu)
cudaError_t launch_reinit_ncc(int ssm, int threads, const DevBatch &b, const double *d_corners, cudaStream_t st);
cudaError_t launch_stage(int ssm, int threads, const DevBatch &b, const StageTaps &t, cudaStream_t st);
// peer_gather.cu: corners -> every rank's gathered array (producers other than the update kernels); flags to the peers +
// wait for theirs (one warp; *d_err set on time-out)
cudaError_t launch_peer_push(const double *corners, int P, const PeerOut &peers, cudaStream_t st);
cudaError_t launch_peer_signal_wait(unsigned *const *peer_flags, unsigned *my_flags, int rank, int world, unsigned seq, int *d_err,
	cudaStream_t st);
// lk_ssd_f32.cu
cudaError_t launch_update_ssd_f32(int ssm, int sm, int threads, const DevBatch &b, cudaStream_t st);
cudaError_t launch_stage_f32(int ssm, const DevBatch &b, const StageTapsF32 &t, cudaStream_t st);
// lk_ncc_f32.cu: NCC under ESM / FCLK in the F32 precision (one fp32 sweep per pass)
cudaError_t launch_update_ncc_f32(int ssm, int sm, int threads, const DevBatch &b, cudaStream_t st);
// lk_ssd_mom.cu: FCLK in the F32 precision with column-fixed threads (d_work: per thread {column, first row, rows, 0})
// frame_tensor_map: a CUtensorMap of the frame (2-D, fp32, box MOM_WINP x F32_WIN = 64 x 56) for the TMA window copy, or null
// (window filled by plain loads)
cudaError_t launch_update_ssd_mom(int ssm, int threads, const DevBatch &b, const int4 *d_work, const void *frame_tensor_map, cudaStream_t st);
cudaError_t launch_pf_evaluate_f32(int ssm, const DevBatch &b, const double *d_states, int n_particles, double *d_likelihood,
	double *d_similarity, double alpha, cudaStream_t st);
// lk_ncc.cu
cudaError_t launch_init_ncc(int ssm, int threads, const DevBatch &b, const double *d_corners, cudaStream_t st);
cudaError_t launch_update_ncc(int ssm, int sm, int threads, const DevBatch &b, cudaStream_t st);
// lk_mi.cu (mi_tab: P x 32 doubles of per-template histogram tables)
cudaError_t launch_init_mi(int ssm, int threads, const DevBatch &b, const double *d_corners, int n_bins, double pre_seed,
	double *mi_tab, cudaStream_t st);
cudaError_t launch_update_mi(int ssm, int sm, int threads, const DevBatch &b, int n_bins, double pre_seed, const double *mi_tab,
	cudaStream_t st);
// preproc.cu: uint8 gray / BGR frame (pitch in bytes) -> float frame (pitch in elements); k5 = the 5 kernel taps
cudaError_t launch_preproc_gauss5(const unsigned char *d_src, int src_pitch, int channels, float *d_dst, int dst_pitch, int h, int w,
	const float *k5, cudaStream_t st);
// MI particle evaluation (lk_mi.cu)
cudaError_t launch_pf_evaluate_mi(int ssm, const DevBatch &b, int n_bins, double pre_seed, const double *mi_tab, const double *d_states,
	int n_particles, double *d_likelihood, double *d_similarity, double alpha, cudaStream_t st);
// debug_kernels.cu
cudaError_t launch_debug_qr_solve(int n, int fast, int n_sys, const double *d_A, const double *d_b, double *d_x, int *d_nz, int *d_perm,
	cudaStream_t st);
// pf_kernels.cu
cudaError_t launch_pf_evaluate(int am, int ssm, const DevBatch &b, const double *d_states, int n_particles,
	double *d_likelihood, double *d_similarity, double alpha, cudaStream_t st);

} // namespace mtfb
