// mtfb_api.cu -- the extern "C" boundary (include/mtf_b200.h): context, device buffers, launches.
// Host code only; no arithmetic of the path lives here except the LinSpaced grid of
// utils::getNormUnitSquarePts (Utilities/src/warpUtils.cc:15-33), which is per-context, not per-pixel.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>
#include <cuda.h>
#include "lk_kernels.cuh"
#include "pf_tracker.cuh"
#include "grid_estimator.cuh"

using namespace mtfb;

namespace {

thread_local std::string g_last_error;

mtfb_status fail(mtfb_status st, const char *fmt, ...){
	char buf[512];
	va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
	g_last_error = buf;
	return st;
}
#define CUDA_TRY(expr) do{ cudaError_t e_ = (expr); if(e_ != cudaSuccess) \
	return fail(e_ == cudaErrorMemoryAllocation ? MTFB_ERR_NO_MEMORY : MTFB_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); }while(0)

// DenseBase::LinSpaced(size, low, high), Eigen 3.3 linspaced_op_impl<double, false>
void lin_spaced(std::vector<double> &out, int size, double low, double high){
	out.resize(size);
	int size1 = size == 1 ? 1 : size - 1;
	double step = size == 1 ? 0.0 : (high - low) / double(size - 1);
	bool flip = std::fabs(high) < std::fabs(low);
	for(int i = 0; i < size; ++i){
		if(flip){ out[i] = (i == 0) ? low : (high - (size1 - i)*step); }
		else{ out[i] = (i == size1) ? high : (low + i*step); }
	}
}

bool combo_supported(const mtfb_params *p, const char **why){
	*why = "";
	const bool gn = p->sm == MTFB_SM_ESM || p->sm == MTFB_SM_FCLK || p->sm == MTFB_SM_ICLK;
	if(!(p->ssm == MTFB_SSM_HOMOGRAPHY || p->ssm == MTFB_SSM_AFFINE || p->ssm == MTFB_SSM_TRANSLATION)){ *why = "ssm must be homography, affine or translation"; return false; }
	if(p->ssm == MTFB_SSM_TRANSLATION){
		// Translation (SSM/src/Translation.cc) runs on the SSD skeleton of the F64 precision
		if(p->am != MTFB_AM_SSD || p->precision != MTFB_PRECISION_F64 || p->sm == MTFB_SM_PF){
			*why = "the translation SSM is implemented for SSD in the F64 precision (ESM / FCLK / ICLK / FALK / IALK)"; return false; }
	}
	if(p->sm == MTFB_SM_FALK || p->sm == MTFB_SM_IALK){
		// the additive searches: ssm.cmptPixJacobian / cmptApproxPixJacobian + ssm.additiveUpdate on the SSD skeleton
		if(p->am != MTFB_AM_SSD){ *why = "FALK / IALK are implemented for SSD"; return false; }
		if(p->precision != MTFB_PRECISION_F64){ *why = "FALK / IALK are implemented in the F64 precision"; return false; }
		if(p->hess_type < MTFB_LK_HESS_INITIAL_SELF || p->hess_type > MTFB_LK_HESS_STD){ *why = "unknown Hessian type"; return false; }
		return true;
	}
	if(p->sm == MTFB_SM_PF){
		if(p->am == MTFB_AM_MI && (p->mi_n_bins < 4 || p->mi_n_bins > 16)){ *why = "MI: 4 <= mi_n_bins <= 16"; return false; }
		if(p->am == MTFB_AM_MI && p->mi_pou && p->mi_n_bins < 6){ *why = "MI: partition of unity needs mi_n_bins >= 6"; return false; }
		return p->am == MTFB_AM_SSD || p->am == MTFB_AM_NCC || p->am == MTFB_AM_MI;
	}
	if(!gn){ *why = "sm must be esm, fclk, iclk, falk, ialk or pf"; return false; }
	if(p->am == MTFB_AM_SSD) return true;
	if(p->am == MTFB_AM_NCC){
		// the self Hessians (NCC.cc:337-389) and the Std forms cmptCurrHessian / cmptInitHessian (NCC.cc:282-336) are
		// implemented, for ESM also on the mean pixel Jacobian (Original Jacobian / Hessian)
		bool ok;
		if(p->sm == MTFB_SM_ESM) ok = (p->jac_type == MTFB_ESM_JAC_DIFF_OF_JACS || p->jac_type == MTFB_ESM_JAC_ORIGINAL) &&
			p->hess_type >= MTFB_ESM_HESS_INITIAL_SELF && p->hess_type <= MTFB_ESM_HESS_STD;
		else if(p->sm == MTFB_SM_FCLK) ok = p->hess_type >= MTFB_LK_HESS_INITIAL_SELF && p->hess_type <= MTFB_LK_HESS_STD;
		else ok = p->hess_type >= MTFB_LK_HESS_INITIAL_SELF && p->hess_type <= MTFB_LK_HESS_STD;
		if(!ok) *why = "NCC: unknown Hessian / Jacobian type";
		return ok;
	}
	if(p->am == MTFB_AM_MI){
		// the self Hessians (init_self_hessian and the per-pass cmptSelfHessian, MI.cc:515-594) and the Std forms
		// cmptInitHessian / cmptCurrHessian (MI.cc:461-514, 603-637) are implemented, for ESM also on the mean pixel Jacobian
		const bool ok = (p->sm == MTFB_SM_ESM) ? ((p->jac_type == MTFB_ESM_JAC_DIFF_OF_JACS || p->jac_type == MTFB_ESM_JAC_ORIGINAL) &&
				p->hess_type >= MTFB_ESM_HESS_INITIAL_SELF && p->hess_type <= MTFB_ESM_HESS_STD)
			: (p->hess_type >= MTFB_LK_HESS_INITIAL_SELF && p->hess_type <= MTFB_LK_HESS_STD);
		if(!ok){ *why = "MI: unknown Hessian / Jacobian type"; return false; }
		if(p->mi_n_bins < 4 || p->mi_n_bins > 16){ *why = "MI: 4 <= mi_n_bins <= 16"; return false; }
		if(p->mi_pou && p->mi_n_bins < 6){ *why = "MI: partition of unity needs mi_n_bins >= 6"; return false; }
		return true;
	}
	*why = "am must be ssd, ncc or mi";
	return false;
}

// Which pixels each thread of the moment kernel (lk_ssd_mom.cu) owns: a thread stays in one column, so a column's rows
// are split among n or n + 1 threads (n = T / resx).  Items are ordered longest first, class by class and column by
// column inside a class, so that the 32 threads of a warp have (nearly) the same number of rows and walk adjacent columns
// of the same row.  Returns false if the resolution does not fit (a thread may own at most 64 rows).
bool build_mom_work(int resx, int resy, int T, std::vector<int> &tab /* T x 4 */){
	tab.assign((size_t)T * 4, 0);
	if(resx > T) return false;
	int n_lo = T / resx;
	if(n_lo > resy) n_lo = resy;
	int n_more = (n_lo < resy) ? T - n_lo*resx : 0;          // columns that get n_lo + 1 threads
	if(n_more > resx) n_more = resx;
	struct Item { int col, row0, nrows, cls; };
	std::vector<Item> items;
	for(int c = 0; c < resx; ++c){
		// the columns with more threads are the first n_more: their (shorter) items go last
		const int n = n_lo + (c < n_more ? 1 : 0);
		for(int k = 0; k < n; ++k){
			const int r0 = (int)((long long)resy*k / n), r1 = (int)((long long)resy*(k + 1) / n);
			items.push_back({ c, r0, r1 - r0, k });
		}
	}
	std::stable_sort(items.begin(), items.end(), [](const Item &a, const Item &b){
		if(a.nrows != b.nrows) return a.nrows > b.nrows;
		if(a.cls != b.cls) return a.cls < b.cls;
		return a.col < b.col;
	});
	if((int)items.size() > T || items.empty() || items[0].nrows > 64) return false;
	for(size_t i = 0; i < items.size(); ++i){
		tab[4 * i] = items[i].col; tab[4 * i + 1] = items[i].row0; tab[4 * i + 2] = items[i].nrows;
	}
	return true;
}

cudaError_t launch_init(const mtfb_params &p, int threads, const DevBatch &b, const double *d_corners, double *mi_tab, cudaStream_t st){
	if(p.am == MTFB_AM_MI) return launch_init_mi(p.ssm, threads, b, d_corners, p.mi_n_bins, p.mi_pre_seed, mi_tab, st);
	if(p.am == MTFB_AM_NCC) return launch_init_ncc(p.ssm, threads, b, d_corners, st);
	return launch_init_ssd(p.ssm, threads, b, d_corners, st);
}
cudaError_t launch_update(const mtfb_params &p, int threads, int occ, const DevBatch &b, const double *mi_tab, cudaStream_t st,
	const int4 *mom_work = nullptr, int mom_threads = 0, const void *frame_map = nullptr){
	if(p.precision == MTFB_PRECISION_F32 && mom_work) return launch_update_ssd_mom(p.ssm, mom_threads, b, mom_work, frame_map, st);
	if(p.precision == MTFB_PRECISION_F32 && p.am == MTFB_AM_NCC) return launch_update_ncc_f32(p.ssm, p.sm, threads, b, st);
	if(p.precision == MTFB_PRECISION_F32) return launch_update_ssd_f32(p.ssm, p.sm, threads, b, st);
	if(p.am == MTFB_AM_MI) return launch_update_mi(p.ssm, p.sm, threads, b, p.mi_n_bins, p.mi_pre_seed, mi_tab, st);
	if(p.am == MTFB_AM_NCC) return launch_update_ncc(p.ssm, p.sm, threads, b, st);
	return launch_update_ssd(p.ssm, p.sm, threads, occ, b, st);
}

} // namespace

struct mtfb_ctx {
	mtfb_params prm;
	int S, N, P, threads, occ;
	cudaStream_t own_stream, stream;
	DevBatch b;
	// owned device memory
	float *d_img_own; size_t img_capacity;      // elements
	unsigned char *d_raw; size_t raw_capacity;  // bytes: the raw uint8 frame of mtfb_set_image_u8
	double *d_grid;                             // xv | yv | norm_corners
	double *d_patch;                            // all per-patch fp64 arrays in one allocation
	double *d_mi_tab;                           // MI: P x 32 histogram tables
	float *d_f32;                               // precision F32: xvf | yvf | I0f | G0f
	int *d_ints;                                // n_iters | status
	double *d_corners_in;                       // staging for initialize()/set_region()
	mtfb_iter_log *d_log;
	double *d_scratch; size_t scratch_bytes;    // getters / pf
	bool have_image, initialized;
	long launches;
	// F32 + FCLK: the moment kernel (lk_ssd_mom.cu) and its thread -> pixels table; the Affine SSM takes it only while
	// every initial region is a parallelogram (the kernel's grid frame needs an affine DLT there)
	int4 *d_mom_work; int mom_threads; bool all_parallelograms;
	// sm = PF after mtfb_pf_configure (pf_tracker.cu)
	bool pf_configured; mtfb_pf_params pfp; PFDev pf;
	double *d_pf; int *d_pf_ints;               // particle arrays in one allocation each
	double *d_pf_rand_in; size_t pf_rand_in_capacity; bool pf_normals_pending, pf_uniforms_pending;
	double *d_pf_rand_out;
	long pf_frame;
	// mtfb_set_image_async: frames uploaded on a copy stream into one of two buffers while the previous frame is tracked
	cudaStream_t copy_stream; cudaEvent_t ev_upload, ev_read[2];
	float *d_pre_img[2]; size_t pre_img_capacity[2];
	unsigned char *d_pre_raw[2]; size_t pre_raw_capacity[2];
	Image pre_image[2];
	int pre_next, pre_pending, pre_current;      // slot of the next upload; slot uploaded but not adopted yet (-1); slot b.img points to (-1)
	// tensor map of the current frame for the moment kernel's 2-D TMA window copy (re-encoded when the frame buffer changes)
	// robust warp estimation from point pairs (grid_estimator.cu): staging for host points, outputs, LMedS scratch; the grid's
	// own prev_pts / curr_pts after mtfb_grid_enable
	float *d_est_pts; unsigned char *d_est_mask; float *d_est_err; double *d_est_out; size_t est_capacity;
	float *d_grid_prev, *d_grid_curr; bool grid_enabled;
	// the all-gather over NVLink peer memory (peer_gather.cu): this rank's two gathered arrays + flag row + error word in one
	// allocation, the other ranks' mapped through CUDA IPC
	double *d_peer_buf; int peer_n_total, peer_rank, peer_world, peer_row0; bool peer_attached, peer_pending;
	unsigned peer_seq;
	double *peer_base[MTFB_MAX_PEERS]; unsigned *peer_flags[MTFB_MAX_PEERS]; void *peer_mapped[MTFB_MAX_PEERS];
	alignas(64) CUtensorMap frame_map; const float *frame_map_ptr; int frame_map_h, frame_map_w, frame_map_pitch; bool frame_map_ok;
};
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
static const void *frame_map_for(mtfb_ctx *c){
	if(!c->d_mom_work) return nullptr;
	const Image &im = c->b.img;
	if(c->frame_map_ptr == im.data && c->frame_map_h == im.h && c->frame_map_w == im.w && c->frame_map_pitch == im.pitch)
		return c->frame_map_ok ? &c->frame_map : nullptr;
	c->frame_map_ptr = im.data; c->frame_map_h = im.h; c->frame_map_w = im.w; c->frame_map_pitch = im.pitch; c->frame_map_ok = false;
	if(std::getenv("MTFB_NO_TMA2D")) return nullptr;
	// the copy needs a 16-byte aligned base and row stride, and a frame at least as large as the box
	if(((uintptr_t)im.data & 15) != 0 || (im.pitch & 3) != 0 || im.w < 64 || im.h < 56) return nullptr;
	typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
		const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	static encode_fn encode = nullptr;
	static bool looked = false;
	if(!looked){
		void *ptr = nullptr; cudaDriverEntryPointQueryResult q;
		if(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
			encode = reinterpret_cast<encode_fn>(ptr);
		looked = true;
	}
	if(!encode) return nullptr;
	const cuuint64_t dims[2] = { (cuuint64_t)im.w, (cuuint64_t)im.h }; const cuuint64_t strides[1] = { (cuuint64_t)im.pitch * sizeof(float) };
	const cuuint32_t box[2] = { 64, 56 }; const cuuint32_t estr[2] = { 1, 1 };
	const CUresult r = encode(&c->frame_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(im.data), dims, strides, box, estr,
		CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	c->frame_map_ok = (r == CUDA_SUCCESS);
	return c->frame_map_ok ? &c->frame_map : nullptr;
}
static void peer_out_for(const mtfb_ctx *c, unsigned seq, PeerOut &po){
	po.n = c->peer_world; po.row0 = c->peer_row0;
	for(int r = 0; r < MTFB_MAX_PEERS; ++r)
		po.dst[r] = r < c->peer_world ? c->peer_base[r] + (size_t)(seq & 1u)*c->peer_n_total * 8 : nullptr;
}
static const int4 *mom_work_for(const mtfb_ctx *c){
	if(!c->d_mom_work) return nullptr;
	if(c->prm.ssm == MTFB_SSM_AFFINE && !c->all_parallelograms) return nullptr;
	return c->d_mom_work;
}

extern "C" {

const char *mtfb_last_error(void){ return g_last_error.c_str(); }
const char *mtfb_version(void){ return "mtf_b200 0.1 (sm_100a)"; }
long mtfb_launch_count(const mtfb_ctx *ctx){ return ctx ? ctx->launches : 0; }
int mtfb_state_size(const mtfb_ctx *ctx){ return ctx ? ctx->S : 0; }

void mtfb_default_params(mtfb_params *p){
	// shipped configuration: Config/mtf.cfg:14,24 (max_iters 30, epsilon 1e-4), Config/modules.cfg
	// ({esm,fc,ic}_chained_warp 1, mi_n_bins / pre_seed as parameters.h:344-346), LM off (parameters.h:176)
	std::memset(p, 0, sizeof(*p));
	p->am = MTFB_AM_SSD; p->ssm = MTFB_SSM_HOMOGRAPHY; p->sm = MTFB_SM_FCLK;
	p->resx = 50; p->resy = 50; p->n_patches = 1;
	p->max_iters = 30; p->epsilon = 1e-4;
	p->hess_type = MTFB_LK_HESS_CURRENT_SELF; p->jac_type = MTFB_ESM_JAC_DIFF_OF_JACS;
	p->chained_warp = 1; p->leven_marq = 0; p->lm_delta_init = 0.01; p->lm_delta_update = 10;
	p->nt_semantics = 1; p->grad_eps = 1e-8; p->hom_normalized_init = 0;
	p->mi_n_bins = 8; p->mi_pre_seed = 10; p->mi_pou = 0; p->likelihood_alpha = 1;
	p->device = 0; p->threads_per_patch = 0; p->occupancy = 0; p->precision = MTFB_PRECISION_F64;
	p->f32_solve = MTFB_F32_SOLVE_REFERENCE;
}

mtfb_status mtfb_destroy(mtfb_ctx *c){
	if(!c) return MTFB_OK;
	cudaSetDevice(c->prm.device);
	if(c->own_stream) cudaStreamSynchronize(c->own_stream);
	cudaFree(c->d_img_own); cudaFree(c->d_grid); cudaFree(c->d_patch); cudaFree(c->d_ints);
	cudaFree(c->d_corners_in); cudaFree(c->d_log); cudaFree(c->d_scratch); cudaFree(c->d_f32); cudaFree(c->d_raw);
	cudaFree(c->d_mom_work);
	cudaFree(c->d_pf); cudaFree(c->d_pf_ints); cudaFree(c->d_pf_rand_in); cudaFree(c->d_pf_rand_out);
	cudaFree(c->d_est_pts); cudaFree(c->d_est_mask); cudaFree(c->d_est_err); cudaFree(c->d_est_out);
	cudaFree(c->d_grid_prev); cudaFree(c->d_grid_curr);
	for(int r = 0; r < MTFB_MAX_PEERS; ++r) if(c->peer_mapped[r]) cudaIpcCloseMemHandle(c->peer_mapped[r]);
	cudaFree(c->d_peer_buf);
	if(c->copy_stream){
		cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream);
		cudaEventDestroy(c->ev_upload); cudaEventDestroy(c->ev_read[0]); cudaEventDestroy(c->ev_read[1]);
	}
	for(int k = 0; k < 2; ++k){ cudaFree(c->d_pre_img[k]); cudaFree(c->d_pre_raw[k]); }
	if(c->own_stream) cudaStreamDestroy(c->own_stream);
	delete c;
	return MTFB_OK;
}

mtfb_status mtfb_create(const mtfb_params *p, mtfb_ctx **out){
	if(!p || !out) return fail(MTFB_ERR_INVALID_ARG, "mtfb_create: null argument");
	*out = nullptr;
	if(p->resx < 2 || p->resy < 2 || p->n_patches < 1 || p->max_iters < 1)
		return fail(MTFB_ERR_INVALID_ARG, "mtfb_create: resx/resy >= 2, n_patches >= 1, max_iters >= 1 required");
	const char *why;
	if(!combo_supported(p, &why))
		return fail(MTFB_ERR_NOT_SUPPORTED, "mtfb_create: (am %d, ssm %d, sm %d, hess %d, jac %d) is not implemented: %s",
			p->am, p->ssm, p->sm, p->hess_type, p->jac_type, why);
	if(p->hom_normalized_init && p->ssm == MTFB_SSM_TRANSLATION)
		return fail(MTFB_ERR_INVALID_ARG, "mtfb_create: the Translation SSM has no normalized_init (TranslationParams.h)");
	if(p->hom_normalized_init && p->ssm == MTFB_SSM_AFFINE && p->precision != MTFB_PRECISION_F64)
		return fail(MTFB_ERR_NOT_SUPPORTED, "mtfb_create: the affine normalized_init (computeAffineNDLT, warpUtils.cc:378-386) is "
			"implemented in the F64 precision");
	if(!(p->grad_eps > 0)) return fail(MTFB_ERR_INVALID_ARG, "mtfb_create: grad_eps must be > 0");
	if(p->precision != MTFB_PRECISION_F64 && p->precision != MTFB_PRECISION_F32)
		return fail(MTFB_ERR_INVALID_ARG, "mtfb_create: precision must be MTFB_PRECISION_F64 or MTFB_PRECISION_F32");
	if(p->f32_solve != MTFB_F32_SOLVE_REFERENCE && p->f32_solve != MTFB_F32_SOLVE_LOCAL)
		return fail(MTFB_ERR_INVALID_ARG, "mtfb_create: f32_solve must be MTFB_F32_SOLVE_REFERENCE or MTFB_F32_SOLVE_LOCAL");
	if(p->precision == MTFB_PRECISION_F32){
		const bool gn = p->sm == MTFB_SM_ESM || p->sm == MTFB_SM_FCLK || p->sm == MTFB_SM_ICLK;
		const bool pf = p->sm == MTFB_SM_PF;
		// NCC: ESM / FCLK with the self Hessians (and ESM's DiffOfJacs Jacobian): the one-sweep kernel of lk_ncc_f32.cu
		const bool ncc_f32 = p->am == MTFB_AM_NCC && p->ssm != MTFB_SSM_TRANSLATION && (p->chained_warp || !p->nt_semantics) &&
			((p->sm == MTFB_SM_ESM && p->jac_type == MTFB_ESM_JAC_DIFF_OF_JACS && (p->hess_type == MTFB_ESM_HESS_INITIAL_SELF ||
				p->hess_type == MTFB_ESM_HESS_CURRENT_SELF || p->hess_type == MTFB_ESM_HESS_SUM_OF_SELF)) ||
			 (p->sm == MTFB_SM_FCLK && (p->hess_type == MTFB_LK_HESS_INITIAL_SELF || p->hess_type == MTFB_LK_HESS_CURRENT_SELF)));
		if(!ncc_f32 && (p->am != MTFB_AM_SSD || !(gn || pf) || (gn && !(p->chained_warp || !p->nt_semantics))))
			return fail(MTFB_ERR_NOT_SUPPORTED, "mtfb_create: MTFB_PRECISION_F32 is implemented for SSD with ESM / FCLK / ICLK (chained "
				"warp) and PF, and for NCC with ESM (DiffOfJacs) / FCLK and the self Hessians (chained warp); use MTFB_PRECISION_F64");
		if(!(p->grad_eps < 1e-6)) return fail(MTFB_ERR_NOT_SUPPORTED, "mtfb_create: MTFB_PRECISION_F32 needs grad_eps < 1e-6 (the fp32 "
			"path returns the cell slope, the eps -> 0 limit of the reference's finite difference)");
	}
	// default work split, measured on B200 (profiles/README.md): one warp per patch once the batch alone fills the
	// ~8 warps per SM the fp64 accumulators leave room for; more warps per patch (and a tighter register budget,
	// so that the whole batch is resident in one wave) for smaller batches
	int threads = p->threads_per_patch;
	int occ = p->occupancy;
	if(!threads && p->am == MTFB_AM_MI){
		// MI keeps 18 KB of private histograms per warp (B = 8) next to the N current pixel values in shared memory
		threads = p->n_patches >= 600 ? 32 : 64; occ = 0;
	}
	if(!threads && p->precision == MTFB_PRECISION_F32){
		// measured on B200 (profiles/r01_pscale.txt): four warps per patch while the batch fits one wave of four 128-thread
		// CTAs per SM (148 x 4 = 592 patches: 0.355 ms against 0.412 ms at two warps), two warps per patch above that
		// (888 patches: 0.512 against 0.600 ms), eight warps per patch for small batches
		threads = p->n_patches > 592 ? 64 : p->n_patches >= 150 ? 128 : 256;
		// NCC's one-sweep kernel on small cells (lk_ncc_f32.cu): 10 x 10 cells 0.33 ms at one warp against 0.35 at two
		if(p->am == MTFB_AM_NCC && p->n_patches > 592 && p->resx*p->resy < 256) threads = 32;
	}
	if(!threads){
		if(p->n_patches >= 900){ threads = 32; occ = 0; }
		else if(p->n_patches >= 450){ threads = 64; occ = 2; }
		else if(p->n_patches >= 200){ threads = 128; occ = 2; }
		else{ threads = 256; occ = 2; }
	}
	if(threads != 32 && threads != 64 && threads != 128 && threads != 256)
		return fail(MTFB_ERR_INVALID_ARG, "mtfb_create: threads_per_patch must be 0, 32, 64, 128 or 256");
	if(occ < 0 || occ > 2) return fail(MTFB_ERR_INVALID_ARG, "mtfb_create: occupancy must be 0, 1 or 2");
	int n_dev = 0;
	if(cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
		return fail(MTFB_ERR_CUDA, "mtfb_create: no CUDA device visible (this library has no CPU path)");
	if(p->device < 0 || p->device >= n_dev) return fail(MTFB_ERR_INVALID_ARG, "mtfb_create: device %d of %d", p->device, n_dev);
	CUDA_TRY(cudaSetDevice(p->device));
	cudaDeviceProp prop;
	CUDA_TRY(cudaGetDeviceProperties(&prop, p->device));
	if(prop.major != 10) return fail(MTFB_ERR_CUDA, "mtfb_create: built for sm_100a, device is sm_%d%d", prop.major, prop.minor);

	mtfb_ctx *c = new (std::nothrow) mtfb_ctx();
	if(!c) return fail(MTFB_ERR_NO_MEMORY, "mtfb_create: out of host memory");
	std::memset(static_cast<void*>(c), 0, sizeof(*c));
	c->prm = *p; c->threads = threads; c->occ = occ;
	c->S = p->ssm == MTFB_SSM_HOMOGRAPHY ? 8 : (p->ssm == MTFB_SSM_AFFINE ? 6 : 2);
	c->N = p->resx * p->resy; c->P = p->n_patches;
	const int S = c->S, N = c->N, P = c->P;
	mtfb_status st = MTFB_OK;
	do{
		if(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess){ st = MTFB_ERR_CUDA; break; }
		c->stream = c->own_stream;
		// grid: Homography uses the unit square (Homography.cc:32-48 -> ProjectiveBase.cc:9-18), Affine a
		// pixel-scaled square [1 - res/2, res/2] (Affine.cc:53-56)
		std::vector<double> xv, yv;
		double min_x = -0.5, min_y = -0.5, max_x = 0.5, max_y = 0.5;
		if(p->ssm == MTFB_SSM_AFFINE){ min_x = 1 - p->resx / 2.0; min_y = 1 - p->resy / 2.0; max_x = p->resx / 2.0; max_y = p->resy / 2.0; }
		lin_spaced(xv, p->resx, min_x, max_x);
		lin_spaced(yv, p->resy, min_y, max_y);
		std::vector<double> grid(xv); grid.insert(grid.end(), yv.begin(), yv.end());
		const double nc[8] = { min_x, max_x, max_x, min_x, min_y, min_y, max_y, max_y };
		grid.insert(grid.end(), nc, nc + 8);
		if(cudaMalloc(&c->d_grid, grid.size()*sizeof(double)) != cudaSuccess){ st = MTFB_ERR_NO_MEMORY; break; }
		if(cudaMemcpy(c->d_grid, grid.data(), grid.size()*sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess){ st = MTFB_ERR_CUDA; break; }
		// per-patch arrays
		// SSD with ESM / FCLK keeps the un-chained template gradient for setRegion (NT/ESM.cc:150-168, NT/FCLK.cc:360-376)
		// ... and IALK pushes it through cmptApproxPixJacobian on every pass (NT/IALK.cc:131)
		const bool keep_raw_grad = p->am == MTFB_AM_SSD && (p->sm == MTFB_SM_ESM || p->sm == MTFB_SM_FCLK || p->sm == MTFB_SM_IALK);
		// MI, templates beyond the 24 KB the update kernel keeps in shared memory (lk_mi.cu launch_one): a pass's pixel values
		const bool mi_it_scratch = p->am == MTFB_AM_MI && (size_t)N*sizeof(double) > 24 * 1024 && !std::getenv("MTFB_MI_RESAMPLE");
		// NCC with ESM / FCLK: the pixel values of the last pass and the un-chained template gradient, for setRegion (lk_ncc.cu
		// ncc_reinit_kernel)
		const bool ncc_it_last = (p->am == MTFB_AM_NCC || p->am == MTFB_AM_MI) && (p->sm == MTFB_SM_ESM || p->sm == MTFB_SM_FCLK);   // (and MI: mi_init_kernel<REINIT>)
		size_t per_patch = 9 + 9 + S + 8 + 8 + (size_t)N + 2 * (size_t)N + 64 + 1 + 8 + 32 + 64 + ((keep_raw_grad || ncc_it_last) ? 2 * (size_t)N : 0) +
			((mi_it_scratch || ncc_it_last) ? (size_t)N : 0);
		if(cudaMalloc(&c->d_patch, per_patch*P*sizeof(double)) != cudaSuccess){ st = MTFB_ERR_NO_MEMORY; break; }
		if(cudaMemset(c->d_patch, 0, per_patch*P*sizeof(double)) != cudaSuccess){ st = MTFB_ERR_CUDA; break; }
		if(cudaMalloc(&c->d_ints, 2 * (size_t)P*sizeof(int)) != cudaSuccess){ st = MTFB_ERR_NO_MEMORY; break; }
		if(cudaMemset(c->d_ints, 0, 2 * (size_t)P*sizeof(int)) != cudaSuccess){ st = MTFB_ERR_CUDA; break; }
		if(cudaMalloc(&c->d_corners_in, 8 * (size_t)P*sizeof(double)) != cudaSuccess){ st = MTFB_ERR_NO_MEMORY; break; }
		DevBatch &b = c->b;
		b.P = P; b.N = N; b.resx = p->resx; b.resy = p->resy;
		b.xv = c->d_grid; b.yv = c->d_grid + p->resx; b.norm_corners = c->d_grid + p->resx + p->resy;
		double *q = c->d_patch;
		b.dlt = q; q += 9 * (size_t)P;
		b.warp = q; q += 9 * (size_t)P;
		b.state = q; q += (size_t)S*P;
		b.corners = q; q += 8 * (size_t)P;
		b.init_corners = q; q += 8 * (size_t)P;
		b.Hinit = q; q += 64 * (size_t)P;
		b.f = q; q += (size_t)P;
		b.am_scal = q; q += 8 * (size_t)P;
		c->d_mi_tab = q; q += 32 * (size_t)P;
		b.ncc_tab = q; q += 64 * (size_t)P;
		b.I0 = q; q += (size_t)N*P;
		b.G0 = q; q += 2 * (size_t)N*P;
		b.G0raw = nullptr;
		b.It_scratch = nullptr;
		if(mi_it_scratch || ncc_it_last){ b.It_scratch = q; q += (size_t)N*P; }
		if(keep_raw_grad || ncc_it_last){ b.G0raw = q; q += 2 * (size_t)N*P; }
		b.n_iters = c->d_ints; b.status = c->d_ints + P;
		b.I0f = b.G0f = nullptr; b.I0f_stride = 0;
		b.gx_lo = b.gx_step = b.gy_lo = b.gy_step = 0;
		if(p->precision == MTFB_PRECISION_F32){
			const size_t stride = ((size_t)N + 3) & ~(size_t)3, n_f32 = (stride + 2 * (size_t)N)*P;
			if(cudaMalloc(&c->d_f32, n_f32*sizeof(float)) != cudaSuccess){ st = MTFB_ERR_NO_MEMORY; break; }
			if(cudaMemset(c->d_f32, 0, n_f32*sizeof(float)) != cudaSuccess){ st = MTFB_ERR_CUDA; break; }
			b.I0f = c->d_f32; b.I0f_stride = (int)stride; b.G0f = c->d_f32 + stride*P;
			// the LinSpaced grid as low + i * step: the kernel's coordinate error bound accounts for the <= 1.5 ulp this
			// differs from the fp64 grid values by (lk_ssd_f32.cu pass_constants)
			b.gx_lo = (float)xv[0]; b.gx_step = (float)((xv[p->resx - 1] - xv[0]) / (p->resx - 1));
			b.gy_lo = (float)yv[0]; b.gy_step = (float)((yv[p->resy - 1] - yv[0]) / (p->resy - 1));
		}
		if(p->precision == MTFB_PRECISION_F32 && p->am == MTFB_AM_SSD && p->sm == MTFB_SM_FCLK){
			// the moment kernel: four warps per patch (one per SM sub-partition) from ~1000 pixels per patch on; fewer threads
			// for small cells so that a thread still owns several rows.  MTFB_F32_KERNEL=classic selects the first-generation
			// kernel (experiments, A/B timing)
			const char *sel = std::getenv("MTFB_F32_KERNEL");
			int mt = p->threads_per_patch;
			if(!mt) mt = N >= 1024 ? 128 : (N >= 400 ? 64 : 32);
			std::vector<int> tab;
			bool ok = !(sel && std::strcmp(sel, "classic") == 0) && build_mom_work(p->resx, p->resy, mt, tab);
			if(!ok && !p->threads_per_patch && !(sel && std::strcmp(sel, "classic") == 0)){
				mt = 256; ok = build_mom_work(p->resx, p->resy, mt, tab);
			}
			if(ok){
				if(cudaMalloc(&c->d_mom_work, tab.size()*sizeof(int)) != cudaSuccess){ st = MTFB_ERR_NO_MEMORY; break; }
				if(cudaMemcpy(c->d_mom_work, tab.data(), tab.size()*sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess){ st = MTFB_ERR_CUDA; break; }
				c->mom_threads = mt;
			}
		}
		b.n_iters_prof = nullptr;
#if MTFB_PROF
		if(cudaMalloc(&b.n_iters_prof, 64 * sizeof(long long)) != cudaSuccess || cudaMemset(b.n_iters_prof, 0, 64 * sizeof(long long)) != cudaSuccess){ st = MTFB_ERR_NO_MEMORY; break; }
#endif
		b.log = nullptr; b.log_slots = 0;
		b.max_iters = p->max_iters; b.hess_type = p->hess_type; b.jac_type = p->jac_type;
		b.leven_marq = p->leven_marq; b.nt_semantics = p->nt_semantics;
		// the templated search methods have no chained_warp switch: always chained (ESM.cc:94-95, FCLK.cc:82-86, ICLK.cc:80-90)
		b.chained = (p->chained_warp || !p->nt_semantics) ? 1 : 0;
		b.additive_sm = (p->sm == MTFB_SM_FALK || p->sm == MTFB_SM_IALK) ? 1 : 0;
		// FALK / IALK have no such switch: am->updatePixGrad(ssm->getPts()) (NT/FALK.cc:166, NT/IALK.cc:64)
		if(p->sm == MTFB_SM_FALK || p->sm == MTFB_SM_IALK) b.chained = 1;
		b.norm_init = p->hom_normalized_init ? 1 : 0;
		b.f32_local_solve = (p->f32_solve == MTFB_F32_SOLVE_LOCAL) ? 1 : 0;
		b.epsilon = p->epsilon; b.lm_delta_init = p->lm_delta_init; b.lm_delta_update = p->lm_delta_update;
		b.grad_eps = p->grad_eps;
		b.pix_mult = 1; b.pix_add = 0;
		if(p->am == MTFB_AM_MI){
			// MI::MI (MI.cc:84-94): pixel values are rescaled to histogram-bin units
			double norm_pix_min = 0, norm_pix_max = p->mi_n_bins - 1;
			if(p->mi_pou){ norm_pix_min = 1; norm_pix_max = p->mi_n_bins - 2; }
			b.pix_mult = (norm_pix_max - norm_pix_min) / (255.0 - 0.0 + 1);
			b.pix_add = norm_pix_min;
		}
		b.grad_mult = b.pix_mult / (2 * p->grad_eps);
		b.img = make_image(nullptr, 0, 0, 0);
		c->pre_pending = -1; c->pre_current = -1; c->pre_next = 0;
	} while(0);
	if(st != MTFB_OK){
		cudaError_t e = cudaGetLastError();
		fail(st, "mtfb_create: device allocation failed: %s", cudaGetErrorString(e));
		std::string keep = g_last_error;
		mtfb_destroy(c);
		g_last_error = keep;
		return st;
	}
	*out = c;
	return MTFB_OK;
}

mtfb_status mtfb_set_stream(mtfb_ctx *c, void *cuda_stream){
	if(!c) return fail(MTFB_ERR_INVALID_ARG, "mtfb_set_stream: null context");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	c->stream = static_cast<cudaStream_t>(cuda_stream);
	return MTFB_OK;
}

mtfb_status mtfb_synchronize(mtfb_ctx *c){
	if(!c) return fail(MTFB_ERR_INVALID_ARG, "mtfb_synchronize: null context");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return MTFB_OK;
}

mtfb_status mtfb_set_image(mtfb_ctx *c, const float *host_img, int h, int w, int row_stride){
	if(!c || !host_img) return fail(MTFB_ERR_INVALID_ARG, "mtfb_set_image: null argument");
	if(h < 2 || w < 2 || row_stride < w) return fail(MTFB_ERR_INVALID_ARG, "mtfb_set_image: bad geometry %d x %d stride %d", h, w, row_stride);
	CUDA_TRY(cudaSetDevice(c->prm.device));
	const int pitch = (w + 31) & ~31;            // 128-byte rows
	const size_t need = (size_t)pitch*h;
	if(need > c->img_capacity){
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		cudaFree(c->d_img_own); c->d_img_own = nullptr; c->img_capacity = 0;
		CUDA_TRY(cudaMalloc(&c->d_img_own, need*sizeof(float)));
		c->img_capacity = need;
	}
	CUDA_TRY(cudaMemcpy2DAsync(c->d_img_own, (size_t)pitch*sizeof(float), host_img, (size_t)row_stride*sizeof(float),
		(size_t)w*sizeof(float), h, cudaMemcpyHostToDevice, c->stream));
	c->b.img = make_image(c->d_img_own, h, w, pitch);
	c->have_image = true; c->pre_current = -1;
	return MTFB_OK;
}

mtfb_status mtfb_set_image_u8(mtfb_ctx *c, const unsigned char *host_img, int h, int w, int row_stride, int channels, int kernel_size,
	double sigma){
	if(!c || !host_img) return fail(MTFB_ERR_INVALID_ARG, "mtfb_set_image_u8: null argument");
	if(channels != 1 && channels != 3) return fail(MTFB_ERR_INVALID_ARG, "mtfb_set_image_u8: channels must be 1 (gray) or 3 (BGR)");
	if(h < 3 || w < 3 || row_stride < w*channels) return fail(MTFB_ERR_INVALID_ARG, "mtfb_set_image_u8: bad geometry %d x %d x %d stride %d", h, w, channels, row_stride);
	if(kernel_size != 5) return fail(MTFB_ERR_NOT_SUPPORTED, "mtfb_set_image_u8: gauss_kernel_size %d (only the default 5 is implemented)", kernel_size);
	if(!(sigma > 0)) return fail(MTFB_ERR_INVALID_ARG, "mtfb_set_image_u8: sigma must be > 0");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	const int pitch = (w + 31) & ~31;
	const size_t need = (size_t)pitch*h, raw_pitch = ((size_t)w*channels + 127) & ~(size_t)127, raw_need = raw_pitch*h;
	if(need > c->img_capacity || raw_need > c->raw_capacity){
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		if(need > c->img_capacity){
			cudaFree(c->d_img_own); c->d_img_own = nullptr; c->img_capacity = 0;
			CUDA_TRY(cudaMalloc(&c->d_img_own, need*sizeof(float)));
			c->img_capacity = need;
		}
		if(raw_need > c->raw_capacity){
			cudaFree(c->d_raw); c->d_raw = nullptr; c->raw_capacity = 0;
			CUDA_TRY(cudaMalloc(&c->d_raw, raw_need));
			c->raw_capacity = raw_need;
		}
	}
	CUDA_TRY(cudaMemcpy2DAsync(c->d_raw, raw_pitch, host_img, (size_t)row_stride, (size_t)w*channels, h, cudaMemcpyHostToDevice, c->stream));
	// cv::getGaussianKernel(5, sigma, CV_32F) as OpenCV 2.4 / 3.x computes it
	float k5[5];
	{
		const double scale2X = -0.5 / (sigma*sigma);
		double sum = 0;
		for(int i = 0; i < 5; ++i){ const double x = i - 2.0; k5[i] = (float)std::exp(scale2X*x*x); sum += k5[i]; }
		sum = 1. / sum;
		for(int i = 0; i < 5; ++i) k5[i] = (float)(k5[i] * sum);
	}
	CUDA_TRY(launch_preproc_gauss5(c->d_raw, (int)raw_pitch, channels, c->d_img_own, pitch, h, w, k5, c->stream));
	++c->launches;
	c->b.img = make_image(c->d_img_own, h, w, pitch);
	c->have_image = true; c->pre_current = -1;
	return MTFB_OK;
}

// ---- prefetched frames (mtfb_set_image_async / mtfb_set_image_u8_async)
static cudaError_t gauss_kernel5(double sigma, float *k5){
	// cv::getGaussianKernel(5, sigma, CV_32F) as OpenCV 2.4 / 3.x computes it
	const double scale2X = -0.5 / (sigma*sigma);
	double sum = 0;
	for(int i = 0; i < 5; ++i){ const double x = i - 2.0; k5[i] = (float)std::exp(scale2X*x*x); sum += k5[i]; }
	sum = 1. / sum;
	for(int i = 0; i < 5; ++i) k5[i] = (float)(k5[i] * sum);
	return cudaSuccess;
}
// called by everything that samples the frame: the launch waits for the pending upload and reads that buffer from now on
static mtfb_status adopt_prefetched(mtfb_ctx *c){
	if(c->pre_pending < 0) return MTFB_OK;
	CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_upload, 0));
	c->b.img = c->pre_image[c->pre_pending];
	c->pre_current = c->pre_pending; c->pre_pending = -1;
	c->have_image = true;
	return MTFB_OK;
}
// after a launch that sampled b.img: the copy stream must not overwrite that buffer before the launch has finished
static mtfb_status mark_frame_read(mtfb_ctx *c){
	if(c->pre_current >= 0) CUDA_TRY(cudaEventRecord(c->ev_read[c->pre_current], c->stream));
	return MTFB_OK;
}
static mtfb_status prefetch_begin(mtfb_ctx *c, size_t need_img, size_t need_raw, int *slot_out){
	CUDA_TRY(cudaSetDevice(c->prm.device));
	if(!c->copy_stream){
		CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
		CUDA_TRY(cudaEventCreateWithFlags(&c->ev_upload, cudaEventDisableTiming));
		CUDA_TRY(cudaEventCreateWithFlags(&c->ev_read[0], cudaEventDisableTiming));
		CUDA_TRY(cudaEventCreateWithFlags(&c->ev_read[1], cudaEventDisableTiming));
	}
	if(c->pre_pending >= 0) return fail(MTFB_ERR_LOGIC, "mtfb_set_image_async: the previous prefetched frame has not been consumed by "
		"initialize() / update() yet (one frame may be in flight)");
	const int slot = c->pre_next;
	if(need_img > c->pre_img_capacity[slot] || need_raw > c->pre_raw_capacity[slot]){
		CUDA_TRY(cudaStreamSynchronize(c->stream)); CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
		if(need_img > c->pre_img_capacity[slot]){
			cudaFree(c->d_pre_img[slot]); c->d_pre_img[slot] = nullptr; c->pre_img_capacity[slot] = 0;
			CUDA_TRY(cudaMalloc(&c->d_pre_img[slot], need_img*sizeof(float)));
			c->pre_img_capacity[slot] = need_img;
		}
		if(need_raw > c->pre_raw_capacity[slot]){
			cudaFree(c->d_pre_raw[slot]); c->d_pre_raw[slot] = nullptr; c->pre_raw_capacity[slot] = 0;
			CUDA_TRY(cudaMalloc(&c->d_pre_raw[slot], need_raw));
			c->pre_raw_capacity[slot] = need_raw;
		}
	}
	// the buffer of this slot was last sampled by the launch that recorded ev_read[slot] (two frames ago)
	CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_read[slot], 0));
	*slot_out = slot;
	return MTFB_OK;
}
static mtfb_status prefetch_end(mtfb_ctx *c, int slot, int h, int w, int pitch){
	CUDA_TRY(cudaEventRecord(c->ev_upload, c->copy_stream));
	c->pre_image[slot] = make_image(c->d_pre_img[slot], h, w, pitch);
	c->pre_pending = slot; c->pre_next = slot ^ 1;
	return MTFB_OK;
}

mtfb_status mtfb_set_image_async(mtfb_ctx *c, const float *host_img, int h, int w, int row_stride){
	if(!c || !host_img) return fail(MTFB_ERR_INVALID_ARG, "mtfb_set_image_async: null argument");
	if(h < 2 || w < 2 || row_stride < w) return fail(MTFB_ERR_INVALID_ARG, "mtfb_set_image_async: bad geometry %d x %d stride %d", h, w, row_stride);
	const int pitch = (w + 31) & ~31;
	int slot;
	mtfb_status st = prefetch_begin(c, (size_t)pitch*h, 0, &slot);
	if(st != MTFB_OK) return st;
	CUDA_TRY(cudaMemcpy2DAsync(c->d_pre_img[slot], (size_t)pitch*sizeof(float), host_img, (size_t)row_stride*sizeof(float),
		(size_t)w*sizeof(float), h, cudaMemcpyHostToDevice, c->copy_stream));
	return prefetch_end(c, slot, h, w, pitch);
}

mtfb_status mtfb_set_image_u8_async(mtfb_ctx *c, const unsigned char *host_img, int h, int w, int row_stride, int channels,
	int kernel_size, double sigma){
	if(!c || !host_img) return fail(MTFB_ERR_INVALID_ARG, "mtfb_set_image_u8_async: null argument");
	if(channels != 1 && channels != 3) return fail(MTFB_ERR_INVALID_ARG, "mtfb_set_image_u8_async: channels must be 1 (gray) or 3 (BGR)");
	if(h < 3 || w < 3 || row_stride < w*channels) return fail(MTFB_ERR_INVALID_ARG, "mtfb_set_image_u8_async: bad geometry %d x %d x %d stride %d", h, w, channels, row_stride);
	if(kernel_size != 5) return fail(MTFB_ERR_NOT_SUPPORTED, "mtfb_set_image_u8_async: gauss_kernel_size %d (only the default 5 is implemented)", kernel_size);
	if(!(sigma > 0)) return fail(MTFB_ERR_INVALID_ARG, "mtfb_set_image_u8_async: sigma must be > 0");
	const int pitch = (w + 31) & ~31;
	const size_t raw_pitch = ((size_t)w*channels + 127) & ~(size_t)127;
	int slot;
	mtfb_status st = prefetch_begin(c, (size_t)pitch*h, raw_pitch*h, &slot);
	if(st != MTFB_OK) return st;
	CUDA_TRY(cudaMemcpy2DAsync(c->d_pre_raw[slot], raw_pitch, host_img, (size_t)row_stride, (size_t)w*channels, h, cudaMemcpyHostToDevice, c->copy_stream));
	float k5[5];
	gauss_kernel5(sigma, k5);
	CUDA_TRY(launch_preproc_gauss5(c->d_pre_raw[slot], (int)raw_pitch, channels, c->d_pre_img[slot], pitch, h, w, k5, c->copy_stream));
	++c->launches;
	return prefetch_end(c, slot, h, w, pitch);
}

mtfb_status mtfb_get_image(mtfb_ctx *c, float *out){
	if(!c || !out) return fail(MTFB_ERR_INVALID_ARG, "mtfb_get_image: null argument");
	{ mtfb_status st = adopt_prefetched(c); if(st != MTFB_OK) return st; }
	if(!c->have_image) return fail(MTFB_ERR_LOGIC, "mtfb_get_image: setImage has not been called");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	CUDA_TRY(cudaMemcpy2DAsync(out, (size_t)c->b.img.w*sizeof(float), c->b.img.data, (size_t)c->b.img.pitch*sizeof(float),
		(size_t)c->b.img.w*sizeof(float), c->b.img.h, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return MTFB_OK;
}

mtfb_status mtfb_set_image_device(mtfb_ctx *c, const float *dev_img, int h, int w, int pitch){
	if(!c || !dev_img) return fail(MTFB_ERR_INVALID_ARG, "mtfb_set_image_device: null argument");
	if(h < 2 || w < 2 || pitch < w) return fail(MTFB_ERR_INVALID_ARG, "mtfb_set_image_device: bad geometry %d x %d pitch %d", h, w, pitch);
	c->b.img = make_image(dev_img, h, w, pitch);
	c->have_image = true; c->pre_current = -1;
	return MTFB_OK;
}

static mtfb_status upload_corners(mtfb_ctx *c, const double *corners, const char *who){
	if(!c || !corners) return fail(MTFB_ERR_INVALID_ARG, "%s: null argument", who);
	{ mtfb_status st0 = adopt_prefetched(c); if(st0 != MTFB_OK) return st0; }
	if(!c->have_image) return fail(MTFB_ERR_LOGIC, "%s: setImage has not been called", who);
	for(size_t i = 0; i < 8 * (size_t)c->P; ++i)
		if(!std::isfinite(corners[i])) return fail(MTFB_ERR_INVALID_ARG, "%s: non-finite corner coordinate in patch %zu", who, i / 8);
	c->all_parallelograms = true;
	for(size_t q = 0; q < (size_t)c->P && c->all_parallelograms; ++q){
		const double *k = corners + 8 * q;
		double scale = 0;
		for(int i = 0; i < 8; ++i) scale = std::max(scale, std::fabs(k[i]));
		// UL - UR + LR - LL = 0 in both coordinates
		if(std::fabs(k[0] - k[1] + k[2] - k[3]) > 1e-9*scale || std::fabs(k[4] - k[5] + k[6] - k[7]) > 1e-9*scale) c->all_parallelograms = false;
	}
	CUDA_TRY(cudaSetDevice(c->prm.device));
	CUDA_TRY(cudaMemcpyAsync(c->d_corners_in, corners, 8 * (size_t)c->P*sizeof(double), cudaMemcpyHostToDevice, c->stream));
	// the host buffer may be pageable: wait, so that the caller can reuse it (the copy is 64 B per patch)
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return MTFB_OK;
}

mtfb_status mtfb_initialize(mtfb_ctx *c, const double *corners){
	mtfb_status st = upload_corners(c, corners, "mtfb_initialize");
	if(st != MTFB_OK) return st;
	CUDA_TRY(launch_init(c->prm, c->threads, c->b, c->d_corners_in, c->d_mi_tab, c->stream));
	++c->launches;
	{ mtfb_status st1 = mark_frame_read(c); if(st1 != MTFB_OK) return st1; }
	c->initialized = true;
	if(c->grid_enabled){
		// GridTracker::resetTrackers (GridTracker.cc:389): prev_pts = centroid of the cell's region
		CUDA_TRY(launch_centroids(c->b.corners, c->P, c->d_grid_prev, c->stream));
		++c->launches;
	}
	if(c->pf_configured){
		// PF::initialize (NT/PF.cc:136-183): initializeParticles, prev_corners = ssm->getCorners()
		CUDA_TRY(launch_pf_init_particles(c->prm.ssm, c->pf, c->b, true, c->stream));
		++c->launches;
		c->pf_frame = 0;
	}
	return MTFB_OK;
}

mtfb_status mtfb_set_region(mtfb_ctx *c, const double *corners){
	if(c && !c->initialized) return fail(MTFB_ERR_LOGIC, "mtfb_set_region: initialize has not been called");
	bool ssm_only = true, ncc_reinit = false, mi_reinit = false;
	if(c){
		ssm_only = (c->prm.sm == MTFB_SM_ICLK) || (c->prm.sm == MTFB_SM_PF) || (c->prm.sm == MTFB_SM_FALK) || (c->prm.sm == MTFB_SM_IALK) ||
			(c->prm.sm == MTFB_SM_FCLK && c->prm.hess_type != MTFB_LK_HESS_INITIAL_SELF);
		// ESM and FCLK-InitialSelf also rebuild the template Jacobian and init_self_hessian at the new points: SSD, and NCC
		// from an identity start (its kept template gradient is then the un-chained one)
		ncc_reinit = !ssm_only && c->prm.am == MTFB_AM_NCC && c->b.It_scratch && c->b.G0raw;
		mi_reinit = !ssm_only && c->prm.am == MTFB_AM_MI && c->b.It_scratch && c->b.G0raw;
		if(!ssm_only && !(c->prm.am == MTFB_AM_SSD && c->b.G0raw) && !ncc_reinit && !mi_reinit) return fail(MTFB_ERR_NOT_SUPPORTED,
			"mtfb_set_region: this context keeps neither the un-chained template gradient nor the last pass's pixel values");
	}
	mtfb_status st = upload_corners(c, corners, "mtfb_set_region");
	if(st != MTFB_OK) return st;
	if(ssm_only) CUDA_TRY(launch_set_region(c->prm.ssm, c->b, c->d_corners_in, c->stream));
	else if(ncc_reinit) CUDA_TRY(launch_reinit_ncc(c->prm.ssm, c->threads, c->b, c->d_corners_in, c->stream));
	else if(mi_reinit) CUDA_TRY(launch_reinit_mi(c->prm.ssm, c->threads, c->b, c->d_corners_in, c->prm.mi_n_bins, c->prm.mi_pre_seed, c->d_mi_tab, c->stream));
	else CUDA_TRY(launch_reinit_ssd(c->prm.ssm, c->threads, c->b, c->d_corners_in, c->stream));
	++c->launches;
	{ mtfb_status st1 = mark_frame_read(c); if(st1 != MTFB_OK) return st1; }
	if(c->grid_enabled){
		CUDA_TRY(launch_centroids(c->b.corners, c->P, c->d_grid_prev, c->stream));
		++c->launches;
	}
	if(c->pf_configured){
		// PF::setRegion (NT/PF.cc:596-600): ssm->setCorners, initializeParticles
		CUDA_TRY(launch_pf_init_particles(c->prm.ssm, c->pf, c->b, false, c->stream));
		++c->launches;
	}
	return MTFB_OK;
}

// ------------------------------------------------------------------------------------------------ particle filter tracker
static mtfb_status pf_update(mtfb_ctx *c){
	// PF::update (NT/PF.cc:207-446) for all objects
	CUDA_TRY(cudaSetDevice(c->prm.device));
	const int ssm = c->prm.ssm;
	const size_t n = (size_t)c->P*c->pfp.n_particles;
	PFDev pf = c->pf;
	++c->pf_frame;
	CUDA_TRY(launch_pf_frame_begin(pf, c->b, c->stream)); ++c->launches;
	for(int it = 0; it < c->pfp.max_iters; ++it){
		const unsigned tag = (unsigned)(c->pf_frame*(long)c->pfp.max_iters + it);
		pf.normals_in = c->pf_normals_pending ? c->d_pf_rand_in + (size_t)it*n*pf.n_normals : nullptr;
		pf.uniforms_in = c->pf_uniforms_pending ? c->d_pf_rand_in + (size_t)c->pfp.max_iters*n*pf.n_normals + (size_t)it*n : nullptr;
		pf.normals_out = c->d_pf_rand_out ? c->d_pf_rand_out + (size_t)it*n*pf.n_normals : nullptr;
		pf.uniforms_out = c->d_pf_rand_out ? c->d_pf_rand_out + (size_t)c->pfp.max_iters*n*pf.n_normals + (size_t)it*n : nullptr;
		CUDA_TRY(launch_pf_perturb(ssm, pf, c->b, tag, c->stream)); ++c->launches;
		// ssm->setState -> am->updatePixVals -> am->updateSimilarity(false) -> am->getLikelihood() of every particle
		mtfb_status st = mtfb_pf_evaluate_device(c, pf.states, c->pfp.n_particles, pf.weights, pf.similarity);
		if(st != MTFB_OK) return st;
		CUDA_TRY(launch_pf_weights(pf, c->b, c->stream)); ++c->launches;
		if(c->pfp.resampling_type != MTFB_PF_RESAMPLE_NONE){
			CUDA_TRY(launch_pf_resample(ssm, pf, c->b, tag, c->stream)); c->launches += 2;
		}
		CUDA_TRY(launch_pf_mean(ssm, pf, c->b, c->stream)); ++c->launches;
	}
	if(c->pfp.reset_to_mean){ CUDA_TRY(launch_pf_init_particles(ssm, pf, c->b, false, c->stream)); ++c->launches; }
	c->pf_normals_pending = c->pf_uniforms_pending = false;
	return MTFB_OK;
}

void mtfb_pf_default_params(mtfb_pf_params *p){
	std::memset(p, 0, sizeof(*p));
	p->n_particles = 500; p->max_iters = 1; p->epsilon = 0.01;
	p->dynamic_model = MTFB_PF_AUTO_REGRESSION1; p->update_type = MTFB_PF_UPDATE_COMPOSITIONAL;
	p->likelihood_func = MTFB_PF_LIKELIHOOD_AM; p->resampling_type = MTFB_PF_RESAMPLE_BINARY_MULTINOMIAL;
	p->mean_type = MTFB_PF_MEAN_NONE; p->reset_to_mean = 0;
	p->adaptive_resampling_thresh = 0.2; p->measurement_sigma = 0.1; p->ar_coeff = 0.5;
	p->corner_based_sampling = 1; p->seed = 0x9E3779B97F4A7C15ull; p->object_offset = 0; p->record_randoms = 0;
}

mtfb_status mtfb_pf_configure(mtfb_ctx *c, const mtfb_pf_params *p){
	if(!c || !p) return fail(MTFB_ERR_INVALID_ARG, "mtfb_pf_configure: null argument");
	if(c->prm.sm != MTFB_SM_PF) return fail(MTFB_ERR_LOGIC, "mtfb_pf_configure: the context was not created with sm = MTFB_SM_PF");
	if(c->prm.ssm != MTFB_SSM_HOMOGRAPHY) return fail(MTFB_ERR_NOT_SUPPORTED, "mtfb_pf_configure: the Affine samplers (Affine.cc:464-560) "
		"are not implemented; particle evaluation (mtfb_pf_evaluate) works for both");
	if(p->n_particles < 1 || p->max_iters < 1) return fail(MTFB_ERR_INVALID_ARG, "mtfb_pf_configure: n_particles, max_iters >= 1 required");
	if(p->dynamic_model < 0 || p->dynamic_model > 1 || p->update_type < 0 || p->update_type > 1 || p->likelihood_func < 0 ||
		p->likelihood_func > 2 || p->mean_type < 0 || p->mean_type > 2 || p->resampling_type < 0 || p->resampling_type > 3)
		return fail(MTFB_ERR_INVALID_ARG, "mtfb_pf_configure: enum parameter out of range");
	if(p->resampling_type == MTFB_PF_RESAMPLE_RESIDUAL) return fail(MTFB_ERR_NOT_SUPPORTED, "mtfb_pf_configure: Residual resampling "
		"(NT/PF.cc:542-585, an unstable std::sort over all but the last particle) is not implemented");
	if(p->likelihood_func != MTFB_PF_LIKELIHOOD_AM && c->prm.am == MTFB_AM_MI) return fail(MTFB_ERR_NOT_SUPPORTED,
		"mtfb_pf_configure: MI particles support likelihood_func = AM only");
	if(p->likelihood_func == MTFB_PF_LIKELIHOOD_GAUSSIAN && !(p->measurement_sigma > 0)) return fail(MTFB_ERR_INVALID_ARG,
		"mtfb_pf_configure: measurement_sigma must be > 0");
	for(int i = 0; i < 8; ++i) if(!(p->ssm_sigma[i] >= 0) || !std::isfinite(p->ssm_mean[i])) return fail(MTFB_ERR_INVALID_ARG,
		"mtfb_pf_configure: ssm_sigma must be >= 0 and ssm_mean finite");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	cudaFree(c->d_pf); cudaFree(c->d_pf_ints); cudaFree(c->d_pf_rand_out); cudaFree(c->d_pf_rand_in);
	c->d_pf = nullptr; c->d_pf_ints = nullptr; c->d_pf_rand_out = nullptr; c->d_pf_rand_in = nullptr; c->pf_rand_in_capacity = 0;
	c->pf_configured = false;
	const size_t P = c->P, n = p->n_particles, S = c->S, pn = P*n;
	const size_t n_d = 4 * pn*S + 3 * pn + 8 * P;
	CUDA_TRY(cudaMalloc(&c->d_pf, n_d*sizeof(double)));
	CUDA_TRY(cudaMemset(c->d_pf, 0, n_d*sizeof(double)));
	CUDA_TRY(cudaMalloc(&c->d_pf_ints, (pn + 3 * P)*sizeof(int)));
	CUDA_TRY(cudaMemset(c->d_pf_ints, 0, (pn + 3 * P)*sizeof(int)));
	PFDev &pf = c->pf;
	std::memset(&pf, 0, sizeof(pf));
	pf.n_particles = p->n_particles;
	pf.corner_based = (p->corner_based_sampling != 0) ? 1 : 0;
	pf.n_normals = pf.corner_based ? 10 : (int)S;
	double *q = c->d_pf;
	pf.states = q; q += pn*S; pf.ar = q; q += pn*S; pf.states_other = q; q += pn*S; pf.ar_other = q; q += pn*S;
	pf.weights = q; q += pn; pf.cum_weights = q; q += pn; pf.similarity = q; q += pn; pf.prev_corners = q; q += 8 * P;
	int *qi = c->d_pf_ints;
	pf.src_id = qi; qi += pn; pf.max_wt_id = qi; qi += P; pf.resample_flag = qi; qi += P; pf.done = qi; qi += P;
	for(int i = 0; i < 8; ++i){ pf.sigma[i] = p->ssm_sigma[i]; pf.mean[i] = p->ssm_mean[i]; }
	pf.dynamic_model = p->dynamic_model; pf.update_type = p->update_type; pf.likelihood_func = p->likelihood_func;
	pf.resampling_type = p->resampling_type; pf.mean_type = p->mean_type;
	pf.adaptive = (p->adaptive_resampling_thresh > 0 && p->adaptive_resampling_thresh <= 1) ? 1 : 0;       // NT/PF.cc:113-117
	pf.min_eff_particles = p->adaptive_resampling_thresh*p->n_particles;
	pf.weights_in_smem = (n*sizeof(double) <= 160 * 1024) ? 1 : 0;
	const double pi = 3.14159265358979323846;
	pf.measurement_factor = 1.0 / std::sqrt(2 * pi*p->measurement_sigma);                                  // NT/PF.cc:81-82
	pf.measurement_sigma = p->measurement_sigma;
	// max_similarity = am->getSimilarity() after initializeSimilarity (NT/PF.cc:155-156): 0 for SSD, 1 for NCC
	pf.max_similarity = c->prm.am == MTFB_AM_NCC ? 1.0 : 0.0;
	pf.ar_coeff = p->ar_coeff; pf.epsilon = p->epsilon; pf.seed = p->seed; pf.object_offset = p->object_offset;
	if(p->record_randoms){
		const size_t n_r = (size_t)p->max_iters*pn*(pf.n_normals + 1);
		CUDA_TRY(cudaMalloc(&c->d_pf_rand_out, n_r*sizeof(double)));
		CUDA_TRY(cudaMemset(c->d_pf_rand_out, 0, n_r*sizeof(double)));
	}
	c->pfp = *p;
	c->pf_configured = true;
	c->pf_normals_pending = c->pf_uniforms_pending = false;
	c->pf_frame = 0;
	if(c->initialized){ CUDA_TRY(launch_pf_init_particles(c->prm.ssm, c->pf, c->b, true, c->stream)); ++c->launches; }
	return MTFB_OK;
}

mtfb_status mtfb_pf_set_random_stream(mtfb_ctx *c, const double *normals, const double *uniforms){
	if(!c) return fail(MTFB_ERR_INVALID_ARG, "mtfb_pf_set_random_stream: null context");
	if(!c->pf_configured) return fail(MTFB_ERR_LOGIC, "mtfb_pf_set_random_stream: mtfb_pf_configure has not been called");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	const size_t pn = (size_t)c->P*c->pfp.n_particles, n_norm = (size_t)c->pfp.max_iters*pn*c->pf.n_normals, n_uni = (size_t)c->pfp.max_iters*pn;
	if(!c->d_pf_rand_in){
		CUDA_TRY(cudaMalloc(&c->d_pf_rand_in, (n_norm + n_uni)*sizeof(double)));
		c->pf_rand_in_capacity = n_norm + n_uni;
	}
	if(normals) CUDA_TRY(cudaMemcpyAsync(c->d_pf_rand_in, normals, n_norm*sizeof(double), cudaMemcpyHostToDevice, c->stream));
	if(uniforms) CUDA_TRY(cudaMemcpyAsync(c->d_pf_rand_in + n_norm, uniforms, n_uni*sizeof(double), cudaMemcpyHostToDevice, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	c->pf_normals_pending = normals != nullptr; c->pf_uniforms_pending = uniforms != nullptr;
	return MTFB_OK;
}

mtfb_status mtfb_pf_get_random_stream(mtfb_ctx *c, double *normals, double *uniforms){
	if(!c) return fail(MTFB_ERR_INVALID_ARG, "mtfb_pf_get_random_stream: null context");
	if(!c->pf_configured || !c->d_pf_rand_out) return fail(MTFB_ERR_LOGIC, "mtfb_pf_get_random_stream: configure with record_randoms = 1");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	const size_t pn = (size_t)c->P*c->pfp.n_particles, n_norm = (size_t)c->pfp.max_iters*pn*c->pf.n_normals, n_uni = (size_t)c->pfp.max_iters*pn;
	if(normals) CUDA_TRY(cudaMemcpyAsync(normals, c->d_pf_rand_out, n_norm*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	if(uniforms) CUDA_TRY(cudaMemcpyAsync(uniforms, c->d_pf_rand_out + n_norm, n_uni*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return MTFB_OK;
}

mtfb_status mtfb_pf_get_particles(mtfb_ctx *c, double *states, double *weights, double *cum_weights, int *max_wt_id){
	if(!c) return fail(MTFB_ERR_INVALID_ARG, "mtfb_pf_get_particles: null context");
	if(!c->pf_configured || !c->initialized) return fail(MTFB_ERR_LOGIC, "mtfb_pf_get_particles: configure and initialize first");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	const size_t pn = (size_t)c->P*c->pfp.n_particles;
	if(states) CUDA_TRY(cudaMemcpyAsync(states, c->pf.states, pn*c->S*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	if(weights) CUDA_TRY(cudaMemcpyAsync(weights, c->pf.weights, pn*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	if(cum_weights) CUDA_TRY(cudaMemcpyAsync(cum_weights, c->pf.cum_weights, pn*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	if(max_wt_id) CUDA_TRY(cudaMemcpyAsync(max_wt_id, c->pf.max_wt_id, (size_t)c->P*sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return MTFB_OK;
}

// ------------------------------------------------------------------------------------------------ robust warp estimation
void mtfb_est_default_params(mtfb_est_params *p){
	if(!p) return;
	// SSMEstimatorParams.cc:5-13
	p->method = MTFB_EST_RANSAC; p->ransac_reproj_thresh = 10.0; p->n_model_pts = 4; p->refine = 1; p->max_iters = 2000;
	p->max_subset_attempts = 300; p->confidence = 0.995; p->lm_max_iters = 10; p->seed = 0;
}

static mtfb_status est_reserve(mtfb_ctx *c, size_t n){
	if(n <= c->est_capacity && c->d_est_out) return MTFB_OK;
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	cudaFree(c->d_est_pts); cudaFree(c->d_est_mask); cudaFree(c->d_est_err); c->d_est_pts = nullptr; c->d_est_mask = nullptr; c->d_est_err = nullptr;
	c->est_capacity = 0;
	CUDA_TRY(cudaMalloc(&c->d_est_pts, 4 * n*sizeof(float)));
	CUDA_TRY(cudaMalloc(&c->d_est_mask, n));
	CUDA_TRY(cudaMalloc(&c->d_est_err, (size_t)EST_WARPS*n*sizeof(float)));
	if(!c->d_est_out) CUDA_TRY(cudaMalloc(&c->d_est_out, 48 * sizeof(double)));   // 32 outputs | the grid's region (mtfb_grid_advance)
	c->est_capacity = n;
	return MTFB_OK;
}

static mtfb_status est_fetch(mtfb_ctx *c, int ssm, int n, double *state_update, unsigned char *mask, double *warp, int *info);
static mtfb_status est_run(mtfb_ctx *c, const char *who, int ssm, const float *d_in, const float *d_out, int n, const mtfb_est_params *ep,
	double *state_update, unsigned char *mask, double *warp, int *info, bool launch_only = false){
	if(ssm != MTFB_SSM_HOMOGRAPHY && ssm != MTFB_SSM_AFFINE)
		return fail(MTFB_ERR_NOT_SUPPORTED, "%s: the estimators of Homography and Affine are implemented (ssm = %d)", who, ssm);
	if(ep->method != MTFB_EST_RANSAC && ep->method != MTFB_EST_LMEDS && ep->method != MTFB_EST_LEAST_SQUARES)
		return fail(MTFB_ERR_INVALID_ARG, "%s: invalid estimation method %d", who, ep->method);
	const int min_pts = ssm == MTFB_SSM_HOMOGRAPHY ? 4 : 3;           // the estimators' assert(_modelPoints >= 4 / 3)
	if(ep->n_model_pts < min_pts) return fail(MTFB_ERR_INVALID_ARG, "%s: n_model_pts = %d, at least %d needed", who, ep->n_model_pts, min_pts);
	if(ep->n_model_pts > EST_MAX_MODEL_PTS) return fail(MTFB_ERR_NOT_SUPPORTED, "%s: n_model_pts = %d, at most %d implemented", who, ep->n_model_pts, (int)EST_MAX_MODEL_PTS);
	if(n < ep->n_model_pts) return fail(MTFB_ERR_INVALID_ARG, "%s: %d points, n_model_pts = %d (CV_Assert(n_pts >= params.n_model_pts))", who, n, ep->n_model_pts);
	if(ep->max_iters < 1 || ep->max_subset_attempts < 1) return fail(MTFB_ERR_INVALID_ARG, "%s: max_iters and max_subset_attempts must be positive", who);
	EstDev e;
	e.in_pts = d_in; e.out_pts = d_out; e.n = n; e.homography = ssm == MTFB_SSM_HOMOGRAPHY;
	e.method = ep->method; e.model_points = ep->n_model_pts; e.refine = ep->refine; e.max_iters = ep->max_iters;
	e.max_attempts = ep->max_subset_attempts; e.lm_max_iters = ep->lm_max_iters;
	e.thresh = ep->ransac_reproj_thresh > 0 ? ep->ransac_reproj_thresh : 3.0;
	e.confidence = ep->confidence; e.seed = ep->seed;
	e.out = c->d_est_out; e.mask = c->d_est_mask; e.err = c->d_est_err;
	CUDA_TRY(launch_estimate(e, c->stream));
	++c->launches;
	if(launch_only) return MTFB_OK;
	return est_fetch(c, ssm, n, state_update, mask, warp, info);
}
static mtfb_status est_fetch(mtfb_ctx *c, int ssm, int n, double *state_update, unsigned char *mask, double *warp, int *info){
	double out[32];
	CUDA_TRY(cudaMemcpyAsync(out, c->d_est_out, sizeof(out), cudaMemcpyDeviceToHost, c->stream));
	if(mask) CUDA_TRY(cudaMemcpyAsync(mask, c->d_est_mask, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	if(warp) std::memcpy(warp, out, 9 * sizeof(double));
	if(state_update) std::memcpy(state_update, out + 9, (ssm == MTFB_SSM_HOMOGRAPHY ? 8 : 6)*sizeof(double));
	if(info) for(int i = 0; i < 4; ++i) info[i] = (int)out[17 + i];
#ifdef EST_PROF
	if(std::getenv("MTFB_EST_PROF")){
		std::fprintf(stderr, "est_prof cycles: draw %.0f  round %.0f  fit_inliers %.0f  lm %.0f (ctl+step %.0f  points %.0f  points+sum %.0f)  kernel %.0f\n",
			out[21], out[22], out[23], out[24] - out[23], out[25], out[26], out[27], out[31]);
	}
#endif
	return MTFB_OK;
}

mtfb_status mtfb_estimate_warp_from_pts(mtfb_ctx *c, int ssm, const float *in_pts, const float *out_pts, int n, const mtfb_est_params *ep,
	double *state_update, unsigned char *mask, double *warp, int *info){
	if(!c || !in_pts || !out_pts || !ep) return fail(MTFB_ERR_INVALID_ARG, "mtfb_estimate_warp_from_pts: null argument");
	if(n < 1) return fail(MTFB_ERR_INVALID_ARG, "mtfb_estimate_warp_from_pts: no points");
	for(size_t i = 0; i < 2 * (size_t)n; ++i)
		if(!std::isfinite(in_pts[i]) || !std::isfinite(out_pts[i])) return fail(MTFB_ERR_INVALID_ARG, "mtfb_estimate_warp_from_pts: non-finite point %zu", i / 2);
	CUDA_TRY(cudaSetDevice(c->prm.device));
	{ mtfb_status st = est_reserve(c, (size_t)n); if(st != MTFB_OK) return st; }
	CUDA_TRY(cudaMemcpyAsync(c->d_est_pts, in_pts, 2 * (size_t)n*sizeof(float), cudaMemcpyHostToDevice, c->stream));
	CUDA_TRY(cudaMemcpyAsync(c->d_est_pts + 2 * (size_t)n, out_pts, 2 * (size_t)n*sizeof(float), cudaMemcpyHostToDevice, c->stream));
	return est_run(c, "mtfb_estimate_warp_from_pts", ssm, c->d_est_pts, c->d_est_pts + 2 * (size_t)n, n, ep, state_update, mask, warp, info);
}

mtfb_status mtfb_estimate_warp_from_corners_device(mtfb_ctx *c, int ssm, const double *d_in_corners, const double *d_out_corners, int n,
	const mtfb_est_params *ep, double *state_update, unsigned char *mask, double *warp, int *info){
	if(!c || !d_in_corners || !d_out_corners || !ep) return fail(MTFB_ERR_INVALID_ARG, "mtfb_estimate_warp_from_corners_device: null argument");
	if(n < 1) return fail(MTFB_ERR_INVALID_ARG, "mtfb_estimate_warp_from_corners_device: no points");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	{ mtfb_status st = est_reserve(c, (size_t)n); if(st != MTFB_OK) return st; }
	CUDA_TRY(launch_centroids(d_in_corners, n, c->d_est_pts, c->stream));
	CUDA_TRY(launch_centroids(d_out_corners, n, c->d_est_pts + 2 * (size_t)n, c->stream));
	c->launches += 2;
	return est_run(c, "mtfb_estimate_warp_from_corners_device", ssm, c->d_est_pts, c->d_est_pts + 2 * (size_t)n, n, ep, state_update, mask, warp, info);
}

mtfb_status mtfb_grid_enable(mtfb_ctx *c){
	if(!c) return fail(MTFB_ERR_INVALID_ARG, "mtfb_grid_enable: null context");
	if(c->grid_enabled) return MTFB_OK;
	CUDA_TRY(cudaSetDevice(c->prm.device));
	CUDA_TRY(cudaMalloc(&c->d_grid_prev, 2 * (size_t)c->P*sizeof(float)));
	CUDA_TRY(cudaMalloc(&c->d_grid_curr, 2 * (size_t)c->P*sizeof(float)));
	c->grid_enabled = true;
	if(c->initialized){
		CUDA_TRY(launch_centroids(c->b.corners, c->P, c->d_grid_prev, c->stream));
		++c->launches;
	}
	return MTFB_OK;
}

mtfb_status mtfb_grid_estimate(mtfb_ctx *c, int ssm, const mtfb_est_params *ep, double *state_update, unsigned char *mask, double *warp, int *info){
	if(!c || !ep) return fail(MTFB_ERR_INVALID_ARG, "mtfb_grid_estimate: null argument");
	if(!c->grid_enabled) return fail(MTFB_ERR_LOGIC, "mtfb_grid_estimate: mtfb_grid_enable has not been called");
	if(!c->initialized) return fail(MTFB_ERR_LOGIC, "mtfb_grid_estimate: initialize has not been called");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	{ mtfb_status st = est_reserve(c, (size_t)c->P); if(st != MTFB_OK) return st; }
	CUDA_TRY(launch_centroids(c->b.corners, c->P, c->d_grid_curr, c->stream));
	++c->launches;
	return est_run(c, "mtfb_grid_estimate", ssm, c->d_grid_prev, c->d_grid_curr, c->P, ep, state_update, mask, warp, info);
}

mtfb_status mtfb_grid_initialize(mtfb_ctx *c, int grid_size_x, int grid_size_y, double patch_size_x, double patch_size_y, const double *region){
	if(!c || !region) return fail(MTFB_ERR_INVALID_ARG, "mtfb_grid_initialize: null argument");
	if(!c->grid_enabled) return fail(MTFB_ERR_LOGIC, "mtfb_grid_initialize: mtfb_grid_enable has not been called");
	{ mtfb_status st0 = adopt_prefetched(c); if(st0 != MTFB_OK) return st0; }
	if(!c->have_image) return fail(MTFB_ERR_LOGIC, "mtfb_grid_initialize: setImage has not been called");
	if(grid_size_x < 1 || grid_size_y < 1 || grid_size_x*grid_size_y != c->P)
		return fail(MTFB_ERR_INVALID_ARG, "mtfb_grid_initialize: %d x %d cells, the batch has %d patches", grid_size_x, grid_size_y, c->P);
	if(!(patch_size_x > 0) || !(patch_size_y > 0)) return fail(MTFB_ERR_INVALID_ARG, "mtfb_grid_initialize: patch sizes must be positive");
	for(int i = 0; i < 8; ++i) if(!std::isfinite(region[i])) return fail(MTFB_ERR_INVALID_ARG, "mtfb_grid_initialize: non-finite region corner");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	{ mtfb_status st = est_reserve(c, (size_t)c->P); if(st != MTFB_OK) return st; }
	c->all_parallelograms = true;                        // the cells are axis-aligned boxes
	// GridTracker::initialize (GridTracker.cc:232-245): ssm.initialize(corners); resetTrackers(true) -- the layout kernel with a
	// zero state update (the identity warp) leaves the region as it is
	double *d_region = c->d_est_out + 32, *d_zero = c->d_est_out + 40;
	CUDA_TRY(cudaMemcpyAsync(d_region, region, 8 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	CUDA_TRY(cudaMemsetAsync(d_zero, 0, 8 * sizeof(double), c->stream));
	CUDA_TRY(launch_grid_layout(1, d_zero, d_region, grid_size_x, grid_size_y, patch_size_x, patch_size_y, c->d_corners_in, c->stream));
	++c->launches;
	CUDA_TRY(launch_init(c->prm, c->threads, c->b, c->d_corners_in, c->d_mi_tab, c->stream));
	++c->launches;
	{ mtfb_status st1 = mark_frame_read(c); if(st1 != MTFB_OK) return st1; }
	c->initialized = true;
	CUDA_TRY(launch_centroids(c->b.corners, c->P, c->d_grid_prev, c->stream));
	++c->launches;
	return MTFB_OK;
}

mtfb_status mtfb_grid_advance(mtfb_ctx *c, int ssm, const mtfb_est_params *ep, int grid_size_x, int grid_size_y, double patch_size_x,
	double patch_size_y, double *region, double *state_update, unsigned char *mask, double *warp, int *info){
	if(!c || !ep || !region) return fail(MTFB_ERR_INVALID_ARG, "mtfb_grid_advance: null argument");
	if(!c->grid_enabled) return fail(MTFB_ERR_LOGIC, "mtfb_grid_advance: mtfb_grid_enable has not been called");
	if(!c->initialized) return fail(MTFB_ERR_LOGIC, "mtfb_grid_advance: initialize has not been called");
	if(grid_size_x < 1 || grid_size_y < 1 || grid_size_x*grid_size_y != c->P)
		return fail(MTFB_ERR_INVALID_ARG, "mtfb_grid_advance: %d x %d cells, the batch has %d patches", grid_size_x, grid_size_y, c->P);
	if(!(patch_size_x > 0) || !(patch_size_y > 0)) return fail(MTFB_ERR_INVALID_ARG, "mtfb_grid_advance: patch sizes must be positive");
	for(int i = 0; i < 8; ++i) if(!std::isfinite(region[i])) return fail(MTFB_ERR_INVALID_ARG, "mtfb_grid_advance: non-finite region corner");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	{ mtfb_status st = est_reserve(c, (size_t)c->P); if(st != MTFB_OK) return st; }
	// the region rides behind the estimator's 32 output doubles
	double *d_region = c->d_est_out + 32;
	CUDA_TRY(cudaMemcpyAsync(d_region, region, 8 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	CUDA_TRY(launch_centroids(c->b.corners, c->P, c->d_grid_curr, c->stream));
	++c->launches;
	{ mtfb_status st = est_run(c, "mtfb_grid_advance", ssm, c->d_grid_prev, c->d_grid_curr, c->P, ep, nullptr, nullptr, nullptr, nullptr, true);
	  if(st != MTFB_OK) return st; }
	CUDA_TRY(launch_grid_layout(ssm == MTFB_SSM_HOMOGRAPHY, c->d_est_out + 9, d_region, grid_size_x, grid_size_y, patch_size_x, patch_size_y,
		c->d_corners_in, c->stream));
	++c->launches;
	// GridTracker::resetTrackers(reinit = true): every cell re-initialised on the current frame at its new region
	c->all_parallelograms = true;                        // the cells are axis-aligned boxes
	CUDA_TRY(launch_init(c->prm, c->threads, c->b, c->d_corners_in, c->d_mi_tab, c->stream));
	++c->launches;
	{ mtfb_status st1 = mark_frame_read(c); if(st1 != MTFB_OK) return st1; }
	CUDA_TRY(launch_centroids(c->b.corners, c->P, c->d_grid_prev, c->stream));
	++c->launches;
	CUDA_TRY(cudaMemcpyAsync(region, d_region, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	return est_fetch(c, ssm, c->P, state_update, mask, warp, info);
}

mtfb_status mtfb_grid_commit(mtfb_ctx *c){
	if(!c) return fail(MTFB_ERR_INVALID_ARG, "mtfb_grid_commit: null context");
	if(!c->grid_enabled || !c->initialized) return fail(MTFB_ERR_LOGIC, "mtfb_grid_commit: mtfb_grid_enable and initialize first");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	CUDA_TRY(launch_centroids(c->b.corners, c->P, c->d_grid_prev, c->stream));
	++c->launches;
	return MTFB_OK;
}

mtfb_status mtfb_grid_get_pts(mtfb_ctx *c, float *prev_pts, float *curr_pts){
	if(!c) return fail(MTFB_ERR_INVALID_ARG, "mtfb_grid_get_pts: null context");
	if(!c->grid_enabled || !c->initialized) return fail(MTFB_ERR_LOGIC, "mtfb_grid_get_pts: mtfb_grid_enable and initialize first");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	if(prev_pts) CUDA_TRY(cudaMemcpyAsync(prev_pts, c->d_grid_prev, 2 * (size_t)c->P*sizeof(float), cudaMemcpyDeviceToHost, c->stream));
	if(curr_pts) CUDA_TRY(cudaMemcpyAsync(curr_pts, c->d_grid_curr, 2 * (size_t)c->P*sizeof(float), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return MTFB_OK;
}

mtfb_status mtfb_update(mtfb_ctx *c){
	if(!c) return fail(MTFB_ERR_INVALID_ARG, "mtfb_update: null context");
	if(!c->initialized) return fail(MTFB_ERR_LOGIC, "mtfb_update: initialize has not been called");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	{ mtfb_status st0 = adopt_prefetched(c); if(st0 != MTFB_OK) return st0; }
	if(c->prm.sm == MTFB_SM_PF){
		if(!c->pf_configured) return fail(MTFB_ERR_LOGIC, "mtfb_update: a PF context needs mtfb_pf_configure (before mtfb_initialize) to "
			"track; without it it only evaluates particles (mtfb_pf_evaluate)");
		mtfb_status st = pf_update(c);
		return st != MTFB_OK ? st : mark_frame_read(c);
	}
	if(c->b.log) CUDA_TRY(cudaMemsetAsync(c->b.log, 0, sizeof(mtfb_iter_log)*(size_t)c->P*c->b.log_slots, c->stream));
	if(c->peer_attached){
		// the frame's all-gather rides on the update kernel: every patch's final corners go straight into the gathered
		// array (of this frame's parity) of every rank; mtfb_peer_gather then only signals and waits
		if(c->peer_pending) return fail(MTFB_ERR_LOGIC, "mtfb_update: the previous frame's corners have not been gathered "
			"(mtfb_peer_gather): the ranks of a job exchange every frame");
		DevBatch b = c->b;
		peer_out_for(c, c->peer_seq + 1, b.peers);
		CUDA_TRY(launch_update(c->prm, c->threads, c->occ, b, c->d_mi_tab, c->stream, mom_work_for(c), c->mom_threads, frame_map_for(c)));
		++c->peer_seq; c->peer_pending = true;
	} else
		CUDA_TRY(launch_update(c->prm, c->threads, c->occ, c->b, c->d_mi_tab, c->stream, mom_work_for(c), c->mom_threads, frame_map_for(c)));
	++c->launches;
	return mark_frame_read(c);
}

static mtfb_status ensure_scratch(mtfb_ctx *c, size_t bytes){
	if(bytes <= c->scratch_bytes) return MTFB_OK;
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	cudaFree(c->d_scratch); c->d_scratch = nullptr; c->scratch_bytes = 0;
	CUDA_TRY(cudaMalloc(&c->d_scratch, bytes));
	c->scratch_bytes = bytes;
	return MTFB_OK;
}

static mtfb_status d2h(mtfb_ctx *c, void *dst, const void *src, size_t bytes){
	CUDA_TRY(cudaSetDevice(c->prm.device));
	CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return MTFB_OK;
}

mtfb_status mtfb_enable_iter_log(mtfb_ctx *c, int n_slots){
	if(!c || n_slots < 0) return fail(MTFB_ERR_INVALID_ARG, "mtfb_enable_iter_log: bad argument");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	cudaFree(c->d_log); c->d_log = nullptr; c->b.log = nullptr; c->b.log_slots = 0;
	if(n_slots > 0){
		CUDA_TRY(cudaMalloc(&c->d_log, sizeof(mtfb_iter_log)*(size_t)c->P*n_slots));
		CUDA_TRY(cudaMemset(c->d_log, 0, sizeof(mtfb_iter_log)*(size_t)c->P*n_slots));
		c->b.log = c->d_log; c->b.log_slots = n_slots;
	}
	return MTFB_OK;
}

mtfb_status mtfb_get_iter_log(mtfb_ctx *c, mtfb_iter_log *out){
	if(!c || !out) return fail(MTFB_ERR_INVALID_ARG, "mtfb_get_iter_log: null argument");
	if(!c->b.log) return fail(MTFB_ERR_LOGIC, "mtfb_get_iter_log: no iteration log attached");
	return d2h(c, out, c->b.log, sizeof(mtfb_iter_log)*(size_t)c->P*c->b.log_slots);
}

mtfb_status mtfb_iterate_once(mtfb_ctx *c, double *jacobian, double *hessian, double *similarity, double *state_update){
	if(!c) return fail(MTFB_ERR_INVALID_ARG, "mtfb_iterate_once: null context");
	if(!c->initialized) return fail(MTFB_ERR_LOGIC, "mtfb_iterate_once: initialize has not been called");
	if(c->prm.sm == MTFB_SM_PF) return fail(MTFB_ERR_NOT_SUPPORTED, "mtfb_iterate_once: not a Gauss-Newton context");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	const int S = c->S, P = c->P;
	mtfb_status st = ensure_scratch(c, sizeof(mtfb_iter_log)*(size_t)P);
	if(st != MTFB_OK) return st;
	DevBatch b = c->b;
	b.max_iters = 1; b.epsilon = -1;               // one pass, no early exit bookkeeping differences
	b.log = reinterpret_cast<mtfb_iter_log*>(c->d_scratch); b.log_slots = 1;
	CUDA_TRY(cudaMemsetAsync(b.log, 0, sizeof(mtfb_iter_log)*(size_t)P, c->stream));
	CUDA_TRY(launch_update(c->prm, c->threads, c->occ, b, c->d_mi_tab, c->stream, mom_work_for(c), c->mom_threads, frame_map_for(c)));
	++c->launches;
	std::vector<mtfb_iter_log> host(P);
	st = d2h(c, host.data(), b.log, sizeof(mtfb_iter_log)*(size_t)P);
	if(st != MTFB_OK) return st;
	for(int p = 0; p < P; ++p){
		if(jacobian) std::memcpy(jacobian + (size_t)p*S, host[p].jacobian, S*sizeof(double));
		if(hessian) std::memcpy(hessian + (size_t)p*S*S, host[p].hessian, S*S*sizeof(double));
		if(similarity) similarity[p] = host[p].f;
		if(state_update) std::memcpy(state_update + (size_t)p*S, host[p].state_update, S*sizeof(double));
	}
	return MTFB_OK;
}

mtfb_status mtfb_pf_evaluate_device(mtfb_ctx *c, const double *d_states, int n_particles, double *d_likelihood, double *d_similarity){
	if(!c || !d_states || n_particles < 1) return fail(MTFB_ERR_INVALID_ARG, "mtfb_pf_evaluate_device: bad argument");
	if(!c->initialized) return fail(MTFB_ERR_LOGIC, "mtfb_pf_evaluate_device: initialize has not been called");

	CUDA_TRY(cudaSetDevice(c->prm.device));
	if(c->prm.am == MTFB_AM_MI)
		CUDA_TRY(launch_pf_evaluate_mi(c->prm.ssm, c->b, c->prm.mi_n_bins, c->prm.mi_pre_seed, c->d_mi_tab, d_states, n_particles,
			d_likelihood, d_similarity, c->prm.likelihood_alpha, c->stream));
	else if(c->prm.precision == MTFB_PRECISION_F32)
		CUDA_TRY(launch_pf_evaluate_f32(c->prm.ssm, c->b, d_states, n_particles, d_likelihood, d_similarity, c->prm.likelihood_alpha, c->stream));
	else
		CUDA_TRY(launch_pf_evaluate(c->prm.am, c->prm.ssm, c->b, d_states, n_particles, d_likelihood, d_similarity,
			c->prm.likelihood_alpha, c->stream));
	++c->launches;
	return MTFB_OK;
}

mtfb_status mtfb_pf_evaluate(mtfb_ctx *c, const double *states, int n_particles, double *likelihood, double *similarity){
	if(!c || !states || n_particles < 1) return fail(MTFB_ERR_INVALID_ARG, "mtfb_pf_evaluate: bad argument");
	if(!c->initialized) return fail(MTFB_ERR_LOGIC, "mtfb_pf_evaluate: initialize has not been called");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	const size_t n = (size_t)c->P*n_particles;
	mtfb_status st = ensure_scratch(c, (n*c->S + 2 * n)*sizeof(double));
	if(st != MTFB_OK) return st;
	double *d_states = c->d_scratch, *d_lik = d_states + n*c->S, *d_sim = d_lik + n;
	CUDA_TRY(cudaMemcpyAsync(d_states, states, n*c->S*sizeof(double), cudaMemcpyHostToDevice, c->stream));
	st = mtfb_pf_evaluate_device(c, d_states, n_particles, d_lik, d_sim);
	if(st != MTFB_OK) return st;
	if(likelihood) CUDA_TRY(cudaMemcpyAsync(likelihood, d_lik, n*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	if(similarity) CUDA_TRY(cudaMemcpyAsync(similarity, d_sim, n*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return MTFB_OK;
}

#define GETTER_PRELUDE(name) \
	if(!c || !out) return fail(MTFB_ERR_INVALID_ARG, name ": null argument"); \
	if(!c->initialized) return fail(MTFB_ERR_LOGIC, name ": initialize has not been called");

mtfb_status mtfb_get_corners(mtfb_ctx *c, double *out){ GETTER_PRELUDE("mtfb_get_corners") return d2h(c, out, c->b.corners, 8 * (size_t)c->P*sizeof(double)); }
mtfb_status mtfb_get_state(mtfb_ctx *c, double *out){ GETTER_PRELUDE("mtfb_get_state") return d2h(c, out, c->b.state, (size_t)c->S*c->P*sizeof(double)); }
mtfb_status mtfb_get_n_iters(mtfb_ctx *c, int *out){ GETTER_PRELUDE("mtfb_get_n_iters") return d2h(c, out, c->b.n_iters, (size_t)c->P*sizeof(int)); }
mtfb_status mtfb_get_similarity(mtfb_ctx *c, double *out){ GETTER_PRELUDE("mtfb_get_similarity") return d2h(c, out, c->b.f, (size_t)c->P*sizeof(double)); }
mtfb_status mtfb_get_patch_status(mtfb_ctx *c, int *out){ GETTER_PRELUDE("mtfb_get_patch_status") return d2h(c, out, c->b.status, (size_t)c->P*sizeof(int)); }
mtfb_status mtfb_get_init_warp(mtfb_ctx *c, double *out){ GETTER_PRELUDE("mtfb_get_init_warp") return d2h(c, out, c->b.dlt, 9 * (size_t)c->P*sizeof(double)); }
mtfb_status mtfb_get_init_pix_vals(mtfb_ctx *c, double *out){ GETTER_PRELUDE("mtfb_get_init_pix_vals") return d2h(c, out, c->b.I0, (size_t)c->N*c->P*sizeof(double)); }

mtfb_status mtfb_get_curr_stage(mtfb_ctx *c, double *pts, double *pix_vals, double *pix_grad, double *pix_jac){
	if(!c) return fail(MTFB_ERR_INVALID_ARG, "mtfb_get_curr_stage: null context");
	if(!c->initialized) return fail(MTFB_ERR_LOGIC, "mtfb_get_curr_stage: initialize has not been called");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	const size_t N = c->N, P = c->P, S = c->S;
	const size_t n_pts = pts ? 2 * N*P : 0, n_val = pix_vals ? N*P : 0, n_grad = pix_grad ? 2 * N*P : 0, n_jac = pix_jac ? S*N*P : 0;
	mtfb_status st = ensure_scratch(c, (n_pts + n_val + n_grad + n_jac + 1)*sizeof(double));
	if(st != MTFB_OK) return st;
	StageTaps t;
	double *q = c->d_scratch;
	t.pts = pts ? q : nullptr; q += n_pts;
	t.pix_vals = pix_vals ? q : nullptr; q += n_val;
	t.pix_grad = pix_grad ? q : nullptr; q += n_grad;
	t.pix_jac = pix_jac ? q : nullptr;
	CUDA_TRY(launch_stage(c->prm.ssm, c->threads, c->b, t, c->stream));
	++c->launches;
	if(pts) CUDA_TRY(cudaMemcpyAsync(pts, t.pts, n_pts*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	if(pix_vals) CUDA_TRY(cudaMemcpyAsync(pix_vals, t.pix_vals, n_val*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	if(pix_grad) CUDA_TRY(cudaMemcpyAsync(pix_grad, t.pix_grad, n_grad*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	if(pix_jac) CUDA_TRY(cudaMemcpyAsync(pix_jac, t.pix_jac, n_jac*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return MTFB_OK;
}

mtfb_status mtfb_get_curr_stage_f32(mtfb_ctx *c, int *idx, float *pix_vals, float *pix_grad, double *pix_jac, float *fast_err){
	if(!c) return fail(MTFB_ERR_INVALID_ARG, "mtfb_get_curr_stage_f32: null context");
	if(!c->initialized) return fail(MTFB_ERR_LOGIC, "mtfb_get_curr_stage_f32: initialize has not been called");
	if(c->prm.precision != MTFB_PRECISION_F32) return fail(MTFB_ERR_LOGIC, "mtfb_get_curr_stage_f32: context is not MTFB_PRECISION_F32");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	const size_t N = c->N, P = c->P, S = c->S;
	// 8-byte slots for everything: simple and aligned
	const size_t n_idx = idx ? N*P : 0, n_val = pix_vals ? (N*P + 1) / 2 : 0, n_grad = pix_grad ? N*P : 0, n_jac = pix_jac ? S*N*P : 0,
		n_err = fast_err ? (N*P + 1) / 2 : 0;
	mtfb_status st = ensure_scratch(c, (n_idx + n_val + n_grad + n_jac + n_err + 1)*sizeof(double));
	if(st != MTFB_OK) return st;
	StageTapsF32 t;
	double *q = c->d_scratch;
	t.idx = idx ? reinterpret_cast<int*>(q) : nullptr; q += n_idx;
	t.pix_vals = pix_vals ? reinterpret_cast<float*>(q) : nullptr; q += n_val;
	t.pix_grad = pix_grad ? reinterpret_cast<float*>(q) : nullptr; q += n_grad;
	t.fast_err = fast_err ? reinterpret_cast<float*>(q) : nullptr; q += n_err;
	t.pix_jac = pix_jac ? q : nullptr;
	CUDA_TRY(launch_stage_f32(c->prm.ssm, c->b, t, c->stream));
	++c->launches;
	if(idx) CUDA_TRY(cudaMemcpyAsync(idx, t.idx, 2 * N*P*sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	if(pix_vals) CUDA_TRY(cudaMemcpyAsync(pix_vals, t.pix_vals, N*P*sizeof(float), cudaMemcpyDeviceToHost, c->stream));
	if(pix_grad) CUDA_TRY(cudaMemcpyAsync(pix_grad, t.pix_grad, 2 * N*P*sizeof(float), cudaMemcpyDeviceToHost, c->stream));
	if(fast_err) CUDA_TRY(cudaMemcpyAsync(fast_err, t.fast_err, N*P*sizeof(float), cudaMemcpyDeviceToHost, c->stream));
	if(pix_jac) CUDA_TRY(cudaMemcpyAsync(pix_jac, t.pix_jac, n_jac*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return MTFB_OK;
}

mtfb_status mtfb_get_init_pts(mtfb_ctx *c, double *out){
	GETTER_PRELUDE("mtfb_get_init_pts")
	// init_pts are not stored: they are the points of the identity warp (recomputed from the DLT every pass)
	CUDA_TRY(cudaSetDevice(c->prm.device));
	const size_t N = c->N, P = c->P;
	mtfb_status st = ensure_scratch(c, (2 * N*P + 9 * P)*sizeof(double));
	if(st != MTFB_OK) return st;
	// temporarily evaluate the stage kernel with curr_warp = identity
	DevBatch b = c->b;
	double *d_pts = c->d_scratch, *d_eye = c->d_scratch + 2 * N*P;
	std::vector<double> eye(9 * P, 0.0);
	for(size_t p = 0; p < P; ++p){ eye[9 * p] = eye[9 * p + 4] = eye[9 * p + 8] = 1; }
	CUDA_TRY(cudaMemcpyAsync(d_eye, eye.data(), eye.size()*sizeof(double), cudaMemcpyHostToDevice, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	b.warp = d_eye;
	StageTaps t = { d_pts, nullptr, nullptr, nullptr };
	CUDA_TRY(launch_stage(c->prm.ssm, c->threads, b, t, c->stream));
	++c->launches;
	return d2h(c, out, d_pts, 2 * N*P*sizeof(double));
}

#if MTFB_PROF
extern "C" int mtfb_prof_read(mtfb_ctx *c, long long *out, int n){
	cudaStreamSynchronize(c->stream);
	return (int)cudaMemcpy(out, c->b.n_iters_prof, n*sizeof(long long), cudaMemcpyDeviceToHost);
}
extern "C" int mtfb_prof_reset(mtfb_ctx *c){
	cudaStreamSynchronize(c->stream);
	return (int)cudaMemset(c->b.n_iters_prof, 0, 64 * sizeof(long long));
}
#endif

mtfb_status mtfb_debug_colpiv_qr_solve(int device, int n, int fast, int n_sys, const double *A, const double *b, double *x,
	int *nonzero_pivots, int *perm){
	if(!A || !b || !x || n_sys < 1 || (n != 6 && n != 8)) return fail(MTFB_ERR_INVALID_ARG, "mtfb_debug_colpiv_qr_solve: bad argument");
	CUDA_TRY(cudaSetDevice(device));
	double *d = nullptr; int *di = nullptr;
	const size_t nA = (size_t)n_sys*n*n, nv = (size_t)n_sys*n;
	CUDA_TRY(cudaMalloc(&d, (nA + 2 * nv)*sizeof(double)));
	if(cudaMalloc(&di, (nv + n_sys)*sizeof(int)) != cudaSuccess){ cudaFree(d); return fail(MTFB_ERR_NO_MEMORY, "mtfb_debug_colpiv_qr_solve: out of device memory"); }
	cudaError_t e = cudaMemcpy(d, A, nA*sizeof(double), cudaMemcpyHostToDevice);
	if(e == cudaSuccess) e = cudaMemcpy(d + nA, b, nv*sizeof(double), cudaMemcpyHostToDevice);
	if(e == cudaSuccess) e = launch_debug_qr_solve(n, fast, n_sys, d, d + nA, d + nA + nv, di, di + n_sys, 0);
	if(e == cudaSuccess) e = cudaDeviceSynchronize();
	if(e == cudaSuccess) e = cudaMemcpy(x, d + nA + nv, nv*sizeof(double), cudaMemcpyDeviceToHost);
	if(e == cudaSuccess && nonzero_pivots) e = cudaMemcpy(nonzero_pivots, di, n_sys*sizeof(int), cudaMemcpyDeviceToHost);
	if(e == cudaSuccess && perm) e = cudaMemcpy(perm, di + n_sys, nv*sizeof(int), cudaMemcpyDeviceToHost);
	cudaFree(d); cudaFree(di);
	if(e != cudaSuccess) return fail(MTFB_ERR_CUDA, "mtfb_debug_colpiv_qr_solve: %s", cudaGetErrorString(e));
	return MTFB_OK;
}

// ------------------------------------------------------------------------------------------------ multi-GPU exchange
static size_t peer_flags_offset(int n_total){ return ((size_t)2 * n_total * 8 * sizeof(double) + 255) & ~(size_t)255; }

mtfb_status mtfb_peer_export(mtfb_ctx *c, int n_total, mtfb_peer_handle *out){
	if(!c || !out || n_total < 1) return fail(MTFB_ERR_INVALID_ARG, "mtfb_peer_export: bad argument");
	if(c->d_peer_buf) return fail(MTFB_ERR_LOGIC, "mtfb_peer_export: already exported");
	static_assert(sizeof(cudaIpcMemHandle_t) <= MTFB_PEER_HANDLE_BYTES, "handle size");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	const size_t bytes = peer_flags_offset(n_total) + 256 + 64;              // arrays | 64 flags | error word
	CUDA_TRY(cudaMalloc(&c->d_peer_buf, bytes));
	CUDA_TRY(cudaMemset(c->d_peer_buf, 0, bytes));
	CUDA_TRY(cudaDeviceSynchronize());
	cudaIpcMemHandle_t h;
	CUDA_TRY(cudaIpcGetMemHandle(&h, c->d_peer_buf));
	std::memset(out->bytes, 0, sizeof(out->bytes));
	std::memcpy(out->bytes, &h, sizeof(h));
	c->peer_n_total = n_total;
	return MTFB_OK;
}

mtfb_status mtfb_peer_attach(mtfb_ctx *c, int rank, int world, int row0, const mtfb_peer_handle *handles){
	if(!c || !handles || world < 1 || world > MTFB_MAX_PEERS || rank < 0 || rank >= world)
		return fail(MTFB_ERR_INVALID_ARG, "mtfb_peer_attach: 1 <= world <= %d, 0 <= rank < world required", (int)MTFB_MAX_PEERS);
	if(!c->d_peer_buf) return fail(MTFB_ERR_LOGIC, "mtfb_peer_attach: mtfb_peer_export has not been called");
	if(c->peer_attached) return fail(MTFB_ERR_LOGIC, "mtfb_peer_attach: already attached");
	if(row0 < 0 || row0 + c->P > c->peer_n_total) return fail(MTFB_ERR_INVALID_ARG, "mtfb_peer_attach: rows %d .. %d do not fit %d",
		row0, row0 + c->P, c->peer_n_total);
	CUDA_TRY(cudaSetDevice(c->prm.device));
	for(int r = 0; r < world; ++r){
		void *base = c->d_peer_buf;
		if(r != rank){
			cudaIpcMemHandle_t h;
			std::memcpy(&h, handles[r].bytes, sizeof(h));
			CUDA_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
			c->peer_mapped[r] = base;
		}
		c->peer_base[r] = static_cast<double*>(base);
		c->peer_flags[r] = reinterpret_cast<unsigned*>(static_cast<char*>(base) + peer_flags_offset(c->peer_n_total));
	}
	c->peer_rank = rank; c->peer_world = world; c->peer_row0 = row0; c->peer_seq = 0; c->peer_pending = false;
	c->peer_attached = true;
	return MTFB_OK;
}

mtfb_status mtfb_peer_gather(mtfb_ctx *c){
	if(!c || !c->peer_attached) return fail(MTFB_ERR_LOGIC, "mtfb_peer_gather: mtfb_peer_attach has not been called");
	CUDA_TRY(cudaSetDevice(c->prm.device));
	if(!c->peer_pending){
		// the corners come from something else than an update kernel (initialize, setRegion, the particle filter): push them
		++c->peer_seq;
		PeerOut po;
		peer_out_for(c, c->peer_seq, po);
		CUDA_TRY(launch_peer_push(c->b.corners, c->P, po, c->stream));
		++c->launches;
	}
	int *d_err = reinterpret_cast<int*>(reinterpret_cast<char*>(c->peer_flags[c->peer_rank]) + 256);
	CUDA_TRY(launch_peer_signal_wait(c->peer_flags, c->peer_flags[c->peer_rank], c->peer_rank, c->peer_world, c->peer_seq, d_err, c->stream));
	++c->launches;
	c->peer_pending = false;
	return MTFB_OK;
}

mtfb_status mtfb_peer_gathered(mtfb_ctx *c, const double **d_corners, int *n_total){
	if(!c || !c->peer_attached) return fail(MTFB_ERR_LOGIC, "mtfb_peer_gathered: mtfb_peer_attach has not been called");
	if(d_corners) *d_corners = c->peer_base[c->peer_rank] + (size_t)(c->peer_seq & 1u)*c->peer_n_total * 8;
	if(n_total) *n_total = c->peer_n_total;
	return MTFB_OK;
}

mtfb_status mtfb_get_gathered_region(mtfb_ctx *c, double *out){
	if(!c || !out || !c->peer_attached) return fail(MTFB_ERR_LOGIC, "mtfb_get_gathered_region: mtfb_peer_attach has not been called");
	if(c->peer_pending) return fail(MTFB_ERR_LOGIC, "mtfb_get_gathered_region: mtfb_peer_gather has not been called for this frame");
	const double *src = c->peer_base[c->peer_rank] + (size_t)(c->peer_seq & 1u)*c->peer_n_total * 8;
	int err = 0;
	const int *d_err = reinterpret_cast<const int*>(reinterpret_cast<const char*>(c->peer_flags[c->peer_rank]) + 256);
	CUDA_TRY(cudaSetDevice(c->prm.device));
	CUDA_TRY(cudaMemcpyAsync(out, src, (size_t)c->peer_n_total * 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	if(err) return fail(MTFB_ERR_CUDA, "mtfb_get_gathered_region: rank %d never signalled its frame (5 s)", err - 1);
	return MTFB_OK;
}

mtfb_status mtfb_device_results(mtfb_ctx *c, double **d_corners, double **d_state, int **d_n_iters){
	if(!c) return fail(MTFB_ERR_INVALID_ARG, "mtfb_device_results: null context");
	if(d_corners) *d_corners = c->b.corners;
	if(d_state) *d_state = c->b.state;
	if(d_n_iters) *d_n_iters = c->b.n_iters;
	return MTFB_OK;
}

} // extern "C"
