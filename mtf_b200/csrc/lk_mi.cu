// lk_mi.cu -- the fused Lucas-Kanade kernels for the MI appearance model (sm_100a).
//
// Reference (AM/src/MI.cc): mutual information with a cubic-B-spline Parzen joint histogram.
//   updateSimilarity :346-382   curr_hist (B), joint_hist (B x B) from 4 x 4 B-spline weights per pixel, logs,
//                               f = sum p_ij (log p_ij - log p_i - log q_j)
//   updateCurrGrad  :426-443    df_dIt(pix) = sum_{4x4} curr_hist_grad . init_hist_mat . (1 + log p_ij - log p_i)
//   updateInitGrad  :398-417    df_dI0(pix) likewise with the roles swapped
//   cmptSelfHessian :515-594, cmptSelfHist :639-658   (evaluated once, at initialize(): init_self_hessian)
//   B-splines: Utilities/include/mtf/Utilities/histUtils.h:206-224 (bSpl3WithGrad), :269-280 (bSpl3Hess)
// The reference materialises B x N weight matrices and a B^2 x N joint-gradient matrix (zero-filled every
// pass: 5 MB at N = 10^4); here the 4 weights of a pixel are recomputed from its value where they are used.
// One pass = sweep 1 (warp + sample, values kept in shared memory, histograms built in per-lane private
// shared-memory copies), a 72-entry log table, sweep 2 (per-pixel gradient weight from a 4 x 4 table gather, pixel
// Jacobian row, S or 2S sums).  Implemented Hessian: InitialSelf (the ICLK default, ICLKParams.cc:6).
// The per-pass histograms are accumulated in per-lane private copies and folded in a fixed order: deterministic.
// (Only initialize() uses shared-memory atomics, for the B^2 x S joint-histogram Jacobian of the self Hessian.)
#include <cstdlib>
#include "lk_solve.cuh"

namespace mtfb {

constexpr int MI_BMAX = 16;
constexpr int MI_HBMAX = MI_BMAX + MI_BMAX*MI_BMAX;

// histUtils.h:206-224
__device__ __forceinline__ void bspl3_with_grad(double &val, double &diff, double x){
	val = 0; diff = 0;
	if((x > -2) && (x <= -1)){
		double temp = 2 + x; diff = (temp * temp) / 2; val = (diff * temp) / 3;
	} else if((x > -1) && (x <= 0)){
		double temp = x / 2; val = (2.0 / 3.0) - x*x*(1 + temp); diff = -x * (temp + x + 2);
	} else if((x > 0) && (x <= 1)){
		double temp = x / 2; val = (2.0 / 3.0) - x*x*(1 - temp); diff = x * (temp + x - 2);
	} else if((x > 1) && (x < 2)){
		double temp = 2 - x; diff = -(temp * temp) / 2; val = -(diff * temp) / 3;
	}
}
// histUtils.h:269-280
__device__ __forceinline__ double bspl3_hess(double x){
	if((x > -2) && (x <= -1)) return 2 + x;
	if((x > -1) && (x <= 0)) return -(3 * x + 2);
	if((x > 0) && (x <= 1)) return 3 * x - 2;
	if((x > 1) && (x < 2)) return 2 - x;
	return 0;
}

// The four branches of bSpl3WithGrad (histUtils.h:206-224) one by one, in the reference's operation order; the division by 3
// is the exact quotient from RN(1/3) and one correction (lk_math.cuh div_by) instead of an IEEE division sequence.
__device__ __forceinline__ void bspl3_piece0(double &val, double &diff, double x){      // -2 < x <= -1
	const double temp = 2 + x; diff = (temp * temp) * 0.5; val = div_by(diff * temp, 3.0, 0.33333333333333331);
}
__device__ __forceinline__ void bspl3_piece1(double &val, double &diff, double x){      // -1 < x <= 0
	const double temp = x * 0.5; val = (2.0 / 3.0) - x*x*(1 + temp); diff = -x * (temp + x + 2);
}
__device__ __forceinline__ void bspl3_piece2(double &val, double &diff, double x){      //  0 < x <= 1
	const double temp = x * 0.5; val = (2.0 / 3.0) - x*x*(1 - temp); diff = x * (temp + x - 2);
}
__device__ __forceinline__ void bspl3_piece3(double &val, double &diff, double x){      //  1 < x < 2 (x = 2: both zero)
	const double temp = 2 - x; diff = -(temp * temp) * 0.5; val = -div_by(diff * temp, 3.0, 0.33333333333333331);
}

// the (up to) 4 histogram bins a pixel value touches and its B-spline weights / derivatives (MI.cc:231-243).
// With bin = (int)v and lo = bin - 1 the k-th argument lo - v + k lies in the k-th branch's interval, so the four weights are
// the four pieces in order (one when lo was clamped at 0: the pieces shift by one) -- no interval tests per weight.  Values
// outside [0, B - 1) (not produced by pix_norm_mult on 8-bit frames) take the generic function.
struct BinWeights { int lo, hi; double w[4], d[4]; };
__device__ __forceinline__ BinWeights bin_weights(double v, int B){
	BinWeights o;
	const int bin = static_cast<int>(v);
	o.lo = bin - 1 > 0 ? bin - 1 : 0;
	o.hi = bin + 2 < B - 1 ? bin + 2 : B - 1;
	double x = o.lo - v;
	if(v >= 0.0 && bin <= B - 1){
		const double x1 = x + 1, x2 = x1 + 1, x3 = x2 + 1;              // curr_diff incremented bin by bin, as the reference does
		if(o.lo == bin - 1){
			bspl3_piece0(o.w[0], o.d[0], x); bspl3_piece1(o.w[1], o.d[1], x1);
			bspl3_piece2(o.w[2], o.d[2], x2); bspl3_piece3(o.w[3], o.d[3], x3);
		} else{                                                         // bin = 0: lo = 0 = bin
			bspl3_piece1(o.w[0], o.d[0], x); bspl3_piece2(o.w[1], o.d[1], x1);
			bspl3_piece3(o.w[2], o.d[2], x2); o.w[3] = 0; o.d[3] = 0;
		}
#pragma unroll
		for(int k = 1; k < 4; ++k) if(o.lo + k > o.hi){ o.w[k] = 0; o.d[k] = 0; }
		return o;
	}
#pragma unroll
	for(int k = 0; k < 4; ++k){
		bspl3_with_grad(o.w[k], o.d[k], x);
		if(o.lo + k > o.hi){ o.w[k] = 0; o.d[k] = 0; }
		x += 1;
	}
	return o;
}

struct MiParams { int B; double pre_seed, hist_pre_seed, hist_norm_mult; int copies; };

// per-template table kept in global memory (P x MI_TAB doubles): [0..B) init_hist, [16..16+B) init_hist_log
constexpr int MI_TAB = 32;

// REINIT: setRegion() of the search methods that keep template Jacobians -- nt::ESM::setRegion (NT/ESM.cc:150-168), nt::FCLK::setRegion
// with InitialSelf (NT/FCLK.cc:360-376): ssm.setCorners, init_pix_jacobian = cmptInitPixJacobian(am.getInitPixGrad()) at the NEW
// template points, init_self_hessian = MI::cmptSelfHessian(init_pix_jacobian) (MI.cc:515-594), which reads the appearance model's
// CURRENT state: curr_hist and the self joint histogram of the pixel values of the LAST pass of the last update (the update kernel
// leaves them in It_scratch; after initialize() they are the template's).  The template values, init_hist and f are kept; the
// template gradient becomes the un-chained one (G0raw).
template<int SSM, int T, bool REINIT = false>
__global__ void __launch_bounds__(T) mi_init_kernel(DevBatch b, const double *__restrict__ corners_in, MiParams mp, double *mi_tab){
	constexpr int S = StateSize<SSM>::value;
	constexpr int NH = S*(S + 1) / 2;
	const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int B = mp.B, N = b.N;
	__shared__ double s_dlt[9];
	__shared__ double s_hist[MI_BMAX], s_hist_log[MI_BMAX], s_joint[MI_BMAX*MI_BMAX], s_factor[MI_BMAX*MI_BMAX];
	__shared__ double s_jhj[MI_BMAX*MI_BMAX*S];
	__shared__ double s_part[(T / 32) * NH];
	__shared__ double s_sum[NH];
	const double *c_in = corners_in + (size_t)p * 8;
	if(warp == 0){
		Mat3 dlt = set_corners<SSM>(b, p, lane, c_in);
		if(lane < 9) s_dlt[lane] = dlt.m[lane];
		if(!REINIT && lane == 0){ b.n_iters[p] = 0; b.status[p] = 0; }
	}
	for(int i = tid; i < B; i += T) s_hist[i] = mp.hist_pre_seed;
	for(int i = tid; i < B*B; i += T) s_joint[i] = mp.pre_seed;
	for(int i = tid; i < B*B*S; i += T) s_jhj[i] = 0;
	cta_sync<T>();
	Mat3 dlt, W = mat3_identity();
#pragma unroll
	for(int i = 0; i < 9; ++i) dlt.m[i] = s_dlt[i];
	if(b.norm_init) W = dlt;                                 // Homography.cc:57-62: curr_warp starts as the DLT warp
	// (the affine chain rule's a, b, c, d = curr_state + identity, Affine.cc:220-223: the identity unless the start is the NDLT warp)
	const double abcd[4] = { (W.m[0] - 1) + 1, W.m[1], W.m[3], (W.m[4] - 1) + 1 };
	double *I0 = b.I0 + (size_t)p*N, *G0 = b.G0 + (size_t)p * 2 * N;
	// the pixel values the self Hessian is built on: the template's own at initialize(), the last pass's at setRegion()
	const double *Vh = REINIT ? b.It_scratch + (size_t)p*N : I0;
	// phase 1: template values (scaled to bin units, MI.cc:91-94), chained gradient, init_hist and the self joint histogram
	for(PixIter it(tid, T, b.resx); it.pix < N; it.next(T)){
		double val;
		if(REINIT){
			val = Vh[it.pix];
			G0[it.pix] = b.G0raw[(size_t)p * 2 * N + it.pix]; G0[N + it.pix] = b.G0raw[(size_t)p * 2 * N + N + it.pix];
		} else{
			PixGeom g = pixel_geometry<SSM>(dlt, W, b.xv[it.col], b.yv[it.row], b.norm_init != 0);
			Sample smp;
			pixel_value_and_gradient<SSM, false>(b, W, g, smp);
			val = b.pix_mult*smp.val + b.pix_add;
			double J[S];
			pixel_jacobian_row<SSM>(b, W, abcd, g, smp.gx, smp.gy, J);
			I0[it.pix] = val;
			G0[it.pix] = (SSM == SSM_HOM) ? J[2] : J[0];
			G0[N + it.pix] = (SSM == SSM_HOM) ? J[5] : J[1];
			if(b.G0raw){ b.G0raw[(size_t)p * 2 * N + it.pix] = smp.gx; b.G0raw[(size_t)p * 2 * N + N + it.pix] = smp.gy; }
			if(b.It_scratch) b.It_scratch[(size_t)p*N + it.pix] = val;          // am.initializePixVals: It = I0
		}
		const BinWeights bw = bin_weights(val, B);
#pragma unroll
		for(int k = 0; k < 4; ++k){
			if(bw.lo + k > bw.hi) continue;
			atomicAdd(&s_hist[bw.lo + k], bw.w[k]);
#pragma unroll
			for(int l = 0; l < 4; ++l){
				if(bw.lo + l > bw.hi) continue;
				atomicAdd(&s_joint[(bw.lo + l)*B + (bw.lo + k)], bw.w[k] * bw.w[l]);        // JH(id1, id2): column-major
			}
		}
	}
	cta_sync<T>();
	for(int i = tid; i < B; i += T){
		const double h = s_hist[i] * mp.hist_norm_mult;
		s_hist[i] = h; s_hist_log[i] = log(h);
		if(!REINIT){ mi_tab[(size_t)p*MI_TAB + i] = h; mi_tab[(size_t)p*MI_TAB + 16 + i] = s_hist_log[i]; }
	}
	cta_sync<T>();
	for(int i = tid; i < B*B; i += T){
		const double jh = s_joint[i] * mp.hist_norm_mult;
		s_joint[i] = jh;
		const int r = i % B;                                              // JH(r, c) = m[c*B + r]
		s_factor[i] = 1 + log(jh) - s_hist_log[r];                        // self_grad_factor(curr = r, init = c) (MI.cc:655)
	}
	cta_sync<T>();
	if(!REINIT && tid == 0){
		// f = max_similarity = MI of the template with itself (MI.cc:270-283)
		double f = 0;
		for(int c = 0; c < B; ++c) for(int r = 0; r < B; ++r){
			const double jh = s_joint[c*B + r];
			f += jh * (log(jh) - s_hist_log[r] - s_hist_log[c]);
		}
		b.f[p] = f;
	}
	// phase 2: init_self_hessian = cmptSelfHessian(init_pix_jacobian) (MI.cc:515-594; at initialize() the current
	// histograms are the initial ones)
	double acc[NH];
#pragma unroll
	for(int i = 0; i < NH; ++i) acc[i] = 0;
	for(PixIter it(tid, T, b.resx); it.pix < N; it.next(T)){
		PixGeom g = pixel_geometry<SSM>(dlt, W, b.xv[it.col], b.yv[it.row], b.norm_init != 0);
		double D[S];
		init_pix_jacobian<SSM>(g.ix, g.iy, G0[it.pix], G0[N + it.pix], D);
		const double v = Vh[it.pix];
		const BinWeights bw = bin_weights(v, B);
		double x = bw.lo - v, hist_hess_term = 0;
#pragma unroll
		for(int k = 0; k < 4; ++k){
			const int curr_id = bw.lo + k;
			const double hess_w = mp.hist_norm_mult * bspl3_hess(x);
			x += 1;
			if(curr_id > bw.hi) continue;
			const double grad_k = bw.d[k] * (-mp.hist_norm_mult);           // curr_hist_grad (MI.cc:240)
			double inner = 0;
#pragma unroll
			for(int l = 0; l < 4; ++l){
				const int init_id = bw.lo + l;
				if(init_id > bw.hi) continue;
				const double gq = grad_k * bw.w[l];
#pragma unroll
				for(int s = 0; s < S; ++s) atomicAdd(&s_jhj[(curr_id*B + init_id)*S + s], gq * D[s]);
				inner += bw.w[l] * s_factor[init_id*B + curr_id];
			}
			hist_hess_term += hess_w * inner;
		}
#pragma unroll
		for(int i = 0; i < S; ++i){
#pragma unroll
			for(int j = i; j < S; ++j) acc[i*S - i*(i - 1) / 2 + (j - i)] = fma(hist_hess_term * D[i], D[j], acc[i*S - i*(i - 1) / 2 + (j - i)]);
		}
	}
	block_reduce<NH, T>(acc, s_part, s_sum);             // also orders the atomics on s_jhj before the reads below
	for(int e = tid; e < S*S; e += T){
		const int i = e % S, j = e / S;
		const int lo = i < j ? i : j, hi = i < j ? j : i;
		double h = s_sum[lo*S - lo*(lo - 1) / 2 + (hi - lo)];
		// + sum over joint bins of row^T row (1 / self_joint(curr, init) - 1 / curr_hist(curr)) (MI.cc:578-590)
		for(int curr_id = 0; curr_id < B; ++curr_id) for(int init_id = 0; init_id < B; ++init_id){
			const double hist_factor = (1.0 / s_joint[init_id*B + curr_id]) - (1.0 / s_hist[curr_id]);
			const double *row = &s_jhj[(curr_id*B + init_id)*S];
			h += row[i] * row[j] * hist_factor;
		}
		b.Hinit[(size_t)p * 64 + j*S + i] = h;
	}
}

// KEEP_IT: keep the N current pixel values of sweep 1 in shared memory for sweep 2 (small patches); otherwise sweep 2
// samples them again (same bits) and shared memory holds only the private histograms -- for 100 x 100 patches the
// 80 KB value buffer would leave room for two warps per SM.
// SELF: the pass also builds am.cmptSelfHessian(curr_pix_jacobian) (MI.cc:515-594) -- the CurrentSelf / SumOfSelf Hessians of
// FCLK / ESM / ICLK: cmptSelfHist (the joint histogram of the current patch with itself, MI.cc:639-658), its logs, and one
// more sweep over the pixels with the same code as the template's init_self_hessian in mi_init_kernel.
// HMODE 2: the Std forms instead -- cmptCurrHessian (MI.cc:603-637: FCLK / ESM Std), cmptInitHessian (MI.cc:461-514: ICLK Std),
// half their sum (ESM SumOfStd, NT/ESM.cc:339-343) -- through the same sweep with the template's bins as the partner.
template<int SSM, int SM, int T, bool KEEP_IT, int HMODE>
__global__ void __launch_bounds__(T, min_blocks(T, 1)) mi_update_kernel(DevBatch b, MiParams mp, const double *__restrict__ mi_tab){
	constexpr bool SELF = (HMODE != 0);                 // a per-pass Hessian is assembled in s_Hc
	constexpr int S = StateSize<SSM>::value;
	constexpr int NH = S*(S + 1) / 2;
	constexpr bool CURR = (SM != SM_ICLK), INIT = (SM != SM_FCLK);
	constexpr int NA = (CURR ? S : 0) + (INIT ? S : 0);
	constexpr int oT = 0, o0 = CURR ? S : 0;
	extern __shared__ __align__(16) double s_dyn[];
	const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int B = mp.B, N = b.N;
	const int HB = B + B*B;                                                 // curr_hist | joint_hist entries
	double *s_It = s_dyn;                                                   // N current pixel values (bin units), if KEEP_IT
	const int it_slots = KEEP_IT ? ((N + 1) & ~1) : 0;
	// PRIVATE copies of both histograms, [entry][copy]: C = mp.copies = 32 (one per lane) or 16 (lanes l and l + 16 share
	// copy l and take turns).  A lane only ever touches its copy's column, so the 20 updates per pixel are plain
	// read-modify-writes (conflict-free banks, no atomics, fixed summation order).  16 copies halve the 18 KB per warp that
	// cap the residency at 8 warps per SM (profiles/r02_ncu_mi_summary.txt: 12 % warps active, long-scoreboard bound)
	const int C = mp.copies, cl = lane & (C - 1), n_turns = 32 / C, my_turn = lane / C;
	double *s_priv = s_dyn + it_slots + (size_t)warp*HB * C;
	__shared__ double s_part[(T / 32) * NA];
	__shared__ double s_sum[NA];
	__shared__ double s_hist[MI_BMAX], s_hist_log[MI_BMAX], s_ihist_log[MI_BMAX], s_joint[MI_BMAX*MI_BMAX];
	__shared__ double s_fac_t[MI_BMAX*MI_BMAX], s_fac_0[MI_BMAX*MI_BMAX];
	__shared__ double s_fold[(T / 32) * MI_HBMAX];
	__shared__ double s_W[9], s_dlt[9], s_corners[8], s_init_corners[8], s_J[S], s_f;
	__shared__ int s_ctrl;
	// SELF only (size 1 otherwise): self joint histogram, its gradient factor, joint_hist_jacobian, the Hessian sums
	__shared__ double s_sj[SELF ? MI_BMAX*MI_BMAX : 1], s_sfac[SELF ? MI_BMAX*MI_BMAX : 1], s_jhj[SELF ? MI_BMAX*MI_BMAX*S : 1];
	__shared__ double s_part2[SELF ? (T / 32) * NH : 1], s_sum2[SELF ? NH : 1], s_Hc[SELF ? S*S : 1];
	__shared__ double s_jh[HMODE == 2 ? MI_BMAX*MI_BMAX : 1], s_ihist[HMODE == 2 ? MI_BMAX : 1];      // joint_hist, init_hist
	if(HMODE == 2){ for(int i = tid; i < B; i += T) s_ihist[i] = mi_tab[(size_t)p*MI_TAB + i]; }
	if(tid < 9){ s_W[tid] = b.warp[(size_t)p * 9 + tid]; s_dlt[tid] = b.dlt[(size_t)p * 9 + tid]; }
	if(tid < 8){ s_corners[tid] = b.corners[(size_t)p * 8 + tid]; s_init_corners[tid] = b.init_corners[(size_t)p * 8 + tid]; }
	for(int i = tid; i < B; i += T) s_ihist_log[i] = mi_tab[(size_t)p*MI_TAB + 16 + i];
	cta_sync<T>();
	Mat3 dlt;
#pragma unroll
	for(int i = 0; i < 9; ++i) dlt.m[i] = s_dlt[i];
	const double *I0 = b.I0 + (size_t)p*N, *G0 = b.G0 + (size_t)p * 2 * N;
	double *It_g = (!KEEP_IT && b.It_scratch) ? b.It_scratch + (size_t)p*N : nullptr;
	const bool jac_orig = (SM == SM_ESM) && (b.jac_type == MTFB_ESM_JAC_ORIGINAL);
	const bool hess_orig = (SM == SM_ESM) && (b.hess_type == MTFB_ESM_HESS_ORIGINAL);        // cmptCurrHessian(mean_pix_jacobian)
	const bool jac_half = (SM == SM_ESM) && !jac_orig;
	LMState lm = { 0.0, b.lm_delta_init, 0.0, false };
	int iter_id = 0, n_passes = 0, patch_status = 0;
	double f = 0;
	while(iter_id < b.max_iters){
		Mat3 W;
#pragma unroll
		for(int i = 0; i < 9; ++i) W.m[i] = s_W[i];
		double abcd[4] = { (W.m[0] - 1) + 1, W.m[1], W.m[3], (W.m[4] - 1) + 1 };
		if(lane < C) for(int e = 0; e < HB; ++e) s_priv[e * C + lane] = 0;
		__syncwarp();
		// ---- sweep 1: updatePixVals (MI.cc:166-192) + the histograms of updateSimilarity (MI.cc:346-370)
		// (the trip count is uniform over the warp so that the lanes sharing a copy can take turns; lanes past the end add nothing)
		// (the template value of the NEXT trip is requested before this trip's arithmetic: the stream comes from DRAM, and with
		// ~13 warps per SM nothing else hides its latency)
		double i0_next = tid < N ? __ldcs(I0 + tid) : 0.0;
		for(PixIter it(tid, T, b.resx); it.pix - lane < N; it.next(T)){
			const bool live = it.pix < N;
			const double i0 = i0_next;
			if(it.pix + T < N) i0_next = __ldcs(I0 + it.pix + T);
			BinWeights bc, bi;
			bc.lo = 0; bc.hi = -1; bi.lo = 0; bi.hi = -1;
			if(live){
				PixGeom g = pixel_geometry<SSM>(dlt, W, b.xv[it.col], b.yv[it.row], b.norm_init != 0);
				const double It = b.pix_mult*sample_pixel(b.img, g.wx, g.wy) + b.pix_add;
				if(KEEP_IT) s_It[it.pix] = It;
				else if(It_g) __stcs(It_g + it.pix, It);                           // read back by this thread in sweep 2
				bc = bin_weights(It, B); bi = bin_weights(i0, B);
			}
			for(int turn = 0; turn < n_turns; ++turn){
				if(live && turn == my_turn){
#pragma unroll
					for(int k = 0; k < 4; ++k){
						if(bc.lo + k > bc.hi) continue;
						s_priv[(bc.lo + k) * C + cl] += bc.w[k];
#pragma unroll
						for(int l = 0; l < 4; ++l){
							if(bi.lo + l > bi.hi) continue;
							s_priv[(B + (bi.lo + l)*B + (bc.lo + k)) * C + cl] += bc.w[k] * bi.w[l];    // JH(curr_id, init_id)
						}
					}
				}
				if(n_turns > 1) __syncwarp();
			}
		}
		__syncwarp();
		// fold the private columns (shuffle tree), then the warps, in a fixed order
		for(int e = 0; e < HB; ++e){
			double v = lane < C ? s_priv[e * C + lane] : 0.0;
#pragma unroll
			for(int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(FULL_MASK, v, off);
			if(lane == 0) s_fold[warp*MI_HBMAX + e] = v;
		}
		cta_sync<T>();
		for(int e = tid; e < HB; e += T){
			double v = e < B ? mp.hist_pre_seed : mp.pre_seed;
#pragma unroll
			for(int w = 0; w < T / 32; ++w) v += s_fold[w*MI_HBMAX + e];
			if(e < B) s_hist[e] = v; else s_joint[e - B] = v;
		}
		cta_sync<T>();
		for(int i = tid; i < B; i += T){
			const double h = s_hist[i] * mp.hist_norm_mult;
			s_hist[i] = h; s_hist_log[i] = log(h);
		}
		cta_sync<T>();
		for(int i = tid; i < B*B; i += T){
			const double jh = s_joint[i] * mp.hist_norm_mult, jl = log(jh);
			if(HMODE == 2) s_jh[i] = jh;
			const int curr_id = i % B, init_id = i / B;
			s_joint[i] = jh * (jl - s_hist_log[curr_id] - s_ihist_log[init_id]);      // the term of f (MI.cc:376-380)
			s_fac_t[i] = 1 + jl - s_hist_log[curr_id];                                // curr_grad_factor(curr, init) (MI.cc:430)
			s_fac_0[i] = 1 + jl - s_ihist_log[init_id];                               // init_grad_factor(init, curr) (MI.cc:402)
		}
		cta_sync<T>();
		if(tid == 0){
			double fs = 0;
			for(int curr_id = 0; curr_id < B; ++curr_id) for(int init_id = 0; init_id < B; ++init_id) fs += s_joint[init_id*B + curr_id];
			s_f = fs;
		}
		// ---- sweep 2: df_dIt / df_dI0 per pixel (MI.cc:426-443, 398-417) and the Jacobian sums
		double acc[NA];
#pragma unroll
		for(int i = 0; i < NA; ++i) acc[i] = 0;
		double p_i0 = tid < N ? __ldcs(I0 + tid) : 0.0, p_gx = 0, p_gy = 0, p_it = 0;
		if(INIT && tid < N){ p_gx = __ldcs(G0 + tid); p_gy = __ldcs(G0 + N + tid); }
		if(!KEEP_IT && It_g && tid < N) p_it = __ldcs(It_g + tid);
		for(PixIter it(tid, T, b.resx); it.pix < N; it.next(T)){
			const double c_i0 = p_i0, c_gx = p_gx, c_gy = p_gy, c_it = p_it;
			if(it.pix + T < N){
				p_i0 = __ldcs(I0 + it.pix + T);
				if(INIT){ p_gx = __ldcs(G0 + it.pix + T); p_gy = __ldcs(G0 + N + it.pix + T); }
				if(!KEEP_IT && It_g) p_it = __ldcs(It_g + it.pix + T);
			}
			PixGeom g = pixel_geometry<SSM>(dlt, W, b.xv[it.col], b.yv[it.row], b.norm_init != 0);
			// the pixel's value of sweep 1: from shared memory (small templates), from the pass's scratch row in global memory
			// (this thread wrote it), or -- neither available -- sampled again
			const double It2 = KEEP_IT ? s_It[it.pix] : (It_g ? c_it : (b.pix_mult*sample_pixel(b.img, g.wx, g.wy) + b.pix_add));
			const BinWeights bc = bin_weights(It2, B), bi = bin_weights(c_i0, B);
			double df_t = 0, df_0 = 0;
			if(CURR){
#pragma unroll
				for(int k = 0; k < 4; ++k){
					if(bc.lo + k > bc.hi) continue;
					const double gk = bc.d[k] * (-mp.hist_norm_mult);                   // curr_hist_grad
#pragma unroll
					for(int l = 0; l < 4; ++l){
						if(bi.lo + l > bi.hi) continue;
						df_t += (gk * bi.w[l]) * s_fac_t[(bi.lo + l)*B + (bc.lo + k)];
					}
				}
			}
			if(INIT){
#pragma unroll
				for(int l = 0; l < 4; ++l){
					if(bi.lo + l > bi.hi) continue;
					const double gl = bi.d[l] * (-mp.hist_norm_mult);                   // init_hist_grad
#pragma unroll
					for(int k = 0; k < 4; ++k){
						if(bc.lo + k > bc.hi) continue;
						df_0 += (gl * bc.w[k]) * s_fac_0[(bi.lo + l)*B + (bc.lo + k)];
					}
				}
			}
			double D[S], D0[S];
			if(INIT) init_pix_jacobian<SSM>(g.ix, g.iy, c_gx, c_gy, D0);
			if(CURR){
				Sample smp;
				pixel_value_and_gradient<SSM, false>(b, W, g, smp);
				pixel_jacobian_row<SSM>(b, W, abcd, g, smp.gx, smp.gy, D);
				// ESM's Original Jacobian: df_dIt . mean_pix_jacobian (NT/ESM.cc:246-248, 301-303)
				if(SM == SM_ESM && jac_orig){
#pragma unroll
					for(int i = 0; i < S; ++i) D[i] = (D0[i] + D[i]) / 2.0;
				}
#pragma unroll
				for(int i = 0; i < S; ++i) acc[oT + i] = fma(df_t, D[i], acc[oT + i]);
			}
			if(INIT){
#pragma unroll
				for(int i = 0; i < S; ++i) acc[o0 + i] = fma(df_0, D0[i], acc[o0 + i]);
			}
		}
		block_reduce<NA, T>(acc, s_part, s_sum);
		++n_passes;
		if(SELF){
			// which Hessians this pass assembles: 0 = cmptSelfHessian(curr), 1 = cmptCurrHessian(curr), 2 = cmptInitHessian(init)
			int modes[2] = { 0, 0 }, n_modes = 1;
			double coef = 1.0;
			if(HMODE == 2){
				if(SM == SM_ICLK) modes[0] = 2;
				else if(SM == SM_ESM && b.hess_type == MTFB_ESM_HESS_SUM_OF_STD){ modes[0] = 1; modes[1] = 2; n_modes = 2; coef = 0.5; }
				else modes[0] = 1;
			}
			for(int mi = 0; mi < n_modes; ++mi){
				const int mode = modes[mi];
				for(int i = tid; i < B*B*S; i += T) s_jhj[i] = 0;
				if(mode == 0){
					// ---- cmptSelfHist (MI.cc:639-658): joint histogram of the current patch with itself, logs, self_grad_factor
					for(int i = tid; i < B*B; i += T) s_sj[i] = mp.pre_seed;
					cta_sync<T>();
					for(PixIter it(tid, T, b.resx); it.pix < N; it.next(T)){
						double v;
						if(KEEP_IT) v = s_It[it.pix];
						else{
							PixGeom g = pixel_geometry<SSM>(dlt, W, b.xv[it.col], b.yv[it.row], b.norm_init != 0);
							v = b.pix_mult*sample_pixel(b.img, g.wx, g.wy) + b.pix_add;
						}
						const BinWeights bw = bin_weights(v, B);
#pragma unroll
						for(int k = 0; k < 4; ++k){
							if(bw.lo + k > bw.hi) continue;
#pragma unroll
							for(int l = 0; l < 4; ++l){
								if(bw.lo + l > bw.hi) continue;
								atomicAdd(&s_sj[(bw.lo + l)*B + (bw.lo + k)], bw.w[k] * bw.w[l]);
							}
						}
					}
					cta_sync<T>();
					for(int i = tid; i < B*B; i += T){
						const double jh = s_sj[i] * mp.hist_norm_mult;
						s_sj[i] = jh;
						s_sfac[i] = 1 + log(jh) - s_hist_log[i % B];                      // self_grad_factor(curr, init) (MI.cc:655)
					}
				}
				cta_sync<T>();
				// ---- the Hessian sweep (MI.cc:558-592 / 614-636 / 472-512).  p = the bins of the patch whose pixel Jacobian is
				// used (current patch, or the template for cmptInitHessian), q = the partner's bins
				double acc2[NH];
#pragma unroll
				for(int i = 0; i < NH; ++i) acc2[i] = 0;
				for(PixIter it(tid, T, b.resx); it.pix < N; it.next(T)){
					PixGeom g = pixel_geometry<SSM>(dlt, W, b.xv[it.col], b.yv[it.row], b.norm_init != 0);
					double D[S], vc;
					if(mode == 2){
						init_pix_jacobian<SSM>(g.ix, g.iy, G0[it.pix], G0[N + it.pix], D);
						vc = KEEP_IT ? s_It[it.pix] : (b.pix_mult*sample_pixel(b.img, g.wx, g.wy) + b.pix_add);
					} else{
						Sample smp;
						pixel_value_and_gradient<SSM, false>(b, W, g, smp);
						pixel_jacobian_row<SSM>(b, W, abcd, g, smp.gx, smp.gy, D);
						vc = KEEP_IT ? s_It[it.pix] : (b.pix_mult*smp.val + b.pix_add);
						if(SM == SM_ESM && hess_orig){
							double D0[S];
							init_pix_jacobian<SSM>(g.ix, g.iy, G0[it.pix], G0[N + it.pix], D0);
#pragma unroll
							for(int i = 0; i < S; ++i) D[i] = (D0[i] + D[i]) / 2.0;
						}
					}
					const double vp = (mode == 2) ? I0[it.pix] : vc;
					const BinWeights bp = bin_weights(vp, B);
					const BinWeights bq = (mode == 0) ? bp : bin_weights(mode == 1 ? I0[it.pix] : vc, B);
					double x = bp.lo - vp, hist_hess_term = 0;
#pragma unroll
					for(int k = 0; k < 4; ++k){
						const int p_id = bp.lo + k;
						const double hess_w = mp.hist_norm_mult * bspl3_hess(x);
						x += 1;
						if(p_id > bp.hi) continue;
						const double grad_k = bp.d[k] * (-mp.hist_norm_mult);               // curr_hist_grad / init_hist_grad (MI.cc:240)
						double inner = 0;
#pragma unroll
						for(int l = 0; l < 4; ++l){
							const int q_id = bq.lo + l;
							if(q_id > bq.hi) continue;
							const double gq = grad_k * bq.w[l];
							// joint_hist_jacobian row (curr_id, init_id); the factor tables are indexed [init * B + curr]
							const int curr_id = (mode == 2) ? q_id : p_id, init_id = (mode == 2) ? p_id : q_id;
#pragma unroll
							for(int r = 0; r < S; ++r) atomicAdd(&s_jhj[(curr_id*B + init_id)*S + r], gq * D[r]);
							inner += bq.w[l] * (mode == 0 ? s_sfac[init_id*B + curr_id] : mode == 1 ? s_fac_t[init_id*B + curr_id] : s_fac_0[init_id*B + curr_id]);
						}
						hist_hess_term += hess_w * inner;
					}
#pragma unroll
					for(int i = 0; i < S; ++i){
#pragma unroll
						for(int j = i; j < S; ++j) acc2[i*S - i*(i - 1) / 2 + (j - i)] = fma(hist_hess_term * D[i], D[j], acc2[i*S - i*(i - 1) / 2 + (j - i)]);
					}
				}
				block_reduce<NH, T>(acc2, s_part2, s_sum2);             // also orders the atomics on s_jhj before the reads below
				for(int e = tid; e < S*S; e += T){
					const int i = e % S, j = e / S;
					const int lo = i < j ? i : j, hi = i < j ? j : i;
					double h = s_sum2[lo*S - lo*(lo - 1) / 2 + (hi - lo)];
					for(int curr_id = 0; curr_id < B; ++curr_id) for(int init_id = 0; init_id < B; ++init_id){
						const double joint = (mode == 0) ? s_sj[init_id*B + curr_id] : s_jh[(HMODE == 2) ? init_id*B + curr_id : 0];
						const double marg = (mode == 2) ? s_ihist[(HMODE == 2) ? init_id : 0] : s_hist[curr_id];
						const double hist_factor = (1.0 / joint) - (1.0 / marg);
						const double *row = &s_jhj[(curr_id*B + init_id)*S];
						h += row[i] * row[j] * hist_factor;
					}
					s_Hc[j*S + i] = (mi == 0 ? 0.0 : s_Hc[j*S + i]) + coef*h;
				}
				cta_sync<T>();
			}
		}
		if(tid < S){
			double jv = CURR ? s_sum[oT + tid] : 0.0;
			if(SM == SM_ESM && !jac_orig) jv = jv - s_sum[o0 + tid];                // AppearanceModel.h:162-166
			if(SM == SM_ICLK) jv = s_sum[o0 + tid];
			s_J[tid] = jac_half ? jv * 0.5 : jv;
		}
		cta_sync<T>();
		if(warp == 0){
			f = s_f;
			// HMODE 2: s_Hc is the complete Hessian of the pass
			const int ctrl = (HMODE == 2) ? serial_step<SSM, SM, false, 0>(b, p, lane, iter_id, n_passes, f, s_J, s_Hc, s_W, s_corners,
				s_init_corners, lm, patch_status) : serial_step<SSM, SM>(b, p, lane, iter_id, n_passes, f, s_J, SELF ? s_Hc : nullptr, s_W,
				s_corners, s_init_corners, lm, patch_status);
			if(lane == 0) s_ctrl = ctrl;
		}
		cta_sync<T>();
		const int ctrl = s_ctrl;
		if(ctrl == CTRL_BREAK) break;
		if(counts_as_iteration<SM>(ctrl, b.nt_semantics)) ++iter_id;
	}
	if(warp == 0) store_patch_state<SSM>(b, p, lane, s_W, s_corners, f, n_passes, patch_status);
	// the AM's pixel values stay those of the last pass: what setRegion's cmptSelfHessian will see (mi_init_kernel<REINIT>);
	// large templates already keep them in the scratch row
	if(KEEP_IT && b.It_scratch) for(int pix = tid; pix < N; pix += T) b.It_scratch[(size_t)p*N + pix] = s_It[pix];
}

// ------------------------------------------------------------------------------------------------
static MiParams make_mi_params(const DevBatch &b, int n_bins, double pre_seed){
	MiParams mp;
	mp.B = n_bins; mp.pre_seed = pre_seed;
	mp.hist_pre_seed = n_bins * pre_seed;                                          // MI.cc:101
	mp.hist_norm_mult = 1.0 / (static_cast<double>(b.N) + mp.hist_pre_seed * n_bins);   // MI.cc:104
	mp.copies = 32;
	return mp;
}

#ifndef MTFB_MI_AFFINE_TU
template<int SSM> static cudaError_t launch_init_t(int threads, const DevBatch &b, const double *d_corners, const MiParams &mp,
	double *mi_tab, cudaStream_t st){
	switch(threads){
	case 32: mi_init_kernel<SSM, 32><<<b.P, 32, 0, st>>>(b, d_corners, mp, mi_tab); break;
	case 64: mi_init_kernel<SSM, 64><<<b.P, 64, 0, st>>>(b, d_corners, mp, mi_tab); break;
	case 128: mi_init_kernel<SSM, 128><<<b.P, 128, 0, st>>>(b, d_corners, mp, mi_tab); break;
	case 256: mi_init_kernel<SSM, 256><<<b.P, 256, 0, st>>>(b, d_corners, mp, mi_tab); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}
template<int SSM> static cudaError_t launch_reinit_t(int threads, const DevBatch &b, const double *d_corners, const MiParams &mp,
	double *mi_tab, cudaStream_t st){
	switch(threads){
	case 32: mi_init_kernel<SSM, 32, true><<<b.P, 32, 0, st>>>(b, d_corners, mp, mi_tab); break;
	case 64: mi_init_kernel<SSM, 64, true><<<b.P, 64, 0, st>>>(b, d_corners, mp, mi_tab); break;
	case 128: mi_init_kernel<SSM, 128, true><<<b.P, 128, 0, st>>>(b, d_corners, mp, mi_tab); break;
	case 256: mi_init_kernel<SSM, 256, true><<<b.P, 256, 0, st>>>(b, d_corners, mp, mi_tab); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}
cudaError_t launch_reinit_mi(int ssm, int threads, const DevBatch &b, const double *d_corners, int n_bins, double pre_seed,
	double *mi_tab, cudaStream_t st){
	if(n_bins < 4 || n_bins > MI_BMAX || !b.It_scratch || !b.G0raw) return cudaErrorInvalidValue;
	const MiParams mp = make_mi_params(b, n_bins, pre_seed);
	if(ssm == SSM_HOM) return launch_reinit_t<SSM_HOM>(threads, b, d_corners, mp, mi_tab, st);
	return launch_reinit_t<SSM_AFF>(threads, b, d_corners, mp, mi_tab, st);
}
cudaError_t launch_init_mi(int ssm, int threads, const DevBatch &b, const double *d_corners, int n_bins, double pre_seed,
	double *mi_tab, cudaStream_t st){
	if(n_bins < 4 || n_bins > MI_BMAX) return cudaErrorInvalidValue;
	const MiParams mp = make_mi_params(b, n_bins, pre_seed);
	if(ssm == SSM_HOM) return launch_init_t<SSM_HOM>(threads, b, d_corners, mp, mi_tab, st);
	return launch_init_t<SSM_AFF>(threads, b, d_corners, mp, mi_tab, st);
}

#endif

template<int SSM, int SM, int T, bool KEEP_IT, int HMODE> static cudaError_t launch_self(const DevBatch &b, const MiParams &mp, const double *mi_tab, cudaStream_t st){
	// private histogram copies per warp: 16 when the batch alone can fill the SMs' warp slots (then shared memory, not the
	// batch size, caps the residency), else one per lane
	MiParams mq = mp;
	mq.copies = ((size_t)b.P * (T / 32) >= 148 * 12) ? 16 : 32;
	if(const char *ev = std::getenv("MTFB_MI_COPIES")) mq.copies = (std::atoi(ev) == 16) ? 16 : 32;
	const size_t smem = ((KEEP_IT ? (size_t)((b.N + 1) & ~1) : 0) + (size_t)(T / 32) * (mp.B + mp.B*mp.B) * mq.copies) * sizeof(double);
	if(smem > 180 * 1024) return cudaErrorInvalidValue;
	cudaError_t e = cudaFuncSetAttribute(mi_update_kernel<SSM, SM, T, KEEP_IT, HMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if(e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(mi_update_kernel<SSM, SM, T, KEEP_IT, HMODE>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
	if(e != cudaSuccess) return e;
	mi_update_kernel<SSM, SM, T, KEEP_IT, HMODE><<<b.P, T, smem, st>>>(b, mq, mi_tab);
	return cudaGetLastError();
}
template<int SSM, int SM, int T, bool KEEP_IT> static cudaError_t launch_keep(const DevBatch &b, const MiParams &mp, const double *mi_tab, cudaStream_t st){
	// the pass's own self Hessian is needed by CurrentSelf (all three searches) and ESM's SumOfSelf
	const bool self = (SM == SM_ESM) ? (b.hess_type == MTFB_ESM_HESS_CURRENT_SELF || b.hess_type == MTFB_ESM_HESS_SUM_OF_SELF)
		: (b.hess_type == MTFB_LK_HESS_CURRENT_SELF);
	const bool std_hess = (SM == SM_ESM) ? (b.hess_type == MTFB_ESM_HESS_STD || b.hess_type == MTFB_ESM_HESS_SUM_OF_STD ||
		b.hess_type == MTFB_ESM_HESS_ORIGINAL) : (b.hess_type == MTFB_LK_HESS_STD);
	if(std_hess) return launch_self<SSM, SM, T, KEEP_IT, 2>(b, mp, mi_tab, st);
	return self ? launch_self<SSM, SM, T, KEEP_IT, 1>(b, mp, mi_tab, st) : launch_self<SSM, SM, T, KEEP_IT, 0>(b, mp, mi_tab, st);
}
template<int SSM, int SM, int T> static cudaError_t launch_one(const DevBatch &b, const MiParams &mp, const double *mi_tab, cudaStream_t st){
	if((size_t)b.N * sizeof(double) <= 24 * 1024) return launch_keep<SSM, SM, T, true>(b, mp, mi_tab, st);
	return launch_keep<SSM, SM, T, false>(b, mp, mi_tab, st);
}
template<int SSM, int SM> static cudaError_t launch_update_t(int threads, const DevBatch &b, const MiParams &mp, const double *mi_tab,
	cudaStream_t st){
	// one or two warps per patch: the private histogram copies (18 KB per warp at 8 bins) leave no room for more, and the
	// library never picks more for MI (mtfb_create); wider requests run at two
	switch(threads){
	case 32: return launch_one<SSM, SM, 32>(b, mp, mi_tab, st);
	case 64: case 128: case 256: return launch_one<SSM, SM, 64>(b, mp, mi_tab, st);
	default: return cudaErrorInvalidValue;
	}
}
#ifndef MTFB_MI_AFFINE_TU
// ------------------------------------------------------------------------------------------------
// Particle evaluation for MI (the particle loop of nt::PF::update, SM/src/NT/PF.cc:303-320, with am = MI):
// ssm->setState -> am->updatePixVals (MI.cc:166-192) -> am->updateSimilarity(false) (MI.cc:346-382: curr_hist and
// joint_hist through the 4 x 4 B-spline weights, logs, f = sum joint (log joint - log curr - log init)) ->
// am->getLikelihood() (MI.cc:384-387).  One warp per particle with a private copy of both histograms per lane, as
// in the update kernel (conflict-free read-modify-writes, fixed summation order); the template's values stay in shared memory.
// ------------------------------------------------------------------------------------------------
template<int SSM, int T>
__global__ void __launch_bounds__(T) pf_evaluate_mi_kernel(DevBatch b, MiParams mp, const double *__restrict__ mi_tab,
	const double *__restrict__ states, int n_particles, double *__restrict__ likelihood, double *__restrict__ similarity, double alpha){
	constexpr int S = StateSize<SSM>::value;
	constexpr int NW = T / 32;
	extern __shared__ __align__(16) double s_dyn[];
	const int obj = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int B = mp.B, N = b.N, HB = B + B*B;
	double *s_I0 = s_dyn;                                                   // N template values (bin units)
	double *s_priv = s_dyn + ((N + 1) & ~1) + (size_t)warp*HB * 32;          // [entry][lane] private histograms of this warp
	__shared__ double s_ihist_log[MI_BMAX], s_h[NW][MI_HBMAX], s_hlog[NW][MI_BMAX];
	for(int i = tid; i < N; i += T) s_I0[i] = b.I0[(size_t)obj*N + i];
	for(int i = tid; i < B; i += T) s_ihist_log[i] = mi_tab[(size_t)obj*MI_TAB + 16 + i];
	Mat3 dlt;
#pragma unroll
	for(int i = 0; i < 9; ++i) dlt.m[i] = b.dlt[(size_t)obj * 9 + i];
	__syncthreads();
	for(int pi = blockIdx.x*NW + warp; pi < n_particles; pi += gridDim.x*NW){
		const double *st = states + ((size_t)obj*n_particles + pi)*S;
		double sv[S];
#pragma unroll
		for(int q = 0; q < S; ++q) sv[q] = st[q];
		const Mat3 W = warp_from_state<SSM>(sv);
		for(int e = 0; e < HB; ++e) s_priv[e * 32 + lane] = 0;
		for(PixIter it(lane, 32, b.resx); it.pix < N; it.next(32)){
			PixGeom g = pixel_geometry<SSM>(dlt, W, b.xv[it.col], b.yv[it.row], b.norm_init != 0);
			const double It = b.pix_mult*sample_pixel(b.img, g.wx, g.wy) + b.pix_add;
			const BinWeights bc = bin_weights(It, B), bi = bin_weights(s_I0[it.pix], B);
#pragma unroll
			for(int k = 0; k < 4; ++k){
				if(bc.lo + k > bc.hi) continue;
				s_priv[(bc.lo + k) * 32 + lane] += bc.w[k];
#pragma unroll
				for(int l = 0; l < 4; ++l){
					if(bi.lo + l > bi.hi) continue;
					s_priv[(B + (bi.lo + l)*B + (bc.lo + k)) * 32 + lane] += bc.w[k] * bi.w[l];    // JH(curr_id, init_id)
				}
			}
		}
		__syncwarp();
		for(int e = 0; e < HB; ++e){
			double v = s_priv[e * 32 + lane];
#pragma unroll
			for(int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(FULL_MASK, v, off);
			if(lane == 0) s_h[warp][e] = ((e < B ? mp.hist_pre_seed : mp.pre_seed) + v) * mp.hist_norm_mult;
		}
		__syncwarp();
		for(int i = lane; i < B; i += 32) s_hlog[warp][i] = log(s_h[warp][i]);
		__syncwarp();
		double fs = 0;
		for(int i = lane; i < B*B; i += 32){
			const double jh = s_h[warp][B + i];
			fs += jh * (log(jh) - s_hlog[warp][i % B] - s_ihist_log[i / B]);            // MI.cc:376-380
		}
#pragma unroll
		for(int off = 16; off >= 1; off >>= 1) fs += __shfl_xor_sync(FULL_MASK, fs, off);
		if(lane == 0){
			if(similarity) similarity[(size_t)obj*n_particles + pi] = fs;
			const double d = (1.0 / fs) - 1;                                            // MI.cc:384-387
			if(likelihood) likelihood[(size_t)obj*n_particles + pi] = exp(-alpha * d*d);
		}
		__syncwarp();
	}
}

cudaError_t launch_pf_evaluate_mi(int ssm, const DevBatch &b, int n_bins, double pre_seed, const double *mi_tab, const double *d_states,
	int n_particles, double *d_likelihood, double *d_similarity, double alpha, cudaStream_t st){
	if(n_bins < 4 || n_bins > MI_BMAX) return cudaErrorInvalidValue;
	const MiParams mp = make_mi_params(b, n_bins, pre_seed);
	constexpr int T = 128;
	const size_t smem = ((size_t)((b.N + 1) & ~1) + (size_t)(T / 32) * (mp.B + mp.B*mp.B) * 32) * sizeof(double);
	if(smem > 200 * 1024) return cudaErrorInvalidValue;
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const int ctas_per_sm = (int)((200 * 1024) / (smem + 2048)) < 1 ? 1 : (int)((200 * 1024) / (smem + 2048));
	int per_obj = (sms*ctas_per_sm + b.P - 1) / b.P;
	const int max_useful = (n_particles + T / 32 - 1) / (T / 32);
	if(per_obj > max_useful) per_obj = max_useful;
	if(per_obj < 1) per_obj = 1;
	const dim3 grid(per_obj, b.P);
	cudaError_t e;
	if(ssm == SSM_HOM){
		e = cudaFuncSetAttribute(pf_evaluate_mi_kernel<SSM_HOM, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if(e != cudaSuccess) return e;
		pf_evaluate_mi_kernel<SSM_HOM, T><<<grid, T, smem, st>>>(b, mp, mi_tab, d_states, n_particles, d_likelihood, d_similarity, alpha);
	} else{
		e = cudaFuncSetAttribute(pf_evaluate_mi_kernel<SSM_AFF, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if(e != cudaSuccess) return e;
		pf_evaluate_mi_kernel<SSM_AFF, T><<<grid, T, smem, st>>>(b, mp, mi_tab, d_states, n_particles, d_likelihood, d_similarity, alpha);
	}
	return cudaGetLastError();
}
#endif

// the update kernels are instantiated in two translation units (this file for the homography, lk_mi_aff.cu -- which
// includes this file with MTFB_MI_AFFINE_TU -- for the affine SSM): they compile in parallel
cudaError_t launch_update_mi_aff(int sm, int threads, const DevBatch &b, const MiParams &mp, const double *mi_tab, cudaStream_t st);
#ifndef MTFB_MI_AFFINE_TU
cudaError_t launch_update_mi(int ssm, int sm, int threads, const DevBatch &b, int n_bins, double pre_seed, const double *mi_tab,
	cudaStream_t st){
	const MiParams mp = make_mi_params(b, n_bins, pre_seed);
	if(ssm == SSM_HOM){
		if(sm == SM_ESM) return launch_update_t<SSM_HOM, SM_ESM>(threads, b, mp, mi_tab, st);
		if(sm == SM_FCLK) return launch_update_t<SSM_HOM, SM_FCLK>(threads, b, mp, mi_tab, st);
		return launch_update_t<SSM_HOM, SM_ICLK>(threads, b, mp, mi_tab, st);
	}
	return launch_update_mi_aff(sm, threads, b, mp, mi_tab, st);
}
#else
cudaError_t launch_update_mi_aff(int sm, int threads, const DevBatch &b, const MiParams &mp, const double *mi_tab, cudaStream_t st){
	if(sm == SM_ESM) return launch_update_t<SSM_AFF, SM_ESM>(threads, b, mp, mi_tab, st);
	if(sm == SM_FCLK) return launch_update_t<SSM_AFF, SM_FCLK>(threads, b, mp, mi_tab, st);
	return launch_update_t<SSM_AFF, SM_ICLK>(threads, b, mp, mi_tab, st);
}
#endif

} // namespace mtfb
