// lk_ssd.cu -- the fused Lucas-Kanade kernels for the SSD appearance model (sm_100a), plus the
// AM-independent set-region and stage-tap kernels.
//
// One CTA tracks one patch for a whole frame: the <= max_iters Gauss-Newton loop of
// nt::FCLK::update (SM/src/NT/FCLK.cc:171-358), nt::ESM::update (SM/src/NT/ESM.cc:170-297) and
// nt::ICLK::update (SM/src/NT/ICLK.cc:160-299) runs inside a single launch.  Per pass every thread
// walks its pixels once and does, in registers, what the reference does in eight sweeps over
// materialised N-vectors / N x S matrices:
//   ssm.getPts            warp the grid point                    (lk_math.cuh pixel_geometry)
//   am.updatePixVals      bilinear sample                        (sample_pixel*)
//   am.updatePixGrad      image gradient at the warped point     (sample_pixel_grad)
//   ssm.cmptWarpedPixJacobian   dI/dp row                        (warped_pix_jacobian)
//   am.updateSimilarity / updateCurrGrad / cmptCurrJacobian / cmptSelfHessian
//                         f, J^T r and J^T J accumulated per thread in fp64
// then the CTA reduces the 1 + S + S(S+1)/2 sums (warp butterfly + one shared-memory hop) and warp 0
// runs the Levenberg-Marquardt bookkeeping, the S x S column-pivoted QR solve, the compositional
// update and the corner-change stopping test.
#include "lk_solve.cuh"
#include "lk_ssd_terms.cuh"

namespace mtfb {

// ------------------------------------------------------------------------------------------------
// initialize(): ssm.setCorners + am.initializePixVals + initializePixGrad + cmptWarpedPixJacobian (at the
// identity warp) + am.cmptSelfHessian  (NT/FCLK.cc:102-169, NT/ESM.cc:110-146, NT/ICLK.cc:71-127)
// ------------------------------------------------------------------------------------------------
template<int SSM, int T>
__global__ void __launch_bounds__(T) ssd_init_kernel(DevBatch b, const double *__restrict__ corners_in){
	constexpr int S = StateSize<SSM>::value;
	typedef AccLayout<S> L;
	const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	__shared__ double s_dlt[9];
	__shared__ double s_part[(T / 32) * L::NA];
	__shared__ double s_sum[L::NA];
	const double *c_in = corners_in + (size_t)p * 8;
	if(warp == 0){
		Mat3 dlt = set_corners<SSM>(b, p, lane, c_in);
		if(lane < 9) s_dlt[lane] = dlt.m[lane];
		if(lane == 0){ b.f[p] = 0; b.n_iters[p] = 0; b.status[p] = 0; }
	}
	cta_sync<T>();
	Mat3 dlt, W = mat3_identity();
#pragma unroll
	for(int i = 0; i < 9; ++i) dlt.m[i] = s_dlt[i];
	if(b.norm_init) W = dlt;                                 // Homography.cc:57-62: curr_warp starts as the DLT warp
	// (the affine chain rule's a, b, c, d = curr_state + identity, Affine.cc:220-223: the identity unless the start is the NDLT warp)
	const double abcd[4] = { (W.m[0] - 1) + 1, W.m[1], W.m[3], (W.m[4] - 1) + 1 };
	double acc[L::NA];
#pragma unroll
	for(int i = 0; i < L::NA; ++i) acc[i] = 0;
	double *I0 = b.I0 + (size_t)p*b.N, *G0 = b.G0 + (size_t)p * 2 * b.N;
	for(PixIter it(tid, T, b.resx); it.pix < b.N; it.next(T)){
		PixGeom g = pixel_geometry<SSM>(dlt, W, b.xv[it.col], b.yv[it.row], b.norm_init != 0);
		Sample smp;
		pixel_value_and_gradient<SSM, true>(b, W, g, smp);
		double J[S];
		pixel_jacobian_row<SSM>(b, W, abcd, g, smp.gx, smp.gy, J);
		I0[it.pix] = smp.val;
		// the template gradient as the Jacobian uses it: what cmptWarpedPixJacobian / cmptInitPixJacobian left in the
		// Ix / Iy columns
		G0[it.pix] = (SSM == SSM_HOM) ? J[2] : J[0];
		G0[b.N + it.pix] = (SSM == SSM_HOM) ? J[S > 5 ? 5 : 1] : J[1];
		if(b.G0raw){ b.G0raw[(size_t)p * 2 * b.N + it.pix] = smp.gx; b.G0raw[(size_t)p * 2 * b.N + b.N + it.pix] = smp.gy; }
		if(b.I0f){
			// fp32 copies for the fp32-arithmetic update kernel (lk_ssd_f32.cu)
			b.I0f[(size_t)p*b.I0f_stride + it.pix] = (float)smp.val;
			b.G0f[(size_t)p * 2 * b.N + it.pix] = (float)G0[it.pix];
			b.G0f[(size_t)p * 2 * b.N + b.N + it.pix] = (float)G0[b.N + it.pix];
		}
		// nt::FALK / nt::IALK::initialize (NT/FALK.cc:108-118, NT/IALK.cc:70-80): init_pix_jacobian = ssm.cmptPixJacobian(am.getInitPixGrad())
		// at the START state -- the same row as above from an identity start, not from a normalised one
		if(b.additive_sm) additive_pix_jacobian<SSM>(g, smp.gx, smp.gy, J);
#pragma unroll
		for(int i = 0; i < S; ++i){
#pragma unroll
			for(int j = i; j < S; ++j) acc[1 + S + L::tri(i, j)] = fma(J[i], J[j], acc[1 + S + L::tri(i, j)]);
		}
	}
	block_reduce<L::NA, T>(acc, s_part, s_sum);
	// (strided: T = 32 threads write the 64 entries of the homography's Hessian in two trips)
	for(int e = tid; e < S*S; e += T){
		const int i = e % S, j = e / S;
		const int lo = i < j ? i : j, hi = i < j ? j : i;
		b.Hinit[(size_t)p * 64 + j*S + i] = -s_sum[1 + S + L::tri(lo, hi)];        // SSD self Hessian: -J^T J
	}
}

// ------------------------------------------------------------------------------------------------
// setRegion() of the search methods that keep template Jacobians -- nt::ESM::setRegion (NT/ESM.cc:150-168) and
// nt::FCLK::setRegion with the InitialSelf Hessian (NT/FCLK.cc:360-376): ssm.setCorners, then
// init_pix_jacobian = ssm.cmptInitPixJacobian(am.getInitPixGrad()) at the NEW template points (always the un-chained
// form, whatever chained_warp says: the reference's choice) and init_self_hessian = am.cmptSelfHessian(init_pix_jacobian).
// The template values are kept.
// ------------------------------------------------------------------------------------------------
template<int SSM, int T>
__global__ void __launch_bounds__(T) ssd_reinit_kernel(DevBatch b, const double *__restrict__ corners_in){
	constexpr int S = StateSize<SSM>::value;
	typedef AccLayout<S> L;
	const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	__shared__ double s_dlt[9];
	__shared__ double s_part[(T / 32) * L::NA];
	__shared__ double s_sum[L::NA];
	if(warp == 0){
		Mat3 dlt = set_corners<SSM>(b, p, lane, corners_in + (size_t)p * 8);
		if(lane < 9) s_dlt[lane] = dlt.m[lane];
	}
	cta_sync<T>();
	Mat3 dlt, W = mat3_identity();
#pragma unroll
	for(int i = 0; i < 9; ++i) dlt.m[i] = s_dlt[i];
	if(b.norm_init) W = dlt;
	double acc[L::NA];
#pragma unroll
	for(int i = 0; i < L::NA; ++i) acc[i] = 0;
	double *G0 = b.G0 + (size_t)p * 2 * b.N;
	const double *Gr = b.G0raw + (size_t)p * 2 * b.N;
	for(PixIter it(tid, T, b.resx); it.pix < b.N; it.next(T)){
		const PixGeom g = pixel_geometry<SSM>(dlt, W, b.xv[it.col], b.yv[it.row], b.norm_init != 0);
		const double gx = Gr[it.pix], gy = Gr[b.N + it.pix];
		double J[S];
		init_pix_jacobian<SSM>(g.ix, g.iy, gx, gy, J);
		G0[it.pix] = gx; G0[b.N + it.pix] = gy;                              // the Ix / Iy columns of cmptInitPixJacobian
		if(b.G0f){ b.G0f[(size_t)p * 2 * b.N + it.pix] = (float)gx; b.G0f[(size_t)p * 2 * b.N + b.N + it.pix] = (float)gy; }
#pragma unroll
		for(int i = 0; i < S; ++i){
#pragma unroll
			for(int j = i; j < S; ++j) acc[1 + S + L::tri(i, j)] = fma(J[i], J[j], acc[1 + S + L::tri(i, j)]);
		}
	}
	block_reduce<L::NA, T>(acc, s_part, s_sum);
	for(int e = tid; e < S*S; e += T){
		const int i = e % S, j = e / S;
		const int lo = i < j ? i : j, hi = i < j ? j : i;
		b.Hinit[(size_t)p * 64 + j*S + i] = -s_sum[1 + S + L::tri(lo, hi)];
	}
}

// ------------------------------------------------------------------------------------------------
// update(): the whole per-frame loop
// ------------------------------------------------------------------------------------------------
// MTFB_PROF = 1: clock64() stamps around the phases of a pass, summed into b.am_scal[0..] of patch 0 (experiment builds)
#if MTFB_PROF
#define MTFB_PROF_T(k) const long long prof_t##k = clock64();
#define MTFB_PROF_ADD() if(tid == 0){ atomicAdd((unsigned long long*)b.n_iters_prof + 0, (unsigned long long)(prof_t1 - prof_t0)); \
	atomicAdd((unsigned long long*)b.n_iters_prof + 1, (unsigned long long)(prof_t2 - prof_t1)); atomicAdd((unsigned long long*)b.n_iters_prof + 2, (unsigned long long)(prof_t3 - prof_t2)); }
#else
#define MTFB_PROF_T(k)
#define MTFB_PROF_ADD()
#endif

template<int SSM, int SM, int T, int OCC>
__global__ void __launch_bounds__(T, min_blocks(T, OCC)) ssd_update_kernel(DevBatch b){
	constexpr int S = StateSize<SSM>::value;
	typedef AccLayout<S> L;
	const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	__shared__ double s_part[(T / 32) * L::NA];
	__shared__ double s_sum[L::NA];
	__shared__ double s_W[9], s_corners[8], s_init_corners[8];
	__shared__ double s_J[S], s_Hc[S*S];
	__shared__ int s_ctrl;
	__shared__ double s_dlt[9];
	constexpr bool ADDITIVE = (SM == SM_FALK || SM == SM_IALK);
	__shared__ double s_state[ADDITIVE ? S : 1];
	if(tid < 9){ s_W[tid] = b.warp[(size_t)p * 9 + tid]; s_dlt[tid] = b.dlt[(size_t)p * 9 + tid]; }
	if(tid < 8){ s_corners[tid] = b.corners[(size_t)p * 8 + tid]; s_init_corners[tid] = b.init_corners[(size_t)p * 8 + tid]; }
	if(ADDITIVE && tid < S) s_state[tid] = b.state[(size_t)p*S + tid];
	cta_sync<T>();
	Mat3 dlt;
#pragma unroll
	for(int i = 0; i < 9; ++i) dlt.m[i] = s_dlt[i];
	const double *I0 = b.I0 + (size_t)p*b.N, *G0 = b.G0 + (size_t)p * 2 * b.N;
	const bool esm_mean = (SM == SM_ESM) && (b.jac_type == MTFB_ESM_JAC_ORIGINAL || b.hess_type == MTFB_ESM_HESS_ORIGINAL);
	const bool jac_half = (SM == SM_ESM) && (b.jac_type == MTFB_ESM_JAC_DIFF_OF_JACS);     // NT/ESM.cc:308-309
	// need_grad: the pass needs a pixel Jacobian of the CURRENT patch (the gradient of the current image, or for IALK the
	// template's gradient pushed through cmptApproxPixJacobian); with_hess: it also rebuilds the Hessian from it
	const bool need_grad = (SM != SM_ICLK) || (b.hess_type == MTFB_LK_HESS_CURRENT_SELF);
	const bool with_hess = ADDITIVE ? (b.hess_type != MTFB_LK_HESS_INITIAL_SELF) : need_grad;
	const bool sample_grad = need_grad && (SM != SM_IALK);
	LMState lm = { 0.0, b.lm_delta_init, 0.0, false };          // uniform; only warp 0's copy is used
	int iter_id = 0, n_passes = 0, patch_status = 0;
	double f = 0;
	while(iter_id < b.max_iters){
		MTFB_PROF_T(0)
		double abcd[4] = { 0, 0, 0, 0 };
		if(SSM == SSM_AFF){
			// Affine.cc:217-220 reads curr_state(2)+1, (3), (4), (5)+1 with curr_state = getStateFromWarp(curr_warp)
			abcd[0] = (s_W[0] - 1) + 1; abcd[1] = s_W[1]; abcd[2] = s_W[3]; abcd[3] = (s_W[4] - 1) + 1;
			if(ADDITIVE){ abcd[0] = s_state[2] + 1; abcd[1] = s_state[3]; abcd[2] = s_state[4]; abcd[3] = s_state[5] + 1; }
		}
		double acc[L::NA];
#pragma unroll
		for(int i = 0; i < L::NA; ++i) acc[i] = 0;
		Mat3 Wm;
#pragma unroll
		for(int i = 0; i < 9; ++i) Wm.m[i] = s_W[i];
		for(PixIter it(tid, T, b.resx); it.pix < b.N; it.next(T)){
			// streamed once per pass: cached in L2 only, L1 is left to the four image gathers per pixel
			const double i0 = __ldcg(I0 + it.pix);
			const PixGeom g = pixel_geometry<SSM>(dlt, Wm, b.xv[it.col], b.yv[it.row], b.norm_init != 0);
			Sample smp;
			if(sample_grad){ pixel_value_and_gradient<SSM, true>(b, Wm, g, smp); }
			else{ smp.val = sample_pixel(b.img, g.wx, g.wy); smp.gx = smp.gy = 0; smp.lit = 0; }
			PixTerms<S> t;
			pixel_terms<SSM, SM>(b, Wm, abcd, g, smp, i0, G0, it.pix, need_grad, esm_mean, t);
			accumulate_terms<S>(acc, t, with_hess);
		}
		MTFB_PROF_T(1)
		block_reduce<L::NA, T>(acc, s_part, s_sum);
		MTFB_PROF_T(2)
		++n_passes;
		// SSD: f = -sum r^2 / 2 (SSDBase.cc:94), df_dp = sum df_dI * dI_dp, self Hessian = -J^T J (SSDBase.h:91-94)
		if(tid < S*S){
			const int i = tid % S, j = tid / S;
			const int lo = i < j ? i : j, hi = i < j ? j : i;
			s_Hc[j*S + i] = -s_sum[1 + S + L::tri(lo, hi)];
			if(tid < S) s_J[tid] = jac_half ? s_sum[1 + tid] * 0.5 : s_sum[1 + tid];
		}
		if(T == 32 && S*S > 32){
			for(int e = tid + 32; e < S*S; e += 32){
				const int i = e % S, j = e / S;
				const int lo = i < j ? i : j, hi = i < j ? j : i;
				s_Hc[j*S + i] = -s_sum[1 + S + L::tri(lo, hi)];
			}
		}
		cta_sync<T>();
		if(warp == 0){
			f = -s_sum[0] / 2;
			const int ctrl = serial_step<SSM, SM>(b, p, lane, iter_id, n_passes, f, s_J, s_Hc, s_W, s_corners, s_init_corners,
				lm, patch_status, nullptr, ADDITIVE ? s_state : nullptr);
			if(lane == 0) s_ctrl = ctrl;
		}
		cta_sync<T>();
		MTFB_PROF_T(3)
		MTFB_PROF_ADD()
		const int ctrl = s_ctrl;
		if(ctrl == CTRL_BREAK) break;
		if(counts_as_iteration<SM>(ctrl, b.nt_semantics)) ++iter_id;
	}
	if(warp == 0) store_patch_state<SSM>(b, p, lane, s_W, s_corners, f, n_passes, patch_status, ADDITIVE ? s_state : nullptr);
}

// ------------------------------------------------------------------------------------------------
// setRegion(): ssm.setCorners only -- new DLT, identity warp; the template is kept
// (NT/FCLK.cc:360-376 with CurrentSelf Hessian, NT/ICLK.cc setRegion with update_ssm = 0)
// ------------------------------------------------------------------------------------------------
__global__ void lk_set_region_kernel(DevBatch b, const double *__restrict__ corners_in, int S){
	const int p = blockIdx.x, lane = threadIdx.x;
	const double *c_in = corners_in + (size_t)p * 8;
	if(S == 8) set_corners<SSM_HOM>(b, p, lane, c_in); else if(S == 6) set_corners<SSM_AFF>(b, p, lane, c_in); else set_corners<SSM_TRANS>(b, p, lane, c_in);
}

// ------------------------------------------------------------------------------------------------
// stage taps: pts / It / dIt_dx / dIt_dp at the current state, through the same device functions
// ------------------------------------------------------------------------------------------------
template<int SSM, int T>
__global__ void __launch_bounds__(T) lk_stage_kernel(DevBatch b, StageTaps t){
	constexpr int S = StateSize<SSM>::value;
	const int p = blockIdx.x, tid = threadIdx.x;
	Mat3 dlt, W;
#pragma unroll
	for(int i = 0; i < 9; ++i){ dlt.m[i] = b.dlt[(size_t)p * 9 + i]; W.m[i] = b.warp[(size_t)p * 9 + i]; }
	double abcd[4] = { W.m[0], W.m[1], W.m[3], W.m[4] };
	if(SSM == SSM_AFF){ abcd[0] = (W.m[0] - 1) + 1; abcd[3] = (W.m[4] - 1) + 1; }
	const size_t N = b.N;
	for(PixIter it(tid, T, b.resx); it.pix < b.N; it.next(T)){
		PixGeom g = pixel_geometry<SSM>(dlt, W, b.xv[it.col], b.yv[it.row], b.norm_init != 0);
		Sample smp;
		pixel_value_and_gradient<SSM, false>(b, W, g, smp);
		const double val = b.pix_mult*smp.val + b.pix_add, gx = smp.gx, gy = smp.gy;
		double J[S];
		pixel_jacobian_row<SSM>(b, W, abcd, g, gx, gy, J);
		if(t.pts){ t.pts[(p*N + it.pix) * 2] = g.wx; t.pts[(p*N + it.pix) * 2 + 1] = g.wy; }
		if(t.pix_vals) t.pix_vals[p*N + it.pix] = val;
		if(t.pix_grad){ t.pix_grad[p * 2 * N + it.pix] = gx; t.pix_grad[p * 2 * N + N + it.pix] = gy; }
		if(t.pix_jac){
#pragma unroll
			for(int s = 0; s < S; ++s) t.pix_jac[(p*S + s)*N + it.pix] = J[s];
		}
	}
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
template<int SSM> static cudaError_t launch_init_t(int threads, const DevBatch &b, const double *d_corners, cudaStream_t st){
	switch(threads){
	case 32: ssd_init_kernel<SSM, 32><<<b.P, 32, 0, st>>>(b, d_corners); break;
	case 64: ssd_init_kernel<SSM, 64><<<b.P, 64, 0, st>>>(b, d_corners); break;
	case 128: ssd_init_kernel<SSM, 128><<<b.P, 128, 0, st>>>(b, d_corners); break;
	case 256: ssd_init_kernel<SSM, 256><<<b.P, 256, 0, st>>>(b, d_corners); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}
cudaError_t launch_init_ssd(int ssm, int threads, const DevBatch &b, const double *d_corners, cudaStream_t st){
	if(ssm == SSM_HOM) return launch_init_t<SSM_HOM>(threads, b, d_corners, st);
	if(ssm == SSM_TRANS) return launch_init_t<SSM_TRANS>(threads, b, d_corners, st);
	return launch_init_t<SSM_AFF>(threads, b, d_corners, st);
}

template<int SSM> static cudaError_t launch_reinit_t(int threads, const DevBatch &b, const double *d_corners, cudaStream_t st){
	switch(threads){
	case 32: ssd_reinit_kernel<SSM, 32><<<b.P, 32, 0, st>>>(b, d_corners); break;
	case 64: ssd_reinit_kernel<SSM, 64><<<b.P, 64, 0, st>>>(b, d_corners); break;
	case 128: ssd_reinit_kernel<SSM, 128><<<b.P, 128, 0, st>>>(b, d_corners); break;
	case 256: ssd_reinit_kernel<SSM, 256><<<b.P, 256, 0, st>>>(b, d_corners); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}
cudaError_t launch_reinit_ssd(int ssm, int threads, const DevBatch &b, const double *d_corners, cudaStream_t st){
	if(!b.G0raw) return cudaErrorInvalidValue;
	if(ssm == SSM_HOM) return launch_reinit_t<SSM_HOM>(threads, b, d_corners, st);
	if(ssm == SSM_TRANS) return launch_reinit_t<SSM_TRANS>(threads, b, d_corners, st);
	return launch_reinit_t<SSM_AFF>(threads, b, d_corners, st);
}

cudaError_t launch_set_region(int ssm, const DevBatch &b, const double *d_corners, cudaStream_t st){
	lk_set_region_kernel<<<b.P, 32, 0, st>>>(b, d_corners, ssm == SSM_HOM ? 8 : (ssm == SSM_AFF ? 6 : 2));
	return cudaGetLastError();
}

template<int SSM, int SM, int OCC> static cudaError_t launch_update_o(int threads, const DevBatch &b, cudaStream_t st){
	switch(threads){
	case 32: ssd_update_kernel<SSM, SM, 32, OCC><<<b.P, 32, 0, st>>>(b); break;
	case 64: ssd_update_kernel<SSM, SM, 64, OCC><<<b.P, 64, 0, st>>>(b); break;
	case 128: ssd_update_kernel<SSM, SM, 128, OCC><<<b.P, 128, 0, st>>>(b); break;
	case 256: ssd_update_kernel<SSM, SM, 256, OCC><<<b.P, 256, 0, st>>>(b); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}
template<int SSM, int SM> static cudaError_t launch_update_t(int threads, int occ, const DevBatch &b, cudaStream_t st){
	if(occ == 0) return launch_update_o<SSM, SM, 0>(threads, b, st);
	if(occ == 1) return launch_update_o<SSM, SM, 1>(threads, b, st);
	return launch_update_o<SSM, SM, 2>(threads, b, st);
}
cudaError_t launch_update_ssd(int ssm, int sm, int threads, int occ, const DevBatch &b, cudaStream_t st){
#ifdef MTFB_ONLY_FCLK_HOM      // experiment builds (profiles/): one combination, fast to compile
	if(ssm == SSM_HOM && sm == SM_FCLK) return launch_update_t<SSM_HOM, SM_FCLK>(threads, occ, b, st);
	return cudaErrorNotSupported;
#else
	if(ssm == SSM_HOM){
		if(sm == SM_ESM) return launch_update_t<SSM_HOM, SM_ESM>(threads, occ, b, st);
		if(sm == SM_FCLK) return launch_update_t<SSM_HOM, SM_FCLK>(threads, occ, b, st);
		if(sm == SM_FALK) return launch_update_o<SSM_HOM, SM_FALK, 0>(threads, b, st);
		if(sm == SM_IALK) return launch_update_o<SSM_HOM, SM_IALK, 0>(threads, b, st);
		return launch_update_t<SSM_HOM, SM_ICLK>(threads, occ, b, st);
	}
	if(ssm == SSM_TRANS){
		// the translation SSM (GridTracker's default cell model, parameters.h:502): two sums + three Hessian entries per pixel,
		// one register budget
		if(sm == SM_ESM) return launch_update_o<SSM_TRANS, SM_ESM, 2>(threads, b, st);
		if(sm == SM_FCLK) return launch_update_o<SSM_TRANS, SM_FCLK, 2>(threads, b, st);
		if(sm == SM_FALK) return launch_update_o<SSM_TRANS, SM_FALK, 2>(threads, b, st);
		if(sm == SM_IALK) return launch_update_o<SSM_TRANS, SM_IALK, 2>(threads, b, st);
		return launch_update_o<SSM_TRANS, SM_ICLK, 2>(threads, b, st);
	}
	if(sm == SM_ESM) return launch_update_t<SSM_AFF, SM_ESM>(threads, occ, b, st);
	if(sm == SM_FCLK) return launch_update_t<SSM_AFF, SM_FCLK>(threads, occ, b, st);
	if(sm == SM_FALK) return launch_update_o<SSM_AFF, SM_FALK, 0>(threads, b, st);
	if(sm == SM_IALK) return launch_update_o<SSM_AFF, SM_IALK, 0>(threads, b, st);
	return launch_update_t<SSM_AFF, SM_ICLK>(threads, occ, b, st);
#endif
}

template<int SSM> static cudaError_t launch_stage_t(int threads, const DevBatch &b, const StageTaps &t, cudaStream_t st){
	switch(threads){
	case 32: lk_stage_kernel<SSM, 32><<<b.P, 32, 0, st>>>(b, t); break;
	case 64: lk_stage_kernel<SSM, 64><<<b.P, 64, 0, st>>>(b, t); break;
	case 128: lk_stage_kernel<SSM, 128><<<b.P, 128, 0, st>>>(b, t); break;
	case 256: lk_stage_kernel<SSM, 256><<<b.P, 256, 0, st>>>(b, t); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}
cudaError_t launch_stage(int ssm, int threads, const DevBatch &b, const StageTaps &t, cudaStream_t st){
	if(ssm == SSM_HOM) return launch_stage_t<SSM_HOM>(threads, b, t, st);
	if(ssm == SSM_TRANS) return launch_stage_t<SSM_TRANS>(threads, b, t, st);
	return launch_stage_t<SSM_AFF>(threads, b, t, st);
}

} // namespace mtfb
