// lk_ncc.cu -- the fused Lucas-Kanade kernels for the NCC appearance model (sm_100a).
//
// Reference (AM/src/NCC.cc): updateSimilarity :124-161 (centre both patches, a = <I0c, Itc>, b = |Itc|, c = |I0c|,
// f = a / (b c)), updateCurrGrad :196-234, updateInitGrad :163-194, cmptCurrJacobian / cmptInitJacobian /
// cmptDifferenceOfJacobians :236-280 and cmptSelfHessian :337-389 (H = -Jc^T Jc + (Jc^T i)(i^T Jc) with
// Jc = (dI_dp - column mean) / b and i = Itc / b).  The reference makes five to seven sweeps over N-vectors and a
// centred N x S copy of the pixel Jacobian per pass; here one pass is
//   sweep 1  warp + sample every pixel, keep It in shared memory, sum It                 -> mean
//   sweep 2  from shared memory: a, b^2, sum Itc, sum I0c                                -> f, b, both gradient means
//   sweep 3  warp + gradient + dI/dp row per pixel, accumulate  sum D, sum D D^T, sum w_t D, sum i D (, sum w_0 D0)
// and the centred products are formed from those sums:  Jc^T Jc = (sum D D^T - N m m^T) / b^2,  m = sum D / N.
// The means of the un-centred gradients (NCC.cc:179,219), which are sums of rounding residue in the reference,
// come out of the same algebra.
#include "lk_solve.cuh"

namespace mtfb {

// STD: the pass also needs the sums behind NCC::cmptCurrHessian / cmptInitHessian (NCC.cc:282-336):
// sum I0cc D (oC) for the current image, sum Itcb D0 (oB0) for the template
// ICLK_CURR: ICLK with the CurrentSelf Hessian also needs the current-image terms (NT/ICLK.cc:205-226)
template<int S, int SM, bool STD = false, bool ICLK_CURR = false> struct NccLayout {
	static constexpr int NH = S*(S + 1) / 2;
	static constexpr bool CURR = (SM != SM_ICLK) || ICLK_CURR;   // needs the current-image Jacobian terms
	static constexpr bool INIT = (SM != SM_FCLK);          // needs the template Jacobian term
	static constexpr int oD = 0, oDD = oD + (CURR ? S : 0), oW = oDD + (CURR ? NH : 0), oB = oW + (CURR ? S : 0),
		o0 = oB + (CURR ? S : 0), oC = o0 + (INIT ? S : 0), oB0 = oC + ((STD && CURR) ? S : 0), NA = oB0 + ((STD && INIT) ? S : 0);
	__host__ __device__ static constexpr int tri(int i, int j){ return i*S - i*(i - 1) / 2 + (j - i); }
};

// NCC::cmptCurrHessian (which = 0: ... + 3 u u^T) / cmptInitHessian (which = 1: ... + 3 w w^T) from sums, NCC.cc:282-336:
//   Jc = (D - column mean) / b (b = |Itc| in BOTH, the reference's quirk),  u = Jc^T Itcb,  w = Jc^T I0cc,
//   H = -f Jc^T Jc - u w^T - w u^T + 3 (u u^T | w w^T)
// sD = sum D, sDD = sum D D^T (upper), sB = sum Itcb D, sC = sum I0cc D; thread e = (i, j) of S*S
template<int S> __device__ __forceinline__ double ncc_std_hessian(int i, int j, const double *sD, const double *sDD, const double *sB,
	const double *sC, double sum_itcb, double sum_i0cc, double f, double bnorm, int N, int which){
	const int lo = i < j ? i : j, hi = i < j ? j : i;
	const double mi = sD[i] / N, mj = sD[j] / N;
	const double jcjc = ((sDD[lo*S - lo*(lo - 1) / 2 + (hi - lo)] - N*mi*mj) / bnorm) / bnorm;
	const double ui = (sB[i] - mi*sum_itcb) / bnorm, uj = (sB[j] - mj*sum_itcb) / bnorm;
	const double wi = (sC[i] - mi*sum_i0cc) / bnorm, wj = (sC[j] - mj*sum_i0cc) / bnorm;
	return -f*jcjc - ui*wj - wi*uj + 3 * (which == 0 ? ui*uj : wi*wj);
}

// self Hessian from the accumulated sums (NCC.cc:337-389): thread e = (i, j) of S*S
template<int S> __device__ __forceinline__ double ncc_self_hessian(int i, int j, const double *sD, const double *sDD,
	const double *sB, double sum_i /* sum of i = Itc / b */, double bnorm, int N){
	const int lo = i < j ? i : j, hi = i < j ? j : i;
	const double mi = sD[i] / N, mj = sD[j] / N;
	const double jcjc = ((sDD[lo*S - lo*(lo - 1) / 2 + (hi - lo)] - N*mi*mj) / bnorm) / bnorm;
	const double vi = (sB[i] - mi*sum_i) / bnorm, vj = (sB[j] - mj*sum_i) / bnorm;
	return -jcjc + vi*vj;
}

template<int SSM, int T>
__global__ void __launch_bounds__(T) ncc_init_kernel(DevBatch b, const double *__restrict__ corners_in){
	constexpr int S = StateSize<SSM>::value;
	typedef NccLayout<S, SM_FCLK> L;                        // sum D0 | sum D0 D0^T | (unused) | sum i D0
	const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	__shared__ double s_dlt[9];
	__shared__ double s_part[(T / 32) * L::NA];
	__shared__ double s_sum[L::NA];
	__shared__ double s_tmp[(T / 32) * 2];
	const double *c_in = corners_in + (size_t)p * 8;
	if(warp == 0){
		Mat3 dlt = set_corners<SSM>(b, p, lane, c_in);
		if(lane < 9) s_dlt[lane] = dlt.m[lane];
		if(lane == 0){ b.f[p] = 1; b.n_iters[p] = 0; b.status[p] = 0; }        // NCC.cc:72: f = 1 at initialisation
	}
	cta_sync<T>();
	Mat3 dlt, W = mat3_identity();
#pragma unroll
	for(int i = 0; i < 9; ++i) dlt.m[i] = s_dlt[i];
	if(b.norm_init) W = dlt;                                 // Homography.cc:57-62: curr_warp starts as the DLT warp
	// (the affine chain rule's a, b, c, d = curr_state + identity, Affine.cc:220-223: the identity unless the start is the NDLT warp)
	const double abcd[4] = { (W.m[0] - 1) + 1, W.m[1], W.m[3], (W.m[4] - 1) + 1 };
	double *I0 = b.I0 + (size_t)p*b.N, *G0 = b.G0 + (size_t)p * 2 * b.N;
	const int N = b.N;
	// phase 1: template values and chained gradient, mean
	double s1[1] = { 0 };
	for(PixIter it(tid, T, b.resx); it.pix < N; it.next(T)){
		PixGeom g = pixel_geometry<SSM>(dlt, W, b.xv[it.col], b.yv[it.row], b.norm_init != 0);
		Sample smp;
		pixel_value_and_gradient<SSM, true>(b, W, g, smp);
		const double val = smp.val;
		double J[S];
		pixel_jacobian_row<SSM>(b, W, abcd, g, smp.gx, smp.gy, J);
		I0[it.pix] = val;
		if(b.It_scratch) b.It_scratch[(size_t)p*N + it.pix] = val;          // am.initializePixVals sets It = I0 (ImageBase.cc:62-99)
		if(b.G0raw){ b.G0raw[(size_t)p * 2 * N + it.pix] = smp.gx; b.G0raw[(size_t)p * 2 * N + N + it.pix] = smp.gy; }
		G0[it.pix] = (SSM == SSM_HOM) ? J[2] : J[0];
		G0[N + it.pix] = (SSM == SSM_HOM) ? J[5] : J[1];
		if(b.I0f){
			// fp32 copies for the F32 precision (lk_ncc_f32.cu)
			b.I0f[(size_t)p*b.I0f_stride + it.pix] = (float)val;
			b.G0f[(size_t)p * 2 * N + it.pix] = (float)G0[it.pix];
			b.G0f[(size_t)p * 2 * N + N + it.pix] = (float)G0[N + it.pix];
		}
		s1[0] += val;
	}
	block_allreduce<1, T>(s1, s_tmp);
	const double I0_mean = s1[0] / N;                                       // NCC.cc:63
	// phase 2: c = | I0 - mean |
	double s2[2] = { 0, 0 };
	for(int pix = tid; pix < N; pix += T){
		const double I0c = I0[pix] - I0_mean;
		s2[0] = fma(I0c, I0c, s2[0]); s2[1] += I0c;
	}
	block_allreduce<2, T>(s2, s_tmp);
	const double c = sqrt(s2[0]), rc = ieee_rcp(c);                         // NCC.cc:65
	// phase 3: init_self_hessian = cmptSelfHessian(init_pix_jacobian) with b = c, i = I0c / c (NCC.cc:70-75,114)
	double acc[L::NA];
#pragma unroll
	for(int i = 0; i < L::NA; ++i) acc[i] = 0;
	for(PixIter it(tid, T, b.resx); it.pix < N; it.next(T)){
		PixGeom g = pixel_geometry<SSM>(dlt, W, b.xv[it.col], b.yv[it.row], b.norm_init != 0);
		double D[S];
		init_pix_jacobian<SSM>(g.ix, g.iy, G0[it.pix], G0[N + it.pix], D);
		const double I0cc = div_by(I0[it.pix] - I0_mean, c, rc);
#pragma unroll
		for(int i = 0; i < S; ++i){
			acc[L::oD + i] += D[i];
			acc[L::oB + i] = fma(I0cc, D[i], acc[L::oB + i]);
#pragma unroll
			for(int j = i; j < S; ++j) acc[L::oDD + L::tri(i, j)] = fma(D[i], D[j], acc[L::oDD + L::tri(i, j)]);
		}
	}
	block_reduce<L::NA, T>(acc, s_part, s_sum);
	for(int e = tid; e < S*S; e += T){
		const int i = e % S, j = e / S;
		b.Hinit[(size_t)p * 64 + j*S + i] = ncc_self_hessian<S>(i, j, s_sum + L::oD, s_sum + L::oDD, s_sum + L::oB, s2[1] / c, c, N);
	}
	if(tid == 0){ b.am_scal[(size_t)p * 8] = I0_mean; b.am_scal[(size_t)p * 8 + 1] = c; }
	// the template sums NCC::cmptInitHessian is built from on every pass (ICLK Std, ESM SumOfStd)
	for(int e = tid; e < S; e += T){ b.ncc_tab[(size_t)p * 64 + e] = s_sum[L::oD + e]; b.ncc_tab[(size_t)p * 64 + S + L::NH + e] = s_sum[L::oB + e]; }
	for(int e = tid; e < L::NH; e += T) b.ncc_tab[(size_t)p * 64 + S + e] = s_sum[L::oDD + e];
}

// MODE 0: self Hessians; 1 (STD): the Std forms; 2: ICLK with the CurrentSelf Hessian
template<int SSM, int SM, int T, int MODE>
__global__ void __launch_bounds__(T, min_blocks(T, 1)) ncc_update_kernel(DevBatch b){
	constexpr int S = StateSize<SSM>::value;
	constexpr bool STD = (MODE == 1);
	typedef NccLayout<S, SM, STD, MODE == 2> L;
	extern __shared__ __align__(16) double s_It[];                          // N current pixel values
	const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int N = b.N;
	__shared__ double s_part[(T / 32) * L::NA];
	__shared__ double s_sum[L::NA];
	__shared__ double s_tmp[(T / 32) * 4];
	__shared__ double s_W[9], s_dlt[9], s_corners[8], s_init_corners[8], s_J[S], s_Hc[S*S];
	__shared__ int s_ctrl;
	if(tid < 9){ s_W[tid] = b.warp[(size_t)p * 9 + tid]; s_dlt[tid] = b.dlt[(size_t)p * 9 + tid]; }
	if(tid < 8){ s_corners[tid] = b.corners[(size_t)p * 8 + tid]; s_init_corners[tid] = b.init_corners[(size_t)p * 8 + tid]; }
	cta_sync<T>();
	Mat3 dlt;
#pragma unroll
	for(int i = 0; i < 9; ++i) dlt.m[i] = s_dlt[i];
	const double *I0 = b.I0 + (size_t)p*N, *G0 = b.G0 + (size_t)p * 2 * N;
	const double I0_mean = b.am_scal[(size_t)p * 8], c = b.am_scal[(size_t)p * 8 + 1], rc = ieee_rcp(c);
	// ESM's Original variants work on mean_pix_jacobian = (init + curr) / 2 (NT/ESM.cc:246-248): the Jacobian is
	// df_dIt . mean (NT/ESM.cc:301-303, no halving), the Hessian cmptCurrHessian(mean) (NT/ESM.cc:325-327)
	const bool jac_orig = (SM == SM_ESM) && (b.jac_type == MTFB_ESM_JAC_ORIGINAL);
	const bool hess_orig = (SM == SM_ESM) && (b.hess_type == MTFB_ESM_HESS_ORIGINAL);
	const bool jac_half = (SM == SM_ESM) && !jac_orig;                       // NT/ESM.cc:308-309 (DiffOfJacs)
	LMState lm = { 0.0, b.lm_delta_init, 0.0, false };
	int iter_id = 0, n_passes = 0, patch_status = 0;
	double f = 0;
	while(iter_id < b.max_iters){
		Mat3 W;
#pragma unroll
		for(int i = 0; i < 9; ++i) W.m[i] = s_W[i];
		double abcd[4] = { (W.m[0] - 1) + 1, W.m[1], W.m[3], (W.m[4] - 1) + 1 };
		// ---- sweep 1: am.updatePixVals (ImageBase.cc:268-290) + the mean of NCC.cc:139
		double s1[1] = { 0 };
		for(PixIter it(tid, T, b.resx); it.pix < N; it.next(T)){
			PixGeom g = pixel_geometry<SSM>(dlt, W, b.xv[it.col], b.yv[it.row], b.norm_init != 0);
			const double It = sample_pixel(b.img, g.wx, g.wy);
			s_It[it.pix] = It;
			s1[0] += It;
		}
		block_allreduce<1, T>(s1, s_tmp);
		const double It_mean = s1[0] / N;
		// ---- sweep 2: a, b (NCC.cc:141-147) and the sums behind the two gradient means
		double s2[4] = { 0, 0, 0, 0 };
		for(int pix = tid; pix < N; pix += T){
			const double Itc = s_It[pix] - It_mean, I0c = I0[pix] - I0_mean;
			s2[0] = fma(I0c, Itc, s2[0]); s2[1] = fma(Itc, Itc, s2[1]); s2[2] += Itc; s2[3] += I0c;
		}
		block_allreduce<4, T>(s2, s_tmp);
		const double a = s2[0], bn = sqrt(s2[1]), rb = ieee_rcp(bn);
		f = a / (bn*c);                                                     // NCC.cc:151-152: bc = b*c, f = a / bc
		const double sum_itcb = s2[2] / bn, sum_i0cc = s2[3] / c;
		const double mean_t = ((sum_i0cc - f*sum_itcb) / bn) / N;           // mean of df_dIt_ncntr (NCC.cc:219)
		const double mean_0 = ((sum_itcb - f*sum_i0cc) / c) / N;            // mean of df_dI0_ncntr (NCC.cc:179)
		// ---- sweep 3: gradients, pixel Jacobians, sums
		double acc[L::NA];
#pragma unroll
		for(int i = 0; i < L::NA; ++i) acc[i] = 0;
		for(PixIter it(tid, T, b.resx); it.pix < N; it.next(T)){
			PixGeom g = pixel_geometry<SSM>(dlt, W, b.xv[it.col], b.yv[it.row], b.norm_init != 0);
			const double Itcb = div_by(s_It[it.pix] - It_mean, bn, rb);                  // It_cntr_b (NCC.cc:213)
			const double I0cc = div_by(I0[it.pix] - I0_mean, c, rc);                     // I0_cntr_c (NCC.cc:116)
			double D[S], D0[S];
			if(L::CURR){
				Sample smp;
				pixel_value_and_gradient<SSM, true>(b, W, g, smp);
				pixel_jacobian_row<SSM>(b, W, abcd, g, smp.gx, smp.gy, D);
			}
			if(L::INIT) init_pix_jacobian<SSM>(g.ix, g.iy, G0[it.pix], G0[N + it.pix], D0);
			if(L::CURR){
				const double wt = div_by(I0cc - f*Itcb, bn, rb) - mean_t;                // df_dIt (NCC.cc:214-222)
#pragma unroll
				for(int i = 0; i < S; ++i){
					// the row behind the Hessian sums (Dh) and the one behind the Jacobian (Dj): the current pixel Jacobian, or
					// ESM's mean pixel Jacobian
					double Dh = D[i], Dj = D[i];
					if(SM == SM_ESM && (jac_orig || hess_orig)){
						const double Dm = (D0[i] + D[i]) / 2.0;
						if(hess_orig) Dh = Dm;
						if(jac_orig) Dj = Dm;
					}
					D[i] = Dh;
					acc[L::oD + i] += Dh;
					acc[L::oW + i] = fma(wt, Dj, acc[L::oW + i]);
					acc[L::oB + i] = fma(Itcb, Dh, acc[L::oB + i]);
					if(STD) acc[L::oC + i] = fma(I0cc, Dh, acc[L::oC + i]);
				}
#pragma unroll
				for(int i = 0; i < S; ++i){
#pragma unroll
					for(int j = i; j < S; ++j) acc[L::oDD + L::tri(i, j)] = fma(D[i], D[j], acc[L::oDD + L::tri(i, j)]);
				}
			}
			if(L::INIT){
				const double w0 = div_by(Itcb - f*I0cc, c, rc) - mean_0;                 // df_dI0 (NCC.cc:174-182)
#pragma unroll
				for(int i = 0; i < S; ++i){
					acc[L::o0 + i] = fma(w0, D0[i], acc[L::o0 + i]);
					if(STD) acc[L::oB0 + i] = fma(Itcb, D0[i], acc[L::oB0 + i]);
				}
			}
		}
		block_reduce<L::NA, T>(acc, s_part, s_sum);
		++n_passes;
		for(int e = tid; e < S*S; e += T){
			const int i = e % S, j = e / S;
			if(STD){
				// FCLK / ESM Std: cmptCurrHessian; ICLK Std: cmptInitHessian; ESM SumOfStd: half their sum (NT/ESM.cc:339-352)
				const double *tab = b.ncc_tab + (size_t)p * 64;
				double hc = 0, hi = 0;
				if(L::CURR) hc = ncc_std_hessian<S>(i, j, s_sum + L::oD, s_sum + L::oDD, s_sum + L::oB, s_sum + L::oC, sum_itcb, sum_i0cc, f, bn, N, 0);
				if(L::INIT) hi = ncc_std_hessian<S>(i, j, tab, tab + S, s_sum + L::oB0, tab + S + L::NH, sum_itcb, sum_i0cc, f, bn, N, 1);
				s_Hc[j*S + i] = (SM == SM_ICLK) ? hi : (SM == SM_ESM && b.hess_type == MTFB_ESM_HESS_SUM_OF_STD) ? (hc + hi) * 0.5 : hc;
			} else if(L::CURR) s_Hc[j*S + i] = ncc_self_hessian<S>(i, j, s_sum + L::oD, s_sum + L::oDD, s_sum + L::oB, sum_itcb, bn, N);
			if(e < S){
				// FCLK: df_dIt . dIt_dp (NCC.cc:252-266); ESM: (df_dIt . dIt_dp - df_dI0 . dI0_dp) / 2 (NCC.cc:268-280,
				// NT/ESM.cc:308-309); ICLK: df_dI0 . dI0_dp (NCC.cc:236-250)
				double jv = L::CURR ? s_sum[L::oW + e] : 0.0;
				if(SM == SM_ESM && !jac_orig) jv = jv - s_sum[L::o0 + e];
				if(SM == SM_ICLK) jv = s_sum[L::o0 + e];
				s_J[e] = jac_half ? jv * 0.5 : jv;
			}
		}
		cta_sync<T>();
		if(warp == 0){
			// STD: s_Hc is the complete Hessian of the pass
			const int ctrl = STD ? serial_step<SSM, SM, false, 0>(b, p, lane, iter_id, n_passes, f, s_J, s_Hc, s_W, s_corners, s_init_corners,
				lm, patch_status) : serial_step<SSM, SM>(b, p, lane, iter_id, n_passes, f, s_J, s_Hc, s_W, s_corners, s_init_corners,
				lm, patch_status);
			if(lane == 0) s_ctrl = ctrl;
		}
		cta_sync<T>();
		const int ctrl = s_ctrl;
		if(ctrl == CTRL_BREAK) break;
		if(counts_as_iteration<SM>(ctrl, b.nt_semantics)) ++iter_id;
	}
	if(warp == 0) store_patch_state<SSM>(b, p, lane, s_W, s_corners, f, n_passes, patch_status);
	// the AM's pixel values stay those of the last pass: what setRegion's cmptSelfHessian will see (ncc_reinit_kernel)
	if(b.It_scratch) for(int pix = tid; pix < N; pix += T) b.It_scratch[(size_t)p*N + pix] = s_It[pix];
}

// ------------------------------------------------------------------------------------------------
// setRegion() of the search methods that keep template Jacobians -- nt::ESM::setRegion (NT/ESM.cc:150-168) and nt::FCLK::setRegion
// with the InitialSelf Hessian (NT/FCLK.cc:360-376) -- for NCC: ssm.setCorners, init_pix_jacobian =
// ssm.cmptInitPixJacobian(am.getInitPixGrad()) at the NEW template points, init_self_hessian = am.cmptSelfHessian(init_pix_jacobian).
// NCC::cmptSelfHessian (NCC.cc:337-389) reads the appearance model's CURRENT state -- b = |It - mean| and It_cntr_b of the
// last pass of the last update (or of initialize) -- which the update kernel leaves in It_scratch.  The template values and
// their statistics are kept; the template gradient becomes the un-chained one (G0raw: it differs from the chained one of
// initialize() by the 1 / hz of Homography.cc:68 for general quadrilaterals); the template sums behind cmptInitHessian
// (ncc_tab) follow the new Jacobian.
// ------------------------------------------------------------------------------------------------
template<int SSM, int T>
__global__ void __launch_bounds__(T) ncc_reinit_kernel(DevBatch b, const double *__restrict__ corners_in){
	constexpr int S = StateSize<SSM>::value;
	typedef NccLayout<S, SM_FCLK> L;                        // sum D0 | sum D0 D0^T | sum I0cc D0 (slot oW) | sum i D0
	const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	__shared__ double s_dlt[9];
	__shared__ double s_part[(T / 32) * L::NA];
	__shared__ double s_sum[L::NA];
	__shared__ double s_tmp[(T / 32) * 2];
	if(warp == 0){
		Mat3 dlt = set_corners<SSM>(b, p, lane, corners_in + (size_t)p * 8);
		if(lane < 9) s_dlt[lane] = dlt.m[lane];
	}
	cta_sync<T>();
	Mat3 dlt, W = mat3_identity();
#pragma unroll
	for(int i = 0; i < 9; ++i) dlt.m[i] = s_dlt[i];
	const int N = b.N;
	const double *I0 = b.I0 + (size_t)p*N, *Gr = b.G0raw + (size_t)p * 2 * N, *It = b.It_scratch + (size_t)p*N;
	double *G0 = b.G0 + (size_t)p * 2 * N;
	const double I0_mean = b.am_scal[(size_t)p * 8], c = b.am_scal[(size_t)p * 8 + 1], rc = ieee_rcp(c);
	double s1[1] = { 0 };
	for(int pix = tid; pix < N; pix += T) s1[0] += It[pix];
	block_allreduce<1, T>(s1, s_tmp);
	const double It_mean = s1[0] / N;
	double s2[2] = { 0, 0 };
	for(int pix = tid; pix < N; pix += T){
		const double Itc = It[pix] - It_mean;
		s2[0] = fma(Itc, Itc, s2[0]); s2[1] += Itc;
	}
	block_allreduce<2, T>(s2, s_tmp);
	const double bn = sqrt(s2[0]), rb = ieee_rcp(bn);
	double acc[L::NA];
#pragma unroll
	for(int i = 0; i < L::NA; ++i) acc[i] = 0;
	for(PixIter it(tid, T, b.resx); it.pix < N; it.next(T)){
		const PixGeom g = pixel_geometry<SSM>(dlt, W, b.xv[it.col], b.yv[it.row], b.norm_init != 0);   // (only the template point is used)
		double D[S];
		const double gx = Gr[it.pix], gy = Gr[N + it.pix];
		init_pix_jacobian<SSM>(g.ix, g.iy, gx, gy, D);
		G0[it.pix] = gx; G0[N + it.pix] = gy;                               // the Ix / Iy columns of cmptInitPixJacobian
		if(b.G0f){ b.G0f[(size_t)p * 2 * N + it.pix] = (float)gx; b.G0f[(size_t)p * 2 * N + N + it.pix] = (float)gy; }
		const double Itcb = div_by(It[it.pix] - It_mean, bn, rb);
		const double I0cc = div_by(I0[it.pix] - I0_mean, c, rc);
#pragma unroll
		for(int i = 0; i < S; ++i){
			acc[L::oD + i] += D[i];
			acc[L::oB + i] = fma(Itcb, D[i], acc[L::oB + i]);
			acc[L::oW + i] = fma(I0cc, D[i], acc[L::oW + i]);
#pragma unroll
			for(int j = i; j < S; ++j) acc[L::oDD + L::tri(i, j)] = fma(D[i], D[j], acc[L::oDD + L::tri(i, j)]);
		}
	}
	block_reduce<L::NA, T>(acc, s_part, s_sum);
	for(int e = tid; e < S*S; e += T){
		const int i = e % S, j = e / S;
		b.Hinit[(size_t)p * 64 + j*S + i] = ncc_self_hessian<S>(i, j, s_sum + L::oD, s_sum + L::oDD, s_sum + L::oB, s2[1] / bn, bn, N);
	}
	for(int e = tid; e < S; e += T){ b.ncc_tab[(size_t)p * 64 + e] = s_sum[L::oD + e]; b.ncc_tab[(size_t)p * 64 + S + L::NH + e] = s_sum[L::oW + e]; }
	for(int e = tid; e < L::NH; e += T) b.ncc_tab[(size_t)p * 64 + S + e] = s_sum[L::oDD + e];
}

template<int SSM> static cudaError_t launch_reinit_t(int threads, const DevBatch &b, const double *d_corners, cudaStream_t st){
	switch(threads){
	case 32: ncc_reinit_kernel<SSM, 32><<<b.P, 32, 0, st>>>(b, d_corners); break;
	case 64: ncc_reinit_kernel<SSM, 64><<<b.P, 64, 0, st>>>(b, d_corners); break;
	case 128: ncc_reinit_kernel<SSM, 128><<<b.P, 128, 0, st>>>(b, d_corners); break;
	case 256: ncc_reinit_kernel<SSM, 256><<<b.P, 256, 0, st>>>(b, d_corners); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}
cudaError_t launch_reinit_ncc(int ssm, int threads, const DevBatch &b, const double *d_corners, cudaStream_t st){
	if(ssm == SSM_HOM) return launch_reinit_t<SSM_HOM>(threads, b, d_corners, st);
	return launch_reinit_t<SSM_AFF>(threads, b, d_corners, st);
}

// ------------------------------------------------------------------------------------------------
template<int SSM> static cudaError_t launch_init_t(int threads, const DevBatch &b, const double *d_corners, cudaStream_t st){
	switch(threads){
	case 32: ncc_init_kernel<SSM, 32><<<b.P, 32, 0, st>>>(b, d_corners); break;
	case 64: ncc_init_kernel<SSM, 64><<<b.P, 64, 0, st>>>(b, d_corners); break;
	case 128: ncc_init_kernel<SSM, 128><<<b.P, 128, 0, st>>>(b, d_corners); break;
	case 256: ncc_init_kernel<SSM, 256><<<b.P, 256, 0, st>>>(b, d_corners); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}
cudaError_t launch_init_ncc(int ssm, int threads, const DevBatch &b, const double *d_corners, cudaStream_t st){
	if(ssm == SSM_HOM) return launch_init_t<SSM_HOM>(threads, b, d_corners, st);
	return launch_init_t<SSM_AFF>(threads, b, d_corners, st);
}

template<int SSM, int SM, int T, int MODE> static cudaError_t launch_one_s(const DevBatch &b, cudaStream_t st){
	const size_t smem = (size_t)b.N * sizeof(double);
	if(smem > 200 * 1024) return cudaErrorInvalidValue;
	cudaError_t e = cudaFuncSetAttribute(ncc_update_kernel<SSM, SM, T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if(e != cudaSuccess) return e;
	ncc_update_kernel<SSM, SM, T, MODE><<<b.P, T, smem, st>>>(b);
	return cudaGetLastError();
}
template<int SSM, int SM, int T> static cudaError_t launch_one(const DevBatch &b, cudaStream_t st){
	const bool std_hess = (SM == SM_ESM) ? (b.hess_type == MTFB_ESM_HESS_STD || b.hess_type == MTFB_ESM_HESS_SUM_OF_STD ||
		b.hess_type == MTFB_ESM_HESS_ORIGINAL)
		: (b.hess_type == MTFB_LK_HESS_STD);
	if(SM == SM_ICLK && b.hess_type == MTFB_LK_HESS_CURRENT_SELF) return launch_one_s<SSM, SM, T, (SM == SM_ICLK ? 2 : 0)>(b, st);
	return std_hess ? launch_one_s<SSM, SM, T, 1>(b, st) : launch_one_s<SSM, SM, T, 0>(b, st);
}
template<int SSM, int SM> static cudaError_t launch_update_t(int threads, const DevBatch &b, cudaStream_t st){
	switch(threads){
	case 32: return launch_one<SSM, SM, 32>(b, st);
	case 64: return launch_one<SSM, SM, 64>(b, st);
	case 128: return launch_one<SSM, SM, 128>(b, st);
	case 256: return launch_one<SSM, SM, 256>(b, st);
	default: return cudaErrorInvalidValue;
	}
}
cudaError_t launch_update_ncc(int ssm, int sm, int threads, const DevBatch &b, cudaStream_t st){
	if(ssm == SSM_HOM){
		if(sm == SM_ESM) return launch_update_t<SSM_HOM, SM_ESM>(threads, b, st);
		if(sm == SM_FCLK) return launch_update_t<SSM_HOM, SM_FCLK>(threads, b, st);
		return launch_update_t<SSM_HOM, SM_ICLK>(threads, b, st);
	}
	if(sm == SM_ESM) return launch_update_t<SSM_AFF, SM_ESM>(threads, b, st);
	if(sm == SM_FCLK) return launch_update_t<SSM_AFF, SM_FCLK>(threads, b, st);
	return launch_update_t<SSM_AFF, SM_ICLK>(threads, b, st);
}

} // namespace mtfb
