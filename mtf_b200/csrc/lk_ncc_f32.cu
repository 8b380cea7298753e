// lk_ncc_f32.cu -- NCC (AM/src/NCC.cc) under ESM / FCLK in the F32 precision: ONE fp32 sweep per pass.
//
// The fp64 kernel (lk_ncc.cu) follows the reference's order -- pixel values, then their mean, then a = <I0c, Itc> and
// b = |Itc|, then the gradient weights (which need f = a / bc) times the pixel Jacobians: three sweeps over the pixels per
// pass.  But every sum of a pass is LINEAR in {1, It, I0} x {rows of the current pixel Jacobian D, rows of the template's D0}
// once the pass's scalars are known:
//     sum Itcb D      = (sum It D - mean_t sum D) / b                                         (NCC.cc:213)
//     sum df_dIt D    = ((sum I0 D - mean_0 sum D) / c - f sum Itcb D) / b                    (NCC.cc:214-222, 252-266)
//     sum df_dI0 D0   = ((sum It D0 - mean_t sum D0) / b - f (sum I0 D0 - mean_0 sum D0) / c) / c   (NCC.cc:174-182, 236-250)
//     self Hessian    = -(Jc^T Jc) + (Jc^T Itcb)(Jc^T Itcb)^T,  Jc = (D - column mean) / b    (NCC.cc:337-389)
// (the two "mean of the gradient" terms of NCC.cc:179,219 are sums of centred values: rounding noise, dropped), so one sweep
// accumulates  sum It, It^2, I0 It, I0 | sum D | sum It D | sum I0 D | sum It D0 | sum D D^T  and the fp64 tail assembles
// f, the Jacobian and the Hessian from them.  sum D0 and sum I0 D0 are constants of the template: one light sweep per frame.
//
// The per-pixel arithmetic is the F32 precision's (lk_f32.cuh, shared with lk_ssd_f32.cu): patch-local coordinates, bit-exact
// sampling indices through the fp64 guard band (deferred pixels are re-evaluated by the reference-exact functions), sums in
// the centred / scaled local basis mapped to the reference's parameters in fp64 (H_ref = T^T H_loc T, J_ref = T^T J_loc).
// Pixel values enter the products relative to the template's mean (|It - mean_0| ~ contrast, not ~ 128): no cancellation in
// b^2 = sum It'^2 - N mean'^2.  The solve is the reference's column-pivoted QR (serial_step), Levenberg-Marquardt and the
// stored Hessians (init_self_hessian from the fp64 init kernel) included.
#include <mutex>
#include "lk_f32.cuh"

namespace mtfb {

using namespace f32;

template<int S, bool ESM> struct NccF32Layout {
	static constexpr int NH = S*(S + 1) / 2;
	static constexpr int oD = 4, oItD = oD + S, oI0D = oItD + S, oItD0 = oI0D + S, oDD = oItD0 + (ESM ? S : 0), NA = oDD + NH;
	__host__ __device__ static constexpr int tri(int i, int j){ return i*S - i*(i - 1) / 2 + (j - i); }   // i <= j
};

// self Hessian from the sums (NCC.cc:337-389), entry (i, j): the same algebra as lk_ncc.cu ncc_self_hessian
template<int S> __device__ __forceinline__ double ncc_f32_self_hessian(int i, int j, const double *sD, const double *sDD,
	const double *sB, double bnorm, int N){
	const int lo = i < j ? i : j, hi = i < j ? j : i;
	const double mi = sD[i] / N, mj = sD[j] / N;
	const double jcjc = ((sDD[lo*S - lo*(lo - 1) / 2 + (hi - lo)] - N*mi*mj) / bnorm) / bnorm;
	const double vi = sB[i] / bnorm, vj = sB[j] / bnorm;       // Jc^T Itcb = (sum Itcb D - mean_D sum Itcb) / b, sum Itcb = 0
	return -jcjc + vi*vj;
}

// The solve of a pass by ONE warp, for the tails without Levenberg-Marquardt and without an iteration log (the counterpart of
// solve_reference_warp, lk_f32.cuh): lane j < S builds column j of T^T H_loc T in registers, combines it with the stored
// init_self_hessian as the search method's Hessian type says (hsel 0: this pass's, 1: the stored one, 2: their mean --
// serial_step's rule), lane S carries J_ref = T^T J_loc; the reference's column-pivoted QR (factor_lean) and back-substitution.
template<int S> __device__ __forceinline__ void solve_stored_warp(int lane, const double *s_Hl, const double *s_Jl, const double *s_T,
	const double *Hinit, int hsel, double *s_dp, int &patch_status){
	const int j = lane < S ? lane : 0;
	double tj[S], A[S];
#pragma unroll
	for(int n = 0; n < S; ++n) tj[n] = s_T[n*S + j];
#pragma unroll
	for(int m = 0; m < S; ++m){
		double a0 = 0, a1 = 0;
#pragma unroll
		for(int n = 0; n < S; n += 2){ a0 = fma(s_Hl[m*S + n], tj[n], a0); a1 = fma(s_Hl[m*S + n + 1], tj[n + 1], a1); }
		A[m] = a0 + a1;                                                   // (H_loc T)[m][j]
	}
	if(lane == S){
#pragma unroll
		for(int m = 0; m < S; ++m) A[m] = s_Jl[m];
	}
	WarpColPivQR<S, S> qr;
#pragma unroll
	for(int i = 0; i < S; ++i){
		double a0 = 0, a1 = 0;
#pragma unroll
		for(int m = 0; m < S; m += 2){ a0 = fma(s_T[m*S + i], A[m], a0); a1 = fma(s_T[(m + 1)*S + i], A[m + 1], a1); }
		double v = a0 + a1;
		if(lane < S && hsel != 0){
			const double hi0 = Hinit[j*S + i];
			v = hsel == 1 ? hi0 : (v + hi0) * 0.5;
		}
		qr.a[i] = v;
	}
	qr.factor_lean(lane, true);
	const double x = -qr.solve_fast_cols(lane);                          // state_update = -H^-1 J^T (NT/ESM.cc:266, NT/FCLK.cc:298)
	if(qr.nonzero_pivots < S) patch_status |= MTFB_PATCH_SINGULAR;
	if(lane < S) s_dp[lane] = x;
	__syncwarp();
}

template<int SSM, bool ESM> struct NccF32Acc {
	static constexpr int S = StateSize<SSM>::value;
	typedef NccF32Layout<S, ESM> L;
	float a[L::NA];
	__device__ __forceinline__ void clear(){
#pragma unroll
		for(int i = 0; i < L::NA; ++i) a[i] = 0.0f;
	}
	// it, i0: pixel values relative to the template's mean; D, D0: local-basis rows (zero for a pixel that adds nothing)
	__device__ __forceinline__ void add(float it, float i0, const float *D, const float *D0){
		a[0] += it; a[1] = fmaf(it, it, a[1]); a[2] = fmaf(i0, it, a[2]); a[3] += i0;
#pragma unroll
		for(int s = 0; s < S; ++s){
			a[L::oD + s] += D[s];
			a[L::oItD + s] = fmaf(it, D[s], a[L::oItD + s]);
			a[L::oI0D + s] = fmaf(i0, D[s], a[L::oI0D + s]);
			if(ESM) a[L::oItD0 + s] = fmaf(it, D0[s], a[L::oItD0 + s]);
		}
#pragma unroll
		for(int i = 0; i < S; ++i){
#pragma unroll
			for(int j = i; j < S; ++j) a[L::oDD + L::tri(i, j)] = fmaf(D[i], D[j], a[L::oDD + L::tri(i, j)]);
		}
	}
};

template<int SSM, int SM, int T, int MINB>
__global__ void __launch_bounds__(T, MINB) ncc_update_f32_kernel(DevBatch b, unsigned smem_bytes, unsigned win_elems){
	constexpr int S = StateSize<SSM>::value;
	constexpr bool ESM = (SM == SM_ESM);
	typedef NccF32Layout<S, ESM> L;
	const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	__shared__ double s_part[(T / 32) * L::NA];
	__shared__ double s_sum[L::NA];
	__shared__ double s_t0[2 * S];                     // template constants in the local basis: sum D0 | sum I0' D0
	__shared__ double s_W[9], s_corners[8], s_init_corners[8], s_dlt[9];
	__shared__ double s_J[S], s_Jl[S], s_Hc[S*S], s_Hl[S*S], s_A[S*S], s_T[S*S], s_Tinv[S*S], s_loc[3], s_B[S];
	__shared__ double s_scal[4];                       // f, bn, mean_t', (spare)
	__shared__ double s_dp[S];
	__shared__ float s_cf[C_COUNT], s_dl[9];
	__shared__ int s_ci[2], s_wi[6];
	__shared__ int s_ctrl;
	extern __shared__ __align__(16) float s_tmpl[];
	__shared__ __align__(8) unsigned long long s_bar;
	const bool use_smem = smem_bytes != 0;
	if(use_smem && tid == 0) mbar_init(&s_bar, 1);
	if(tid < 9){ s_W[tid] = b.warp[(size_t)p * 9 + tid]; s_dlt[tid] = b.dlt[(size_t)p * 9 + tid]; }
	if(tid < 8){ s_corners[tid] = b.corners[(size_t)p * 8 + tid]; s_init_corners[tid] = b.init_corners[(size_t)p * 8 + tid]; }
	cta_sync<T>();
	if(use_smem && tid == 0){
		const unsigned bytes = (unsigned)b.I0f_stride * (unsigned)sizeof(float);
		mbar_expect_tx(&s_bar, bytes);
		bulk_copy_g2s(s_tmpl, b.I0f + (size_t)p*b.I0f_stride, bytes, &s_bar);
	}
	if(warp == 0){
		patch_setup<SSM>(b, lane, s_W, s_dlt, s_init_corners, s_loc, s_T, s_Tinv, s_dl, s_cf, s_ci);
		if(lane == 0){ s_wi[4] = 0; window_decide(b, s_corners, win_elems != 0, s_wi); }
	}
	cta_sync<T>();
	float dl[9];
#pragma unroll
	for(int i = 0; i < 9; ++i) dl[i] = s_dl[i];
	const bool dlt_affine = (dl[6] == 0.0f) && (dl[7] == 0.0f);
	const float *I0 = b.I0f + (size_t)p*b.I0f_stride, *G0 = b.G0f + (size_t)p * 2 * b.N;
	if(use_smem) mbar_wait(&s_bar, 0);
	const float *tmpl = use_smem ? (const float*)s_tmpl : I0;
	const int slow_words = ((b.N + T - 1) / T + 31) / 32;
	unsigned *s_slow = reinterpret_cast<unsigned*>(s_tmpl) + (use_smem ? b.I0f_stride : 0);
	float *s_win = reinterpret_cast<float*>(s_slow + slow_words*T);
	const int N = b.N;
	const double m0 = b.am_scal[(size_t)p * 8], c = b.am_scal[(size_t)p * 8 + 1];      // NCC.cc:63-65: I0_mean, |I0 - mean|
	const float m0f = (float)m0;
	const double delta0 = m0 - (double)m0f;            // I0 - mean_0 = (I0 - m0f) - delta0
	const bool jac_half = ESM && (b.jac_type == MTFB_ESM_JAC_DIFF_OF_JACS);
	const double rN = 1.0 / N, rc = 1.0 / c;
	// template-local coordinates of a grid position, as front_fast computes them
	auto local_xy = [&](float rowf, float colf, float &xl, float &yl){
		const float u = fmaf(colf, b.gx_step, b.gx_lo), v = fmaf(rowf, b.gy_step, b.gy_lo);
		xl = fmaf(dl[0], u, fmaf(dl[1], v, dl[2]));
		yl = fmaf(dl[3], u, fmaf(dl[4], v, dl[5]));
		if(!dlt_affine){
			const float rz = rcp_approx(fmaf(dl[6], u, fmaf(dl[7], v, dl[8])));
			xl *= rz; yl *= rz;
		}
	};
	if(ESM){
		// the template's own sums, once per frame: sum D0, sum I0' D0 (local basis)
		float t0[2 * S];
#pragma unroll
		for(int i = 0; i < 2 * S; ++i) t0[i] = 0.0f;
		for(PixIterF it(tid, T, b.resx); it.pix < N; it.next(T)){
			float xl, yl, D0[S];
			local_xy(it.rowf, it.colf, xl, yl);
			local_row<SSM>(xl, yl, __ldcg(G0 + it.pix), __ldcg(G0 + N + it.pix), D0);
			const float i0 = tmpl[it.pix] - m0f;
#pragma unroll
			for(int s = 0; s < S; ++s){ t0[s] += D0[s]; t0[S + s] = fmaf(i0, D0[s], t0[S + s]); }
		}
		block_reduce_f32<2 * S, T>(t0, s_part, s_t0);
	}
	LMState lm = { 0.0, b.lm_delta_init, 0.0, false };
	int iter_id = 0, n_passes = 0, patch_status = 0;
	double f = 0;
	while(iter_id < b.max_iters){
		PassConst k;
#pragma unroll
		for(int i = 0; i < 9; ++i) k.m[i] = s_cf[C_M + i];
#pragma unroll
		for(int i = 0; i < 6; ++i) k.a[i] = s_cf[C_A + i];
		k.delta = s_cf[C_DELTA]; k.lox = s_cf[C_LOX]; k.hix = s_cf[C_HIX]; k.loy = s_cf[C_LOY]; k.hiy = s_cf[C_HIY];
		k.X0 = s_ci[0]; k.Y0 = s_ci[1];
		k.base = b.img.data; k.pitch = b.img.pitch; k.Xr = k.X0; k.Yr = k.Y0;
		if(s_wi[2]){
			const int ox = s_wi[0], oy = s_wi[1];
			if(s_wi[3]){
				for(int i = tid; i < F32_WIN*F32_WIN; i += T){
					const int r = i / F32_WIN, cc = i - r*F32_WIN;
					s_win[r*F32_WINP + cc] = __ldg(b.img.data + (size_t)(oy + r)*b.img.pitch + ox + cc);
				}
				cta_sync<T>();
			}
			k.base = s_win; k.pitch = F32_WINP; k.Xr = k.X0 - ox; k.Yr = k.Y0 - oy;
			k.lox = (float)(ox - k.X0); k.hix = (float)(ox + F32_WIN - 2 - k.X0);
			k.loy = (float)(oy - k.Y0); k.hiy = (float)(oy + F32_WIN - 2 - k.Y0);
		}
		NccF32Acc<SSM, ESM> acc;
		acc.clear();
		for(int w = 0; w < slow_words; ++w) s_slow[w*T + tid] = 0u;
		auto one_pixel = [&](const PixF &px, int pix, bool valid){
			float Gx, Gy, D[S], D0[S];
			chain_gradient<SSM>(k, px, Gx, Gy);
			if(!valid){ Gx = 0; Gy = 0; }
			local_row<SSM>(px.xl, px.yl, Gx, Gy, D);
			if(ESM){
				float g0x = __ldcg(G0 + pix), g0y = __ldcg(G0 + N + pix);
				if(!valid){ g0x = 0; g0y = 0; }
				local_row<SSM>(px.xl, px.yl, g0x, g0y, D0);
			}
			const float it = valid ? px.val - m0f : 0.0f, i0 = valid ? tmpl[pix] - m0f : 0.0f;
			acc.add(it, i0, D, D0);
		};
		{
			int g = 0;
			for(PixIterF it(tid, T, b.resx); it.pix < N; it.next(T), ++g){
				PixF px;
				front_fast<SSM, true>(b, k, dl, dlt_affine, it.rowf, it.colf, px);
				if(!px.fast) s_slow[(g >> 5)*T + tid] |= 1u << (g & 31);
				one_pixel(px, it.pix, px.fast);
			}
		}
		for(int w = 0; w < slow_words; ++w){
			unsigned bits = s_slow[w*T + tid];
			while(bits){
				const int g = 32 * w + __ffs(bits) - 1;
				bits &= bits - 1;
				const int pix = tid + g*T;
				PixF px;
				px.fast = false;
				const float rowf = (float)(pix / b.resx), colf = (float)(pix % b.resx);
				local_xy(rowf, colf, px.xl, px.yl);
				front_exact<SSM>(b, k, s_dlt, s_W, rowf, colf, px);
				one_pixel(px, pix, true);
			}
		}
		block_reduce_f32<L::NA, T>(acc.a, s_part, s_sum);
		++n_passes;
		// ---- the pass's scalars and vectors (fp64, local basis)
		const bool lean_tail = !b.leven_marq && !b.log;
		if(lean_tail){
			// everything by warp 0, no block barrier until the pass ends
			if(warp == 0){
				// (reciprocals instead of the ten IEEE divisions of the straightforward form: each costs ~20 dependent fp64
				// instructions on this warp's critical path, and the sums carry fp32 rounding anyway)
				const double S_it = s_sum[0], S_it2 = s_sum[1], S_i0it = s_sum[2], S_i0 = s_sum[3];
				const double mt = S_it * rN;                              // mean of It' = It - m0f
				const double bn2 = S_it2 - S_it*mt;
				const double rb = rsqrt_newton(bn2), bn = bn2 * rb;      // |It - mean|          (NCC.cc:145-147)
				const double fv = (S_i0it - mt*S_i0) * rb * rc;           // f = <I0c, Itc> / bc  (NCC.cc:141, 151-152)
				f = fv;
				if(lane < S){
					const double sD = s_sum[L::oD + lane];
					const double sB = (s_sum[L::oItD + lane] - mt*sD) * rb;
					const double sC = (s_sum[L::oI0D + lane] - delta0*sD) * rc;
					double jv = (sC - fv*sB) * rb;
					if(ESM){
						const double sD0 = s_t0[lane];
						jv -= ((s_sum[L::oItD0 + lane] - mt*sD0) * rb - fv*(s_t0[S + lane] - delta0*sD0) * rc) * rc;
					}
					s_B[lane] = sB;
					s_Jl[lane] = jac_half ? jv * 0.5 : jv;
				}
				(void)bn;
				__syncwarp();
				if(lane < S){
					const double rb2 = rb*rb, mj = s_sum[L::oD + lane] * rN, vj = s_B[lane];
#pragma unroll
					for(int i = 0; i < S; ++i){
						const int lo = i < lane ? i : lane, hi = i < lane ? lane : i;
						const double mi = s_sum[L::oD + i] * rN;
						s_Hl[i*S + lane] = (vj*s_B[i] - (s_sum[L::oDD + L::tri(lo, hi)] - N*mi*mj)) * rb2;     // NCC.cc:337-389
					}
				}
				__syncwarp();
				solve_stored_warp<S>(lane, s_Hl, s_Jl, s_T, b.Hinit + (size_t)p * 64, hessian_select<SM>(b.hess_type), s_dp, patch_status);
				const int ctrl = apply_update_lean<SSM>(b, lane, f, s_dp, s_W, s_corners, s_init_corners, patch_status);
				if(lane == 0) s_ctrl = ctrl;
				__syncwarp();
				if(ctrl != CTRL_BREAK){
					pass_constants<SSM>(b, lane, s_W, s_dlt, s_loc[0], s_loc[1], s_loc[2], s_cf, s_ci);
					if(lane == 0) window_decide(b, s_corners, win_elems != 0, s_wi);
				}
			}
		} else{
		if(tid == 0){
			const double S_it = s_sum[0], S_it2 = s_sum[1], S_i0it = s_sum[2], S_i0 = s_sum[3];
			const double mt = S_it / N;                                   // mean of It' = It - m0f
			const double bn = sqrt(S_it2 - S_it*mt);                      // |It - mean|          (NCC.cc:145-147)
			const double a = S_i0it - mt*S_i0;                            // <I0c, Itc>           (NCC.cc:141)
			s_scal[0] = a / (bn*c); s_scal[1] = bn; s_scal[2] = mt;       // f = a / bc           (NCC.cc:151-152)
		}
		cta_sync<T>();
		const double fv = s_scal[0], bn = s_scal[1], mt = s_scal[2];
		f = fv;
		if(tid < S){
			const double sD = s_sum[L::oD + tid];
			const double sB = (s_sum[L::oItD + tid] - mt*sD) / bn;         // sum Itcb D
			const double sC = (s_sum[L::oI0D + tid] - delta0*sD) / c;      // sum I0cc D
			double jv = (sC - fv*sB) / bn;                                 // df_dIt . dIt_dp
			if(ESM){
				const double sD0 = s_t0[tid];
				const double j0 = ((s_sum[L::oItD0 + tid] - mt*sD0) / bn - fv*(s_t0[S + tid] - delta0*sD0) / c) / c;   // df_dI0 . dI0_dp
				jv = jv - j0;                                              // NCC.cc:268-280
			}
			s_B[tid] = sB;
			s_Jl[tid] = jac_half ? jv * 0.5 : jv;                          // NT/ESM.cc:308-309
		}
		cta_sync<T>();
		for(int e = tid; e < S*S; e += T){
			const int i = e / S, j = e % S;
			s_Hl[e] = ncc_f32_self_hessian<S>(i, j, s_sum + L::oD, s_sum + L::oDD, s_B, bn, N);
		}
		cta_sync<T>();
		// local basis -> the reference's: H = T^T H_loc T, J = T^T J_loc
		for(int e = tid; e < S*S; e += T){
			const int i = e / S, kk = e % S;
			double a = 0;
#pragma unroll
			for(int m = 0; m < S; ++m) a = fma(s_Hl[i*S + m], s_T[m*S + kk], a);
			s_A[e] = a;
		}
		cta_sync<T>();
		for(int e = tid; e < S*S; e += T){
			const int i = e % S, j = e / S;                                 // s_Hc is column-major
			const int lo = i < j ? i : j, hi = i < j ? j : i;
			double a = 0;
#pragma unroll
			for(int m = 0; m < S; ++m) a = fma(s_T[m*S + lo], s_A[m*S + hi], a);
			s_Hc[e] = a;
		}
		if(tid < S){
			double a = 0;
#pragma unroll
			for(int m = 0; m < S; ++m) a = fma(s_T[m*S + tid], s_Jl[m], a);
			s_J[tid] = a;
		}
		cta_sync<T>();
		if(warp == 0){
			const int ctrl = serial_step<SSM, SM>(b, p, lane, iter_id, n_passes, f, s_J, s_Hc, s_W, s_corners, s_init_corners, lm, patch_status);
			if(lane == 0) s_ctrl = ctrl;
			__syncwarp();
			if(ctrl != CTRL_BREAK){
				pass_constants<SSM>(b, lane, s_W, s_dlt, s_loc[0], s_loc[1], s_loc[2], s_cf, s_ci);
				if(lane == 0) window_decide(b, s_corners, win_elems != 0, s_wi);
			}
		}
		}
		cta_sync<T>();
		const int ctrl = s_ctrl;
		if(ctrl == CTRL_BREAK) break;
		if(counts_as_iteration<SM>(ctrl, b.nt_semantics)) ++iter_id;
	}
	if(warp == 0) store_patch_state<SSM>(b, p, lane, s_W, s_corners, f, n_passes, patch_status);
}

// ------------------------------------------------------------------------------------------------ launchers
template<int SSM, int SM, int T, int MINB> static cudaError_t launch_one(const DevBatch &b, cudaStream_t st){
	size_t smem = (size_t)b.I0f_stride*sizeof(float);
	if(smem*MINB > 96 * 1024) smem = 0;
	const size_t slow_bytes = (size_t)((((b.N + T - 1) / T + 31) / 32)*T) * sizeof(unsigned);
	size_t win_bytes = (size_t)F32_WIN*F32_WINP*sizeof(float);
	static size_t configured_dev[64] = {};
	static std::mutex mu;
	std::lock_guard<std::mutex> lock(mu);
	int dev = 0;
	cudaGetDevice(&dev);
	size_t &configured = configured_dev[dev & 63];
	cudaFuncAttributes fa;
	cudaError_t e = cudaFuncGetAttributes(&fa, ncc_update_f32_kernel<SSM, SM, T, MINB>);
	if(e != cudaSuccess) return e;
	if((fa.sharedSizeBytes + smem + slow_bytes + win_bytes + 1024)*MINB > 224 * 1024) win_bytes = 0;
	const size_t per_cta = fa.sharedSizeBytes + (size_t)b.I0f_stride*sizeof(float) + slow_bytes + (size_t)F32_WIN*F32_WINP*sizeof(float) + 1024;
	if(per_cta > configured){
		e = cudaFuncSetAttribute(ncc_update_f32_kernel<SSM, SM, T, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024 / MINB + 32 * 1024);
		if(e != cudaSuccess) return e;
		const int carve = (int)((per_cta*MINB * 100 + 228 * 1024 - 1) / (228 * 1024)) + 2;
		e = cudaFuncSetAttribute(ncc_update_f32_kernel<SSM, SM, T, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, carve > 100 ? 100 : carve);
		if(e != cudaSuccess) return e;
		configured = per_cta;
	}
	if(slow_bytes > 16 * 1024) return cudaErrorInvalidValue;
	ncc_update_f32_kernel<SSM, SM, T, MINB><<<b.P, T, smem + slow_bytes + win_bytes, st>>>(b, (unsigned)smem, (unsigned)(win_bytes / sizeof(float)));
	return cudaGetLastError();
}
template<int SSM, int SM> static cudaError_t launch_t(int threads, const DevBatch &b, cudaStream_t st){
	switch(threads){
	case 32: return launch_one<SSM, SM, 32, 8>(b, st);
	case 64: return launch_one<SSM, SM, 64, 7>(b, st);
	case 128: return launch_one<SSM, SM, 128, 4>(b, st);
	case 256: return launch_one<SSM, SM, 256, 2>(b, st);
	default: return cudaErrorInvalidValue;
	}
}
cudaError_t launch_update_ncc_f32(int ssm, int sm, int threads, const DevBatch &b, cudaStream_t st){
	if(sm != SM_ESM && sm != SM_FCLK) return cudaErrorInvalidValue;
	if(ssm == SSM_HOM) return sm == SM_ESM ? launch_t<SSM_HOM, SM_ESM>(threads, b, st) : launch_t<SSM_HOM, SM_FCLK>(threads, b, st);
	return sm == SM_ESM ? launch_t<SSM_AFF, SM_ESM>(threads, b, st) : launch_t<SSM_AFF, SM_FCLK>(threads, b, st);
}

} // namespace mtfb
