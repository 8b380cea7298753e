// lk_ssd_f32.cu -- the SSD Lucas-Kanade frame kernel in its "fp32 arithmetic, exact sampling indices" precision
// (mtfb_params::precision = MTFB_PRECISION_F32; BASELINE.json north_star: "bit-exact warped sampling indices,
// Jacobian / Hessian and final corner coordinates within a stated fp32 tolerance").
//
// Same path as lk_ssd.cu (nt::FCLK / ESM / ICLK::update, SM/src/NT/{FCLK.cc:171-358, ESM.cc:170-297, ICLK.cc:160-299}),
// same one-CTA-per-patch / whole-loop-in-one-launch structure, same fp64 reduction, solve and compositional update
// (lk_solve.cuh).  What changes is the per-pixel arithmetic, which moves from the fp64 pipe (64 lanes / SM on B200)
// to the fp32 pipe:
//
//  * PATCH-LOCAL COORDINATES.  fp32 cannot hold image coordinates of ~10^3 px to the 10^-5 px the sampling needs, so
//    every pass folds an INTEGER origin (X0, Y0) = floor(warped patch centre) into the warp in fp64
//    (pass_constants()), and the pixel loop evaluates  wxl = ((M0 - X0 M2).(u,v,1)) / (M2.(u,v,1))  in fp32:
//    |wxl| <= patch extent, error <= ~1e-6 * extent px (bound derived at pass_constants()).
//  * EXACT INDICES.  The reference samples at lx = (int)x (imgUtils.h:91-113).  floor(wxl) + X0 equals it whenever
//    frac(wxl) is further than delta from 0 and 1, delta = a rigorous bound on the fp32 evaluation error.  The
//    <1 % of pixels inside the guard band -- and every pixel within one cell of the image border or exactly on a
//    pixel column / row, where utils::getImgGrad's central difference straddles two cells (imgUtils.cc:233-254) --
//    take the reference-exact fp64 functions of lk_math.cuh instead.  Sampling indices are therefore bit-identical
//    to the fp64 path for every pixel; only dx, dy and everything downstream carry fp32 rounding.
//  * WELL-CONDITIONED ACCUMULATION.  With raw pixel coordinates (hom_normalized_init = 0) the columns of dI/dp are
//    almost parallel (x Ix vs x0 Ix): J^T J is numerically rank deficient even in fp64 and useless in fp32.  The
//    kernel accumulates J^T J and J^T r in a basis built on centred, scaled template coordinates
//    xl = (x - x0) / s in [-1, 1] (fp32 products, fp32 per-thread sums over <= N / T pixels), reduces in fp64 and
//    maps to the reference's basis exactly:  J_ref = J_loc . T  =>  H_ref = T^T H_loc T,  g_ref = T^T g_loc  with
//    the constant S x S matrix T(x0, y0, s) (make_basis_map()), in fp64.  The solve then sees the same matrix the
//    reference builds, to ~1e-7 relative in the well-conditioned basis.
//
// This file is compiled with -fmad=false like the rest (the fp64 slow path must not contract); fp32 fused
// multiply-adds are spelled fmaf().
#include <mutex>
#include "lk_f32.cuh"

namespace mtfb {

using namespace f32;

// MTFB_PROF = 1 (experiment builds): clock64() stamps around the phases of a pass, summed into b.n_iters_prof[8..12]
#if MTFB_PROF
#define F32_PROF_T(k) const long long prof_t##k = clock64();
#define F32_PROF_ADD() if(tid == 0){ \
	atomicAdd((unsigned long long*)b.n_iters_prof + 8, (unsigned long long)(prof_t1 - prof_t0)); \
	atomicAdd((unsigned long long*)b.n_iters_prof + 9, (unsigned long long)(prof_t2 - prof_t1)); \
	atomicAdd((unsigned long long*)b.n_iters_prof + 10, (unsigned long long)(prof_t3 - prof_t2)); \
	atomicAdd((unsigned long long*)b.n_iters_prof + 11, (unsigned long long)(prof_t4 - prof_t3)); \
	atomicAdd((unsigned long long*)b.n_iters_prof + 12, (unsigned long long)(prof_t5 - prof_t4)); }
#else
#define F32_PROF_T(k)
#define F32_PROF_ADD()
#endif

template<int SSM, int SM, int T, int MINB, int U>
__global__ void __launch_bounds__(T, MINB) ssd_update_f32_kernel(DevBatch b, unsigned smem_bytes, unsigned win_elems){
	constexpr int S = StateSize<SSM>::value;
	typedef AccLayout<S> L;
	const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	__shared__ double s_part[(T / 32) * L::NA];
	__shared__ double s_sum[L::NA];
	__shared__ double s_W[9], s_corners[8], s_init_corners[8], s_dlt[9];
	__shared__ double s_J[S], s_Hc[S*S], s_Hl[S*S], s_A[S*S], s_T[S*S], s_Tinv[S*S], s_loc[3], s_x[S], s_dp[S];
	__shared__ float s_cf[C_COUNT], s_dl[9];
	__shared__ int s_ci[2], s_wi[6];
	__shared__ int s_ctrl;
	extern __shared__ __align__(16) float s_tmpl[];
	__shared__ __align__(8) unsigned long long s_bar;
	const bool use_smem = smem_bytes != 0;             // the launcher sized the dynamic shared memory for the template or not
	if(use_smem && tid == 0) mbar_init(&s_bar, 1);
	if(tid < 9){ s_W[tid] = b.warp[(size_t)p * 9 + tid]; s_dlt[tid] = b.dlt[(size_t)p * 9 + tid]; }
	if(tid < 8){ s_corners[tid] = b.corners[(size_t)p * 8 + tid]; s_init_corners[tid] = b.init_corners[(size_t)p * 8 + tid]; }
	cta_sync<T>();
	if(use_smem && tid == 0){
		const unsigned bytes = (unsigned)b.I0f_stride * (unsigned)sizeof(float);          // stride is a multiple of 4 floats
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
	// the template, staged once per frame by one bulk async copy (TMA, SASS UBLKCP) issued before the per-patch setup;
	// templates too large for the launch's dynamic shared memory are read through L2 instead
	if(use_smem) mbar_wait(&s_bar, 0);
	const float *tmpl = use_smem ? (const float*)s_tmpl : I0;     // generic pointer: one load instruction either way
	// per-thread bit columns of the pixels deferred to the fp64 path, behind the template in dynamic shared memory
	const int slow_words = ((b.N + T - 1) / T + 31) / 32;
	unsigned *s_slow = reinterpret_cast<unsigned*>(s_tmpl) + (use_smem ? b.I0f_stride : 0);
	float *s_win = reinterpret_cast<float*>(s_slow + slow_words*T);
	const bool esm_mean = (SM == SM_ESM) && (b.jac_type == MTFB_ESM_JAC_ORIGINAL || b.hess_type == MTFB_ESM_HESS_ORIGINAL);
	const bool jac_half = (SM == SM_ESM) && (b.jac_type == MTFB_ESM_JAC_DIFF_OF_JACS);     // NT/ESM.cc:308-309
	const bool need_grad = (SM != SM_ICLK) || (b.hess_type == MTFB_LK_HESS_CURRENT_SELF);
	// the Hessian of this pass alone, no damping, nobody watching the reference-basis matrices: solve in the local basis
	const bool local_solve = b.f32_local_solve && (hessian_select<SM>(b.hess_type) == 0) && !b.leven_marq && !b.log;
	const bool lean_tail = (SM != SM_ICLK) && (hessian_select<SM>(b.hess_type) == 0) && !b.leven_marq && !b.log;
	LMState lm = { 0.0, b.lm_delta_init, 0.0, false };
	int iter_id = 0, n_passes = 0, patch_status = 0;
	double f = 0;
	// first grid position of each of this thread's U pixel streams (two integer divisions and five conversions: once per
	// frame, not once per pass)
	PixIterF it_first[U];
#pragma unroll
	for(int u = 0; u < U; ++u) it_first[u] = PixIterF(tid + u*T, U*T, b.resx);
	while(iter_id < b.max_iters){
		F32_PROF_T(0)
		PassConst k;
#pragma unroll
		for(int i = 0; i < 9; ++i) k.m[i] = s_cf[C_M + i];
#pragma unroll
		for(int i = 0; i < 6; ++i) k.a[i] = s_cf[C_A + i];
		k.delta = s_cf[C_DELTA]; k.lox = s_cf[C_LOX]; k.hix = s_cf[C_HIX]; k.loy = s_cf[C_LOY]; k.hiy = s_cf[C_HIY];
		k.X0 = s_ci[0]; k.Y0 = s_ci[1];
		k.base = b.img.data; k.pitch = b.img.pitch; k.Xr = k.X0; k.Yr = k.Y0;
		if(s_wi[2]){
			// this pass samples the shared-memory window of the frame (s_wi: window_decide)
			const int ox = s_wi[0], oy = s_wi[1];
			if(s_wi[3]){
				for(int i = tid; i < F32_WIN*F32_WIN; i += T){
					const int r = i / F32_WIN, c = i - r*F32_WIN;
					s_win[r*F32_WINP + c] = __ldg(b.img.data + (size_t)(oy + r)*b.img.pitch + ox + c);
				}
				cta_sync<T>();
			}
			k.base = s_win; k.pitch = F32_WINP; k.Xr = k.X0 - ox; k.Yr = k.Y0 - oy;
			// all four neighbours inside the window (which lies inside the frame)
			k.lox = (float)(ox - k.X0); k.hix = (float)(ox + F32_WIN - 2 - k.X0);
			k.loy = (float)(oy - k.Y0); k.hiy = (float)(oy + F32_WIN - 2 - k.Y0);
		}
		PackedAcc<S> acc;
		acc.clear();
		// U pixels per trip (pix, pix + T, ...): their dependent chains (coordinates -> loads -> chain rule) interleave.
		// The loop is branch-free: a pixel that must take the fp64 path contributes nothing here (weight zero) and sets a
		// bit in this thread's column of s_slow; the bits are worked off after the loop.  (Nearly) every pixel is slow only
		// when the warp sits on the pixel lattice -- the first pass after initialize() with an integer-aligned box.
		for(int w = 0; w < slow_words; ++w) s_slow[w*T + tid] = 0u;
		auto pixel_loop = [&](auto chk){
			constexpr bool CHK = decltype(chk)::value;
			PixIterF it[U];
#pragma unroll
			for(int u = 0; u < U; ++u) it[u] = it_first[u];
			for(int g = 0; it[0].pix < b.N; g += U){
				bool vu[U]; int pixu[U]; float rowu[U], colu[U], i0[U];
				PixF px[U];
#pragma unroll
				for(int u = 0; u < U; ++u){
					vu[u] = (u == 0) || (it[u].pix < b.N);
					pixu[u] = vu[u] ? it[u].pix : it[0].pix; rowu[u] = vu[u] ? it[u].rowf : it[0].rowf; colu[u] = vu[u] ? it[u].colf : it[0].colf;
					i0[u] = tmpl[pixu[u]];
				}
#pragma unroll
				for(int u = 0; u < U; ++u) front_fast<SSM, CHK>(b, k, dl, dlt_affine, rowu[u], colu[u], px[u]);
#pragma unroll
				for(int u = 0; u < U; ++u){
					if(vu[u] && !px[u].fast) s_slow[((g + u) >> 5)*T + tid] |= 1u << ((g + u) & 31);      // pixel tid + (g + u) T
					accumulate_pixel<SSM, SM>(b, k, px[u], i0[u], G0, pixu[u], vu[u] && px[u].fast, need_grad, esm_mean, acc);
				}
#pragma unroll
				for(int u = 0; u < U; ++u) it[u].next(U*T);
			}
		};
		// the window holds the patch's whole hull: no per-sample range test (uniform choice, once per pass)
		if(s_wi[2] && s_wi[5]) pixel_loop(std::false_type{}); else pixel_loop(std::true_type{});
		for(int w = 0; w < slow_words; ++w){
			unsigned bits = s_slow[w*T + tid];
			while(bits){
				const int g = 32 * w + __ffs(bits) - 1;
				bits &= bits - 1;
				const int pix = tid + g*T;
				PixF px;
				px.fast = false;
				const float rowf = (float)(pix / b.resx), colf = (float)(pix % b.resx);
				// template-local coordinates as the fast path computes them
				{
					const float u = fmaf(colf, b.gx_step, b.gx_lo), v = fmaf(rowf, b.gy_step, b.gy_lo);
					px.xl = fmaf(dl[0], u, fmaf(dl[1], v, dl[2]));
					px.yl = fmaf(dl[3], u, fmaf(dl[4], v, dl[5]));
					if(!dlt_affine){
						const float rz = rcp_approx(fmaf(dl[6], u, fmaf(dl[7], v, dl[8])));
						px.xl *= rz; px.yl *= rz;
					}
				}
				front_exact<SSM>(b, k, s_dlt, s_W, rowf, colf, px);
				accumulate_pixel<SSM, SM>(b, k, px, tmpl[pix], G0, pix, true, need_grad, esm_mean, acc);
			}
		}
		float accf[L::NA];
		acc.unpack(accf);
		F32_PROF_T(1)
		block_reduce_f32<L::NA, T>(accf, s_part, s_sum);
		++n_passes;
		F32_PROF_T(2)
		// pass-local Hessian, forward update, no Levenberg-Marquardt, no iteration log: warp 0 solves (local basis if asked for
		// and safe, else the reference's QR in the reference's parameters) and applies the update spread over its lanes
		if(lean_tail){
			if(warp == 0){
				bool solved = false;
				if(local_solve) solved = solve_local<S>(lane, s_sum, s_Tinv, jac_half ? 0.5 : 1.0, s_x, s_dp);
				if(!solved) solve_reference_warp<S>(lane, s_sum, s_T, jac_half ? 0.5 : 1.0, s_dp, patch_status);
				F32_PROF_T(3)
				f = -s_sum[0] / 2;
				const int ctrl = apply_update_lean<SSM>(b, lane, f, s_dp, s_W, s_corners, s_init_corners, patch_status);
				if(lane == 0) s_ctrl = ctrl;
				__syncwarp();
				F32_PROF_T(4)
				if(ctrl != CTRL_BREAK){
					pass_constants<SSM>(b, lane, s_W, s_dlt, s_loc[0], s_loc[1], s_loc[2], s_cf, s_ci);
					if(lane == 0) window_decide(b, s_corners, win_elems != 0, s_wi);
				}
				F32_PROF_T(5)
				F32_PROF_ADD()
			}
		} else{
			bool solved = false;
			if(local_solve){
				if(warp == 0) solved = solve_local<S>(lane, s_sum, s_Tinv, jac_half ? 0.5 : 1.0, s_x, s_dp);
				if(T > 32){
					if(tid == 0) s_ctrl = solved ? 1 : 0;
					__syncthreads();
					solved = s_ctrl != 0;
					__syncthreads();
				}
			}
			if(!solved){
				// local basis -> the reference's: H = T^T H_loc T, g = T^T g_loc (fp64), then the reference's QR
				for(int e = tid; e < S*S; e += T){
					const int i = e / S, m = e % S;
					s_Hl[e] = s_sum[1 + S + L::tri(i < m ? i : m, i < m ? m : i)];
				}
				cta_sync<T>();
				for(int e = tid; e < S*S; e += T){
					const int i = e / S, kk = e % S;
					double a = 0;
#pragma unroll
					for(int m = 0; m < S; ++m) a = fma(s_Hl[i*S + m], s_T[m*S + kk], a);
					s_A[e] = a;                                                     // (H_loc T)[i][kk]
				}
				cta_sync<T>();
				for(int e = tid; e < S*S; e += T){
					const int i = e % S, j = e / S;                                 // s_Hc is column-major: entry (i, j) at j*S + i
					const int lo = i < j ? i : j, hi = i < j ? j : i;
					double a = 0;
#pragma unroll
					for(int m = 0; m < S; ++m) a = fma(s_T[m*S + lo], s_A[m*S + hi], a);
					s_Hc[e] = -a;                                                   // SSD self Hessian: -J^T J (SSDBase.h:91-94)
				}
				if(tid < S){
					double a = 0;
#pragma unroll
					for(int m = 0; m < S; ++m) a = fma(s_T[m*S + tid], s_sum[1 + m], a);
					s_J[tid] = jac_half ? a * 0.5 : a;
				}
				cta_sync<T>();
			}
			F32_PROF_T(3)
			if(warp == 0){
				f = -s_sum[0] / 2;
				int ctrl;
				if(solved && SM != SM_ICLK) ctrl = apply_update_lean<SSM>(b, lane, f, s_dp, s_W, s_corners, s_init_corners, patch_status);
				else if(solved) ctrl = serial_step<SSM, SM, true>(b, p, lane, iter_id, n_passes, f, s_J, s_Hc, s_W, s_corners, s_init_corners,
					lm, patch_status, s_dp);
				else ctrl = serial_step<SSM, SM, false>(b, p, lane, iter_id, n_passes, f, s_J, s_Hc, s_W, s_corners, s_init_corners,
					lm, patch_status);
				if(lane == 0) s_ctrl = ctrl;
				__syncwarp();
				F32_PROF_T(4)
				if(ctrl != CTRL_BREAK){
					pass_constants<SSM>(b, lane, s_W, s_dlt, s_loc[0], s_loc[1], s_loc[2], s_cf, s_ci);
					if(lane == 0) window_decide(b, s_corners, win_elems != 0, s_wi);
				}
				F32_PROF_T(5)
				F32_PROF_ADD()
			}
		}
		cta_sync<T>();
		const int ctrl = s_ctrl;
		if(ctrl == CTRL_BREAK) break;
		if(counts_as_iteration<SM>(ctrl, b.nt_semantics)) ++iter_id;
	}
	if(warp == 0) store_patch_state<SSM>(b, p, lane, s_W, s_corners, f, n_passes, patch_status);
}

// ------------------------------------------------------------------------------------------------
// Particle evaluation in the F32 precision (the body of the particle loop of nt::PF::update, SM/src/NT/PF.cc:303-320:
// ssm->setState -> am->updatePixVals -> am->updateSimilarity(false) -> am->getLikelihood(), SSD.h:41-43): the same
// patch-local fp32 geometry, guard band and bit-exact sampling indices as the Gauss-Newton kernel above; a particle's
// warp takes the place of a pass's warp (pass_constants once per particle, by the warp that evaluates it).
// One CTA = one object's template (fp32 row staged by one bulk async copy) and a slice of its particles; one warp per
// particle, lanes stride over the pixels; the sum of squares is fp32 per lane (<= N / 32 terms), fp64 across lanes.
// ------------------------------------------------------------------------------------------------
template<int SSM, int T>
__global__ void __launch_bounds__(T) pf_evaluate_f32_kernel(DevBatch b, const double *__restrict__ states, int n_particles,
	double *__restrict__ likelihood, double *__restrict__ similarity, double alpha){
	constexpr int S = StateSize<SSM>::value;
	constexpr int NW = T / 32;
	const int obj = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int N = b.N;
	extern __shared__ __align__(16) float s_tmpl[];
	__shared__ __align__(8) unsigned long long s_bar;
	__shared__ double s_dlt[9], s_ic[8], s_T[S*S], s_Tinv[S*S], s_loc[3], s_Wid[9];
	__shared__ double s_W[NW][9];
	__shared__ float s_cf0[C_COUNT], s_dl[9], s_cf[NW][C_COUNT];
	__shared__ int s_ci0[2], s_ci[NW][2];
	if(tid == 0) mbar_init(&s_bar, 1);
	if(tid < 9){ s_dlt[tid] = b.dlt[(size_t)obj * 9 + tid]; s_Wid[tid] = (tid % 4 == 0) ? 1.0 : 0.0; }
	if(tid < 8) s_ic[tid] = b.init_corners[(size_t)obj * 8 + tid];
	__syncthreads();
	if(tid == 0){
		const unsigned bytes = (unsigned)b.I0f_stride * (unsigned)sizeof(float);
		mbar_expect_tx(&s_bar, bytes);
		bulk_copy_g2s(s_tmpl, b.I0f + (size_t)obj*b.I0f_stride, bytes, &s_bar);
	}
	// the object's template frame and centred DLT rows (the per-particle constants come later)
	if(warp == 0) patch_setup<SSM>(b, lane, s_Wid, s_dlt, s_ic, s_loc, s_T, s_Tinv, s_dl, s_cf0, s_ci0);
	__syncthreads();
	float dl[9];
#pragma unroll
	for(int i = 0; i < 9; ++i) dl[i] = s_dl[i];
	const bool dlt_affine = (dl[6] == 0.0f) && (dl[7] == 0.0f);
	mbar_wait(&s_bar, 0);
	for(int pi = blockIdx.x*NW + warp; pi < n_particles; pi += gridDim.x*NW){
		const double *st = states + ((size_t)obj*n_particles + pi)*S;
		// ssm->setState: curr_warp = getWarpFromState(state) (ProjectiveBase.cc:41-49, Affine.cc:109-115)
		if(lane < 9){
			const int r = lane / 3, c = lane - 3 * r;
			s_W[warp][lane] = update_entry<SSM>(st, r, c);
		}
		__syncwarp();
		pass_constants<SSM>(b, lane, s_W[warp], s_dlt, s_loc[0], s_loc[1], s_loc[2], s_cf[warp], s_ci[warp]);
		__syncwarp();
		PassConst k;
#pragma unroll
		for(int i = 0; i < 9; ++i) k.m[i] = s_cf[warp][C_M + i];
#pragma unroll
		for(int i = 0; i < 6; ++i) k.a[i] = s_cf[warp][C_A + i];
		k.delta = s_cf[warp][C_DELTA]; k.lox = s_cf[warp][C_LOX]; k.hix = s_cf[warp][C_HIX]; k.loy = s_cf[warp][C_LOY]; k.hiy = s_cf[warp][C_HIY];
		k.X0 = s_ci[warp][0]; k.Y0 = s_ci[warp][1];
		k.base = b.img.data; k.pitch = b.img.pitch; k.Xr = k.X0; k.Yr = k.Y0;
		float acc = 0.0f;
		for(PixIterF it(lane, 32, b.resx); it.pix < N; it.next(32)){
			PixF px;
			front_fast<SSM>(b, k, dl, dlt_affine, it.rowf, it.colf, px);
			if(!px.fast) front_exact<SSM>(b, k, s_dlt, s_W[warp], it.rowf, it.colf, px);
			const float d = px.val - s_tmpl[it.pix];
			acc = fmaf(d, d, acc);
		}
		double accd = (double)acc;
#pragma unroll
		for(int off = 16; off >= 1; off >>= 1) accd += __shfl_xor_sync(FULL_MASK, accd, off);
		if(lane == 0){
			const double f = -accd / 2;                                      // SSDBase.cc:94
			if(similarity) similarity[(size_t)obj*n_particles + pi] = f;
			if(likelihood) likelihood[(size_t)obj*n_particles + pi] = exp(-alpha * sqrt(-f / double(N)));
		}
		__syncwarp();
	}
}

cudaError_t launch_pf_evaluate_f32(int ssm, const DevBatch &b, const double *d_states, int n_particles, double *d_likelihood,
	double *d_similarity, double alpha, cudaStream_t st){
	constexpr int T = 256;
	const size_t smem = (size_t)b.I0f_stride*sizeof(float);
	if(smem > 160 * 1024) return cudaErrorInvalidValue;
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	// enough CTAs per object to fill the machine (4 CTAs of 8 warps per SM), never more than one particle per warp
	int per_obj = (sms * 4 + b.P - 1) / b.P;
	const int max_useful = (n_particles + T / 32 - 1) / (T / 32);
	if(per_obj > max_useful) per_obj = max_useful;
	if(per_obj < 1) per_obj = 1;
	const dim3 grid(per_obj, b.P);
	cudaError_t e;
	if(ssm == SSM_HOM){
		e = cudaFuncSetAttribute(pf_evaluate_f32_kernel<SSM_HOM, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if(e != cudaSuccess) return e;
		pf_evaluate_f32_kernel<SSM_HOM, T><<<grid, T, smem, st>>>(b, d_states, n_particles, d_likelihood, d_similarity, alpha);
	} else{
		e = cudaFuncSetAttribute(pf_evaluate_f32_kernel<SSM_AFF, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if(e != cudaSuccess) return e;
		pf_evaluate_f32_kernel<SSM_AFF, T><<<grid, T, smem, st>>>(b, d_states, n_particles, d_likelihood, d_similarity, alpha);
	}
	return cudaGetLastError();
}

// debug tap of the fp32 front end at the current state: sampling indices (lx, ly; -1 = outside the image), pixel
// value, image gradient, reference-basis Jacobian row (through the basis map) and the fast path's coordinate error
template<int SSM, int T>
__global__ void __launch_bounds__(T) ssd_stage_f32_kernel(DevBatch b, StageTapsF32 t){
	constexpr int S = StateSize<SSM>::value;
	const int p = blockIdx.x, tid = threadIdx.x;
	__shared__ double s_W[9], s_dlt[9], s_ic[8], s_T[S*S], s_Tinv[S*S], s_loc[3];
	__shared__ float s_cf[C_COUNT], s_dl[9];
	__shared__ int s_ci[2];
	if(tid < 9){ s_W[tid] = b.warp[(size_t)p * 9 + tid]; s_dlt[tid] = b.dlt[(size_t)p * 9 + tid]; }
	if(tid < 8) s_ic[tid] = b.init_corners[(size_t)p * 8 + tid];
	__syncthreads();
	if(tid < 32) patch_setup<SSM>(b, tid, s_W, s_dlt, s_ic, s_loc, s_T, s_Tinv, s_dl, s_cf, s_ci);
	__syncthreads();
	PassConst k;
	float dl[9];
	for(int i = 0; i < 9; ++i){ k.m[i] = s_cf[C_M + i]; dl[i] = s_dl[i]; }
	for(int i = 0; i < 6; ++i) k.a[i] = s_cf[C_A + i];
	k.delta = s_cf[C_DELTA]; k.lox = s_cf[C_LOX]; k.hix = s_cf[C_HIX]; k.loy = s_cf[C_LOY]; k.hiy = s_cf[C_HIY];
	k.X0 = s_ci[0]; k.Y0 = s_ci[1];
	k.base = b.img.data; k.pitch = b.img.pitch; k.Xr = k.X0; k.Yr = k.Y0;
	const bool dlt_affine = (dl[6] == 0.0f) && (dl[7] == 0.0f);
	const size_t N = b.N;
	for(PixIterF it(tid, T, b.resx); it.pix < b.N; it.next(T)){
		PixF px, ex;
		front_fast<SSM>(b, k, dl, dlt_affine, it.rowf, it.colf, px);
		ex = px;
		front_exact<SSM>(b, k, s_dlt, s_W, it.rowf, it.colf, ex);
		// (the fp64 point rounded to fp32: the comparison resolves 2^-24 of the local coordinate, a tenth of the bound)
		const float ferr = px.fast ? fmaxf(fabsf(px.wxl - ex.wxl), fabsf(px.wyl - ex.wyl)) : -1.0f;
		if(!px.fast) px = ex;
		float Gx, Gy, Jl[S];
		chain_gradient<SSM>(k, px, Gx, Gy);
		local_row<SSM>(px.xl, px.yl, Gx, Gy, Jl);
		const size_t q = p*N + it.pix;
		if(t.idx){ t.idx[2 * q] = px.lx; t.idx[2 * q + 1] = px.ly; }
		if(t.pix_vals) t.pix_vals[q] = px.val;
		if(t.pix_grad){ t.pix_grad[p * 2 * N + it.pix] = px.gx; t.pix_grad[p * 2 * N + N + it.pix] = px.gy; }
		if(t.fast_err) t.fast_err[q] = ferr;
		if(t.pix_jac){
			for(int c = 0; c < S; ++c){
				double a = 0;
				for(int m = 0; m < S; ++m) a = fma((double)Jl[m], s_T[m*S + c], a);
				t.pix_jac[(p*S + c)*N + it.pix] = a;
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
#ifndef MTFB_F32_U32
#define MTFB_F32_U32 4      // pixels per trip at one warp per patch (<= 255 registers)
#endif
#ifndef MTFB_F32_U64
#define MTFB_F32_U64 2      // ... at two or more warps per patch (<= 128 registers; 0.681 vs 0.702 ms with one pixel per trip)
#endif
#ifndef MTFB_F32_MINB128
#define MTFB_F32_MINB128 4
#endif
#ifndef MTFB_F32_MINB256
#define MTFB_F32_MINB256 2
#endif
template<int SSM, int SM, int T, int MINB, int U> static cudaError_t launch_one(const DevBatch &b, cudaStream_t st){
	// dynamic shared memory = the template (I0f row), if MINB CTAs of it still leave the SM most of its L1 for the image
	// gathers; else the kernel reads the template through L2
	size_t smem = (size_t)b.I0f_stride*sizeof(float);
	if(smem*MINB > 96 * 1024) smem = 0;
#ifdef MTFB_F32_NO_STAGE     // experiment: template through L2 instead of shared memory
	smem = 0;
#endif
	const size_t slow_bytes = (size_t)((((b.N + T - 1) / T + 31) / 32)*T) * sizeof(unsigned);
	// the frame window, if MINB CTAs with it still fit the SM's shared memory
	size_t win_bytes = (size_t)F32_WIN*F32_WINP*sizeof(float);
	if((smem + slow_bytes + win_bytes + 6 * 1024)*MINB > 216 * 1024) win_bytes = 0;
#ifdef MTFB_F32_NO_WINDOW    // experiment: gather from the frame in global memory
	win_bytes = 0;
#endif
	// per instantiation and device; attributes are per function and sticky: the carve-out follows the largest footprint seen
	static size_t configured_dev[64] = {};
	static std::mutex mu;                            // contexts of different host threads may share an instantiation
	std::lock_guard<std::mutex> lock(mu);
	int dev = 0;
	cudaGetDevice(&dev);
	size_t &configured = configured_dev[dev & 63];
	cudaFuncAttributes fa;
	cudaError_t e = cudaFuncGetAttributes(&fa, ssd_update_f32_kernel<SSM, SM, T, MINB, U>);
	if(e != cudaSuccess) return e;
	// shared memory carve-out: what MINB resident CTAs need (static + dynamic + 1 KB the system reserves per CTA);
	// the rest of the SM's 228 KB stays L1 for the image gathers.  Too small a carve-out would cost resident CTAs.
	const size_t per_cta = fa.sharedSizeBytes + (size_t)b.I0f_stride*sizeof(float) + slow_bytes + (size_t)F32_WIN*F32_WINP*sizeof(float) + 1024;
	if(per_cta > configured){
		e = cudaFuncSetAttribute(ssd_update_f32_kernel<SSM, SM, T, MINB, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024 / MINB + 32 * 1024);
		if(e != cudaSuccess) return e;
		const int carve = (int)((per_cta*MINB * 100 + 228 * 1024 - 1) / (228 * 1024)) + 2;
		e = cudaFuncSetAttribute(ssd_update_f32_kernel<SSM, SM, T, MINB, U>, cudaFuncAttributePreferredSharedMemoryCarveout, carve > 100 ? 100 : carve);
		if(e != cudaSuccess) return e;
		configured = per_cta;
	}
	if(slow_bytes > 16 * 1024) return cudaErrorInvalidValue;      // > 4 M pixels per patch
	ssd_update_f32_kernel<SSM, SM, T, MINB, U><<<b.P, T, smem + slow_bytes + win_bytes, st>>>(b, (unsigned)smem, (unsigned)(win_bytes / sizeof(float)));
	return cudaGetLastError();
}
template<int SSM, int SM> static cudaError_t launch_update_f32_t(int threads, const DevBatch &b, cudaStream_t st){
	switch(threads){
	case 32: return launch_one<SSM, SM, 32, 8, MTFB_F32_U32>(b, st);
	case 64: return launch_one<SSM, SM, 64, 7, MTFB_F32_U64>(b, st);
	case 128: return launch_one<SSM, SM, 128, MTFB_F32_MINB128, MTFB_F32_U64>(b, st);
	case 256: return launch_one<SSM, SM, 256, MTFB_F32_MINB256, MTFB_F32_U64>(b, st);
	default: return cudaErrorInvalidValue;
	}
}
cudaError_t launch_update_ssd_f32(int ssm, int sm, int threads, const DevBatch &b, cudaStream_t st){
#ifdef MTFB_ONLY_FCLK_HOM
	if(ssm == SSM_HOM && sm == SM_FCLK) return launch_update_f32_t<SSM_HOM, SM_FCLK>(threads, b, st);
	return cudaErrorNotSupported;
#else
	if(ssm == SSM_HOM){
		if(sm == SM_ESM) return launch_update_f32_t<SSM_HOM, SM_ESM>(threads, b, st);
		if(sm == SM_FCLK) return launch_update_f32_t<SSM_HOM, SM_FCLK>(threads, b, st);
		return launch_update_f32_t<SSM_HOM, SM_ICLK>(threads, b, st);
	}
	if(sm == SM_ESM) return launch_update_f32_t<SSM_AFF, SM_ESM>(threads, b, st);
	if(sm == SM_FCLK) return launch_update_f32_t<SSM_AFF, SM_FCLK>(threads, b, st);
	return launch_update_f32_t<SSM_AFF, SM_ICLK>(threads, b, st);
#endif
}
cudaError_t launch_stage_f32(int ssm, const DevBatch &b, const StageTapsF32 &t, cudaStream_t st){
	if(ssm == SSM_HOM) ssd_stage_f32_kernel<SSM_HOM, 64><<<b.P, 64, 0, st>>>(b, t);
	else ssd_stage_f32_kernel<SSM_AFF, 64><<<b.P, 64, 0, st>>>(b, t);
	return cudaGetLastError();
}

} // namespace mtfb
