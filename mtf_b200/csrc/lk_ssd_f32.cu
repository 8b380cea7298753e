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
#include "lk_solve.cuh"

namespace mtfb {

namespace f32 {

// indices into the per-pass constant block (floats in shared memory, copied to registers by every thread)
enum { C_M = 0,            // 9: Homography: rows of diag(1,1,1) . [M0 - X0 M2; M1 - Y0 M2; M2], M = curr_warp . dlt
                           //    Affine: 6: (W00 s, W01 s, W00 x0 + W01 y0 + W02 - X0, W10 s, W11 s, ... - Y0)
       C_A = 9,            // 6: Homography: a00 - a20 X0, a01 - a21 X0, a10 - a20 Y0, a11 - a21 Y0, a20, a21
                           //    Affine: a, b, c, d of Affine.cc:217-220
       C_DELTA = 15, C_LOX = 16, C_HIX = 17, C_LOY = 18, C_HIY = 19, C_COUNT = 20 };

struct PassConst {
	float m[9], a[6];
	float delta, lox, hix, loy, hiy;
	int X0, Y0;
};

__device__ __forceinline__ float rcp_approx(float x){
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));       // MUFU.RCP, <= 1 ulp
	return r;
}

// Centre and scale of the template points, from the corners of the initial region: x0, y0 = mean corner, s = the
// largest |corner - centre| coordinate.  Any choice works (the basis map is exact algebra); this one puts the
// template points into [-1, 1]^2.  hom_normalized_init: the template points are the unit grid itself.
template<int SSM> __device__ __forceinline__ void template_frame(const DevBatch &b, const double *init_corners, double &x0, double &y0,
	double &s){
	if(SSM == SSM_HOM && b.norm_init){ x0 = 0; y0 = 0; s = 0.5; return; }
	x0 = 0.25*(init_corners[0] + init_corners[1] + init_corners[2] + init_corners[3]);
	y0 = 0.25*(init_corners[4] + init_corners[5] + init_corners[6] + init_corners[7]);
	s = 0;
#pragma unroll
	for(int i = 0; i < 4; ++i){ s = fmax(s, fabs(init_corners[i] - x0)); s = fmax(s, fabs(init_corners[4 + i] - y0)); }
	if(!(s > 0)) s = 1;
}

// T: the S x S matrix with  J_ref[k] = sum_m J_loc[m] T[m][k]  where J_ref = init_pix_jacobian<SSM>(x, y, Gx, Gy)
// (the reference's dI/dp row, Homography.cc:231-294 / Affine.cc:213-242 after chaining the gradient) and
// J_loc = init_pix_jacobian<SSM>(xl, yl, Gx, Gy), x = x0 + s xl, y = y0 + s yl.  Row-major T[m*S + k].
template<int SSM> __device__ __forceinline__ void make_basis_map(double x0, double y0, double s, double *T){
	constexpr int S = StateSize<SSM>::value;
	for(int i = 0; i < S*S; ++i) T[i] = 0;
#define TT(m, k) T[(m)*S + (k)]
	if(SSM == SSM_HOM){
		// J_loc = [Gx xl, Gx yl, Gx, Gy xl, Gy yl, Gy, -xl q, -yl q],  q = xl Gx + yl Gy
		TT(0, 0) = s; TT(2, 0) = x0;                       // Gx x
		TT(1, 1) = s; TT(2, 1) = y0;                       // Gx y
		TT(2, 2) = 1;
		TT(3, 3) = s; TT(5, 3) = x0;                       // Gy x
		TT(4, 4) = s; TT(5, 4) = y0;                       // Gy y
		TT(5, 5) = 1;
		// -x (x Gx + y Gy)
		TT(6, 6) = s*s; TT(0, 6) = -2 * s*x0; TT(4, 6) = -s*x0; TT(3, 6) = -s*y0; TT(2, 6) = -x0*x0; TT(5, 6) = -x0*y0;
		// -y (x Gx + y Gy)
		TT(7, 7) = s*s; TT(0, 7) = -s*y0; TT(4, 7) = -2 * s*y0; TT(1, 7) = -s*x0; TT(2, 7) = -x0*y0; TT(5, 7) = -y0*y0;
	} else{
		// J_loc = [Gx, Gy, Gx xl, Gx yl, Gy xl, Gy yl]
		TT(0, 0) = 1; TT(1, 1) = 1;
		TT(2, 2) = s; TT(0, 2) = x0;
		TT(3, 3) = s; TT(0, 3) = y0;
		TT(4, 4) = s; TT(1, 4) = x0;
		TT(5, 5) = s; TT(1, 5) = y0;
	}
#undef TT
}

// Per-pass constants, by one thread, in fp64 (about 80 operations per pass and patch).
//
// Error bound behind delta (Homography; u, v in [-1/2, 1/2]).  wxl = num / den with num = c0 u + c1 v + c2 evaluated as
// two fp32 fmaf's on coefficients and grid values rounded to fp32: every term carries <= 2 roundings of 2^-24 and each
// fmaf one more on its partial sum, so |err(num)| <= 4 . 2^-24 . A with A = |c0|/2 + |c1|/2 + |c2|, and likewise
// |err(den)| <= 4 . 2^-24 . B; MUFU.RCP (1 ulp) and the final product add 2 . 2^-24 relative.  With Dmin <= |den|,
// E = A / Dmin >= |wxl| and rho = B / Dmin >= 1:   |err(wxl)| <= 2^-24 . E . (6 + 4 rho)   (= 6e-7 E at rho = 1).
// delta = 2.5 x that bound + 2e-6 px (tests/test_gpu_parity.py measures the actual error through the tap).
// Affine: wxl = c0 xl + c1 yl + c2 with |xl|, |yl| <= 1 known to 10 . 2^-24 (their own quotient): E = |c0| + |c1| + |c2|,
// |err| <= (10 + 4) . 2^-24 . E: the same formula with rho = 2.
template<int SSM> __device__ __forceinline__ void pass_constants(const DevBatch &b, const double *W, const double *dlt, double x0, double y0,
	double s, float *cf, int *ci){
	double E, rho = 2;
	int X0 = 0, Y0 = 0;
	bool sane = true;
	if(SSM == SSM_HOM){
		double M[9];
		if(b.norm_init){
#pragma unroll
			for(int i = 0; i < 9; ++i) M[i] = W[i];
		} else{
#pragma unroll
			for(int r = 0; r < 3; ++r)
#pragma unroll
			for(int c = 0; c < 3; ++c) M[3 * r + c] = W[3 * r] * dlt[c] + W[3 * r + 1] * dlt[3 + c] + W[3 * r + 2] * dlt[6 + c];
		}
		const double cx = M[2] / M[8], cy = M[5] / M[8];
		sane = (fabs(cx) < 1e8) && (fabs(cy) < 1e8);
		if(sane){ X0 = (int)floor(cx); Y0 = (int)floor(cy); }
#pragma unroll
		for(int c = 0; c < 3; ++c){ M[c] -= X0*M[6 + c]; M[3 + c] -= Y0*M[6 + c]; }
#pragma unroll
		for(int i = 0; i < 9; ++i) cf[C_M + i] = (float)M[i];
		cf[C_A + 0] = (float)(W[0] - W[6] * X0); cf[C_A + 1] = (float)(W[1] - W[7] * X0);
		cf[C_A + 2] = (float)(W[3] - W[6] * Y0); cf[C_A + 3] = (float)(W[4] - W[7] * Y0);
		cf[C_A + 4] = (float)W[6]; cf[C_A + 5] = (float)W[7];
		const double Ax = 0.5*fabs(M[0]) + 0.5*fabs(M[1]) + fabs(M[2]), Ay = 0.5*fabs(M[3]) + 0.5*fabs(M[4]) + fabs(M[5]);
		const double Dmin = fabs(M[8]) - 0.5*fabs(M[6]) - 0.5*fabs(M[7]);
		E = fmax(Ax, Ay) / Dmin;
		rho = (fabs(M[8]) + 0.5*fabs(M[6]) + 0.5*fabs(M[7])) / Dmin;
		sane = sane && (Dmin > 0);
	} else{
		const double cx = W[0] * x0 + W[1] * y0 + W[2], cy = W[3] * x0 + W[4] * y0 + W[5];
		sane = (fabs(cx) < 1e8) && (fabs(cy) < 1e8);
		if(sane){ X0 = (int)floor(cx); Y0 = (int)floor(cy); }
		const double c[6] = { W[0] * s, W[1] * s, cx - X0, W[3] * s, W[4] * s, cy - Y0 };
#pragma unroll
		for(int i = 0; i < 6; ++i) cf[C_M + i] = (float)c[i];
		cf[C_M + 6] = 0; cf[C_M + 7] = 0; cf[C_M + 8] = 1;
		// Affine.cc:217-220 reads curr_state(2)+1, (3), (4), (5)+1 with curr_state = getStateFromWarp(curr_warp)
		cf[C_A + 0] = (float)((W[0] - 1) + 1); cf[C_A + 1] = (float)W[1]; cf[C_A + 2] = (float)W[3]; cf[C_A + 3] = (float)((W[4] - 1) + 1);
		cf[C_A + 4] = 0; cf[C_A + 5] = 0;
		E = fmax(fabs(c[0]) + fabs(c[1]) + fabs(c[2]), fabs(c[3]) + fabs(c[4]) + fabs(c[5]));
	}
	double delta = 2.5 * 5.9604644775390625e-8 * (6 + 4 * rho) * E + 2e-6;
	if(!sane || !(delta < 0.25)) delta = 2.0;                    // every pixel takes the fp64 path
	cf[C_DELTA] = (float)delta;
	// fast path only if all four neighbours are inside the image: 0 <= lx, lx + 1 <= w - 1 (same for y)
	cf[C_LOX] = (float)(-(double)X0); cf[C_HIX] = (float)((double)b.img.w - 2 - X0);
	cf[C_LOY] = (float)(-(double)Y0); cf[C_HIY] = (float)((double)b.img.h - 2 - Y0);
	ci[0] = X0; ci[1] = Y0;
}

// what the fp32 front end hands to the chain rule
struct PixF { float xl, yl, wxl, wyl, invD, val, gx, gy; int lx, ly; float ferr; };

// One pixel: template-local coordinates, warped point, sample + image gradient.
//   dl[9]: rows of the centred / scaled DLT ((dlt0 - x0 dlt2) / s, (dlt1 - y0 dlt2) / s, dlt2) -- unused with normalized_init
// WANT_TAP: also report the sampling indices of the fast path and its coordinate error (debug tap).
template<int SSM, bool WANT_TAP> __device__ __forceinline__ void pixel_front(const DevBatch &b, const PassConst &k, const float (&dl)[9],
	const double *s_dlt, const double *s_W, int row, int col, PixF &o){
	const float u = __ldg(b.xvf + col), v = __ldg(b.yvf + row);
	if(SSM == SSM_HOM && b.norm_init){ o.xl = u + u; o.yl = v + v; }             // s = 1/2
	else{
		const float hz = fmaf(dl[6], u, fmaf(dl[7], v, dl[8]));
		const float rz = rcp_approx(hz);
		o.xl = fmaf(dl[0], u, fmaf(dl[1], v, dl[2])) * rz;
		o.yl = fmaf(dl[3], u, fmaf(dl[4], v, dl[5])) * rz;
	}
	if(SSM == SSM_HOM){
		const float D = fmaf(k.m[6], u, fmaf(k.m[7], v, k.m[8]));
		o.invD = rcp_approx(D);
		o.wxl = fmaf(k.m[0], u, fmaf(k.m[1], v, k.m[2])) * o.invD;
		o.wyl = fmaf(k.m[3], u, fmaf(k.m[4], v, k.m[5])) * o.invD;
	} else{
		o.invD = 1.0f;
		o.wxl = fmaf(k.m[0], o.xl, fmaf(k.m[1], o.yl, k.m[2]));
		o.wyl = fmaf(k.m[3], o.xl, fmaf(k.m[4], o.yl, k.m[5]));
	}
	const float fx = floorf(o.wxl), fy = floorf(o.wyl);
	const float dx = o.wxl - fx, dy = o.wyl - fy;                          // exact (Sterbenz)
	const float hi = 1.0f - k.delta;
	// written so that NaN fails
	const bool fast = (dx >= k.delta) && (dx <= hi) && (dy >= k.delta) && (dy <= hi) &&
		(fx >= k.lox) && (fx <= k.hix) && (fy >= k.loy) && (fy <= k.hiy);
	if(fast){
		const int lx = k.X0 + (int)fx, ly = k.Y0 + (int)fy;
		const float *r0 = b.img.data + (size_t)ly*b.img.pitch + lx;
		const float p00 = __ldg(r0), p01 = __ldg(r0 + 1), p10 = __ldg(r0 + b.img.pitch), p11 = __ldg(r0 + b.img.pitch + 1);
		const float t0 = p01 - p00, t1 = p11 - p10;
		const float top = fmaf(dx, t0, p00), bot = fmaf(dx, t1, p10);
		o.gy = bot - top;                                                    // (1 - dx)(p10 - p00) + dx (p11 - p01)
		o.val = fmaf(dy, o.gy, top);
		o.gx = fmaf(dy, t1 - t0, t0);                                        // (1 - dy)(p01 - p00) + dy (p11 - p10)
		if(WANT_TAP){ o.lx = lx; o.ly = ly; }
	}
	if(!fast || WANT_TAP){
		// reference-exact fp64: warped point, indices, value, gradient (incl. the literal finite difference where the
		// two samples straddle a cell or leave the image)
		const PixGeom g = pixel_geometry<SSM>(MemMat3{ s_dlt }, MemMat3{ s_W }, __ldg(b.xv + col), __ldg(b.yv + row), b.norm_init != 0);
		if(WANT_TAP) o.ferr = fast ? (float)fmax(fabs((double)o.wxl - (g.wx - k.X0)), fabs((double)o.wyl - (g.wy - k.Y0))) : -1.0f;
		if(!fast){
			Sample smp = sample_fast<true>(b.img, g.wx, g.wy, b.grad_eps, 1.0);
			if(smp.lit) sample_literal(b.img, g.wx, g.wy, b.grad_eps, b.grad_mult, smp);
			o.val = (float)smp.val; o.gx = (float)smp.gx; o.gy = (float)smp.gy;
			o.wxl = (float)(g.wx - k.X0); o.wyl = (float)(g.wy - k.Y0);
			if(SSM == SSM_HOM) o.invD = (float)g.rD;
			if(WANT_TAP){
				const bool inb = !check_overflow(g.wx, g.wy, b.img.hd, b.img.wd);
				o.lx = inb ? (int)g.wx : -1; o.ly = inb ? (int)g.wy : -1;
			}
		}
	}
}

// chained gradient (Gx, Gy): the image gradient times d(warped point)/d(template point)
// (Homography.cc:250-262: Ix, Iy pre-multiplied by dw/dx / D; Affine.cc:221-230 with Gx = a Ix + c Iy, Gy = b Ix + d Iy)
template<int SSM> __device__ __forceinline__ void chain_gradient(const PassConst &k, const PixF &px, float &Gx, float &Gy){
	if(SSM == SSM_HOM){
		const float dwx_dx = fmaf(-k.a[4], px.wxl, k.a[0]), dwx_dy = fmaf(-k.a[5], px.wxl, k.a[1]);
		const float dwy_dx = fmaf(-k.a[4], px.wyl, k.a[2]), dwy_dy = fmaf(-k.a[5], px.wyl, k.a[3]);
		const float gxd = px.gx*px.invD, gyd = px.gy*px.invD;
		Gx = fmaf(dwx_dx, gxd, dwy_dx*gyd);
		Gy = fmaf(dwx_dy, gxd, dwy_dy*gyd);
	} else{
		Gx = fmaf(px.gx, k.a[0], px.gy*k.a[2]);
		Gy = fmaf(px.gx, k.a[1], px.gy*k.a[3]);
	}
}
// init_pix_jacobian<SSM> in the local frame
template<int SSM> __device__ __forceinline__ void local_row(float xl, float yl, float Gx, float Gy, float *J){
	if(SSM == SSM_HOM){
		J[0] = Gx*xl; J[1] = Gx*yl; J[2] = Gx; J[3] = Gy*xl; J[4] = Gy*yl; J[5] = Gy;
		const float nq = -J[0] - J[4];
		J[6] = xl*nq; J[7] = yl*nq;
	} else{
		J[0] = Gx; J[1] = Gy; J[2] = Gx*xl; J[3] = Gx*yl; J[4] = Gy*xl; J[5] = Gy*yl;
	}
}

} // namespace f32

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

template<int SSM, int SM, int T, int MINB>
__global__ void __launch_bounds__(T, MINB) ssd_update_f32_kernel(DevBatch b){
	constexpr int S = StateSize<SSM>::value;
	typedef AccLayout<S> L;
	const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	__shared__ double s_part[(T / 32) * L::NA];
	__shared__ double s_sum[L::NA];
	__shared__ double s_W[9], s_corners[8], s_init_corners[8], s_dlt[9];
	__shared__ double s_J[S], s_Hc[S*S], s_Hl[S*S], s_A[S*S], s_T[S*S], s_loc[3];
	__shared__ float s_cf[C_COUNT], s_dl[9];
	__shared__ int s_ci[2];
	__shared__ int s_ctrl;
	if(tid < 9){ s_W[tid] = b.warp[(size_t)p * 9 + tid]; s_dlt[tid] = b.dlt[(size_t)p * 9 + tid]; }
	if(tid < 8){ s_corners[tid] = b.corners[(size_t)p * 8 + tid]; s_init_corners[tid] = b.init_corners[(size_t)p * 8 + tid]; }
	cta_sync<T>();
	if(tid == 0){
		double x0, y0, s;
		template_frame<SSM>(b, s_init_corners, x0, y0, s);
		s_loc[0] = x0; s_loc[1] = y0; s_loc[2] = s;
		make_basis_map<SSM>(x0, y0, s, s_T);
		const double rs = 1.0 / s;
#pragma unroll
		for(int c = 0; c < 3; ++c){
			s_dl[c] = (float)((s_dlt[c] - x0*s_dlt[6 + c])*rs);
			s_dl[3 + c] = (float)((s_dlt[3 + c] - y0*s_dlt[6 + c])*rs);
			s_dl[6 + c] = (float)s_dlt[6 + c];
		}
		pass_constants<SSM>(b, s_W, s_dlt, x0, y0, s, s_cf, s_ci);
	}
	cta_sync<T>();
	float dl[9];
#pragma unroll
	for(int i = 0; i < 9; ++i) dl[i] = s_dl[i];
	const float *I0 = b.I0f + (size_t)p*b.N, *G0 = b.G0f + (size_t)p * 2 * b.N;
	const bool esm_mean = (SM == SM_ESM) && (b.jac_type == MTFB_ESM_JAC_ORIGINAL || b.hess_type == MTFB_ESM_HESS_ORIGINAL);
	const bool jac_half = (SM == SM_ESM) && (b.jac_type == MTFB_ESM_JAC_DIFF_OF_JACS);     // NT/ESM.cc:308-309
	const bool need_grad = (SM != SM_ICLK) || (b.hess_type == MTFB_LK_HESS_CURRENT_SELF);
	LMState lm = { 0.0, b.lm_delta_init, 0.0, false };
	int iter_id = 0, n_passes = 0, patch_status = 0;
	double f = 0;
	while(iter_id < b.max_iters){
		F32_PROF_T(0)
		PassConst k;
#pragma unroll
		for(int i = 0; i < 9; ++i) k.m[i] = s_cf[C_M + i];
#pragma unroll
		for(int i = 0; i < 6; ++i) k.a[i] = s_cf[C_A + i];
		k.delta = s_cf[C_DELTA]; k.lox = s_cf[C_LOX]; k.hix = s_cf[C_HIX]; k.loy = s_cf[C_LOY]; k.hiy = s_cf[C_HIY];
		k.X0 = s_ci[0]; k.Y0 = s_ci[1];
		float acc[L::NA];
#pragma unroll
		for(int i = 0; i < L::NA; ++i) acc[i] = 0;
		for(PixIter it(tid, T, b.resx); it.pix < b.N; it.next(T)){
			const float i0 = __ldcg(I0 + it.pix);
			PixF px;
			pixel_front<SSM, false>(b, k, dl, s_dlt, s_W, it.row, it.col, px);
			const float r = px.val - i0;                                   // I_diff (SSDBase.cc:78)
			float Jt[S], Jj[S], wj;
			if(need_grad){
				float Gx, Gy;
				chain_gradient<SSM>(k, px, Gx, Gy);
				local_row<SSM>(px.xl, px.yl, Gx, Gy, Jt);
			}
			if(SM == SM_ICLK){
				wj = r;                                                     // df_dI0 = I_diff (SSDBase.cc:34)
				local_row<SSM>(px.xl, px.yl, __ldcg(G0 + it.pix), __ldcg(G0 + b.N + it.pix), Jj);
			} else if(SM == SM_ESM){
				wj = -r;                                                    // df_dIt = -I_diff (SSDBase.cc:115-121)
				float J0[S];
				local_row<SSM>(px.xl, px.yl, __ldcg(G0 + it.pix), __ldcg(G0 + b.N + it.pix), J0);
				if(esm_mean){
#pragma unroll
					for(int s = 0; s < S; ++s) J0[s] = (J0[s] + Jt[s]) * 0.5f;     // NT/ESM.cc:246-248
				}
				if(b.jac_type == MTFB_ESM_JAC_ORIGINAL){
#pragma unroll
					for(int s = 0; s < S; ++s) Jj[s] = J0[s];
				} else{
#pragma unroll
					for(int s = 0; s < S; ++s) Jj[s] = esm_mean ? (2.0f*J0[s]) : (J0[s] + Jt[s]);   // SSDBase.cc:186
				}
				if(b.hess_type == MTFB_ESM_HESS_ORIGINAL){
#pragma unroll
					for(int s = 0; s < S; ++s) Jt[s] = J0[s];
				}
			} else{
				wj = -r;
#pragma unroll
				for(int s = 0; s < S; ++s) Jj[s] = Jt[s];
			}
			acc[0] = fmaf(r, r, acc[0]);
#pragma unroll
			for(int s = 0; s < S; ++s) acc[1 + s] = fmaf(wj, Jj[s], acc[1 + s]);
			if(need_grad){
#pragma unroll
				for(int i = 0; i < S; ++i){
#pragma unroll
					for(int j = i; j < S; ++j) acc[1 + S + L::tri(i, j)] = fmaf(Jt[i], Jt[j], acc[1 + S + L::tri(i, j)]);
				}
			}
		}
		F32_PROF_T(1)
		double accd[L::NA];
#pragma unroll
		for(int i = 0; i < L::NA; ++i) accd[i] = (double)acc[i];
		block_reduce<L::NA, T>(accd, s_part, s_sum);
		++n_passes;
		F32_PROF_T(2)
		// local basis -> the reference's: H = T^T H_loc T, g = T^T g_loc (fp64)
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
		F32_PROF_T(3)
		if(warp == 0){
			f = -s_sum[0] / 2;
			const int ctrl = serial_step<SSM, SM>(b, p, lane, iter_id, n_passes, f, s_J, s_Hc, s_W, s_corners, s_init_corners,
				lm, patch_status);
			if(lane == 0){
				s_ctrl = ctrl;
			}
			__syncwarp();
			F32_PROF_T(4)
			if(lane == 0 && ctrl != CTRL_BREAK) pass_constants<SSM>(b, s_W, s_dlt, s_loc[0], s_loc[1], s_loc[2], s_cf, s_ci);
			F32_PROF_T(5)
			F32_PROF_ADD()
		}
		cta_sync<T>();
		const int ctrl = s_ctrl;
		if(ctrl == CTRL_BREAK) break;
		if(counts_as_iteration<SM>(ctrl, b.nt_semantics)) ++iter_id;
	}
	if(warp == 0) store_patch_state<SSM>(b, p, lane, s_W, s_corners, f, n_passes, patch_status);
}

// debug tap of the fp32 front end at the current state: sampling indices (lx, ly; -1 = outside the image), pixel
// value, image gradient, reference-basis Jacobian row (through the basis map) and the fast path's coordinate error
template<int SSM, int T>
__global__ void __launch_bounds__(T) ssd_stage_f32_kernel(DevBatch b, StageTapsF32 t){
	constexpr int S = StateSize<SSM>::value;
	const int p = blockIdx.x, tid = threadIdx.x;
	__shared__ double s_W[9], s_dlt[9], s_ic[8], s_T[S*S], s_loc[3];
	__shared__ float s_cf[C_COUNT], s_dl[9];
	__shared__ int s_ci[2];
	if(tid < 9){ s_W[tid] = b.warp[(size_t)p * 9 + tid]; s_dlt[tid] = b.dlt[(size_t)p * 9 + tid]; }
	if(tid < 8) s_ic[tid] = b.init_corners[(size_t)p * 8 + tid];
	__syncthreads();
	if(tid == 0){
		double x0, y0, s;
		template_frame<SSM>(b, s_ic, x0, y0, s);
		s_loc[0] = x0; s_loc[1] = y0; s_loc[2] = s;
		make_basis_map<SSM>(x0, y0, s, s_T);
		const double rs = 1.0 / s;
		for(int c = 0; c < 3; ++c){
			s_dl[c] = (float)((s_dlt[c] - x0*s_dlt[6 + c])*rs);
			s_dl[3 + c] = (float)((s_dlt[3 + c] - y0*s_dlt[6 + c])*rs);
			s_dl[6 + c] = (float)s_dlt[6 + c];
		}
		pass_constants<SSM>(b, s_W, s_dlt, x0, y0, s, s_cf, s_ci);
	}
	__syncthreads();
	PassConst k;
	float dl[9];
	for(int i = 0; i < 9; ++i){ k.m[i] = s_cf[C_M + i]; dl[i] = s_dl[i]; }
	for(int i = 0; i < 6; ++i) k.a[i] = s_cf[C_A + i];
	k.delta = s_cf[C_DELTA]; k.lox = s_cf[C_LOX]; k.hix = s_cf[C_HIX]; k.loy = s_cf[C_LOY]; k.hiy = s_cf[C_HIY];
	k.X0 = s_ci[0]; k.Y0 = s_ci[1];
	const size_t N = b.N;
	for(PixIter it(tid, T, b.resx); it.pix < b.N; it.next(T)){
		PixF px;
		pixel_front<SSM, true>(b, k, dl, s_dlt, s_W, it.row, it.col, px);
		float Gx, Gy, Jl[S];
		chain_gradient<SSM>(k, px, Gx, Gy);
		local_row<SSM>(px.xl, px.yl, Gx, Gy, Jl);
		const size_t q = p*N + it.pix;
		if(t.idx){ t.idx[2 * q] = px.lx; t.idx[2 * q + 1] = px.ly; }
		if(t.pix_vals) t.pix_vals[q] = px.val;
		if(t.pix_grad){ t.pix_grad[p * 2 * N + it.pix] = px.gx; t.pix_grad[p * 2 * N + N + it.pix] = px.gy; }
		if(t.fast_err) t.fast_err[q] = px.ferr;
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
template<int SSM, int SM> static cudaError_t launch_update_f32_t(int threads, const DevBatch &b, cudaStream_t st){
	switch(threads){
	case 32: ssd_update_f32_kernel<SSM, SM, 32, 14><<<b.P, 32, 0, st>>>(b); break;
	case 64: ssd_update_f32_kernel<SSM, SM, 64, 7><<<b.P, 64, 0, st>>>(b); break;
	case 128: ssd_update_f32_kernel<SSM, SM, 128, 4><<<b.P, 128, 0, st>>>(b); break;
	case 256: ssd_update_f32_kernel<SSM, SM, 256, 2><<<b.P, 256, 0, st>>>(b); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
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
