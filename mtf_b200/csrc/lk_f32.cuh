// lk_f32.cuh -- device functions of the "fp32 arithmetic, exact sampling indices" precision (MTFB_PRECISION_F32) shared by
// its kernels: lk_ssd_f32.cu (one accumulator per J^T J entry; ESM / FCLK / ICLK) and lk_ssd_mom.cu (column-fixed threads,
// moment accumulators; FCLK).  See lk_ssd_f32.cu for the description of the precision.
#pragma once
#include <type_traits>
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
	// where the four neighbours are read from: the frame in global memory, or the patch's window of it in shared memory
	// (generic pointer), with the integer origin relative to that array
	const float *base; int pitch, Xr, Yr;
};
// side of the square frame window a CTA keeps in shared memory (pitch F32_WIN + 1: rotated patches walk columns)
constexpr int F32_WIN = 56, F32_WINP = F32_WIN + 1;

__device__ __forceinline__ float rcp_approx(float x){
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));       // MUFU.RCP, <= 1 ulp
	return r;
}

// 1 / d in fp64 from the fp32 reciprocal and two Newton steps (d > 0, normal range)
__device__ __forceinline__ double rcp_newton(double d){
	double r = (double)rcp_approx((float)d);
	r = fma(fma(-d, r, 1.0), r, r);
	r = fma(fma(-d, r, 1.0), r, r);
	return r;
}

// one Newton step: relative error ~4e-15, plenty for the pivots of a solve whose inputs carry fp32 rounding
__device__ __forceinline__ double rcp_newton1(double d){
	const double r = (double)rcp_approx((float)d);
	return fma(fma(-d, r, 1.0), r, r);
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

// Per-pass constants, by the 32 lanes of warp 0 (lane i < 9 owns entry i of the 3 x 3 product).
//
// Error bound behind delta (Homography; u, v in [-1/2, 1/2]).  wxl = num / den with num = c0 u + c1 v + c2 evaluated as
// two fp32 fmaf's on coefficients rounded to fp32 (2^-24 relative) and grid values u = fmaf(i, step, low) that differ
// from the fp64 LinSpaced values by <= 1.5 . 2^-24 absolute (= 3 . 2^-24 relative to |u|max = 1/2): every term carries
// <= 4 roundings of 2^-24 and each fmaf one more on its partial sum, so |err(num)| <= 6 . 2^-24 . A with
// A = |c0|/2 + |c1|/2 + |c2|, and likewise |err(den)| <= 6 . 2^-24 . B; MUFU.RCP (1 ulp) and the final product add
// 2 . 2^-24 relative, the (wxl - 1/2) of the floor another 2^-24.  With Dmin <= |den|, E = A / Dmin >= |wxl| and
// rho = B / Dmin >= 1:      |err(wxl)| <= 2^-24 . E . (9 + 6 rho)          (= 9e-7 E at rho = 1).
// delta = 2.5 x that bound + 2e-6 px (tests/test_gpu_f32.py measures the actual error through the tap).
// Affine: wxl = c0 xl + c1 yl + c2 with |xl|, |yl| <= 1 known to 14 . 2^-24 (their own sums and quotient):
// E = |c0| + |c1| + |c2|, |err| <= (14 + 4) . 2^-24 . E: covered by the same formula with rho = 2.
template<int SSM> __device__ __forceinline__ void pass_constants(const DevBatch &b, int lane, const double *W, const double *dlt,
	double x0, double y0, double s, float *cf, int *ci){
	// The integer origin only has to be NEAR the patch centre (any integer is a valid origin), and delta only has to be an
	// upper bound with its 2.5x margin: both are computed in fp32.  What must be exact -- the centred rows
	// M0 - X0 M2, M1 - Y0 M2 -- is one fp64 fma per entry before the rounding to fp32.
	float E, rho = 2.0f;
	int X0 = 0, Y0 = 0;
	bool sane = true;
	if(SSM == SSM_HOM){
		double Mi;
		{
			const int i = lane < 9 ? lane : 0, r = i / 3, c = i - 3 * r;
			Mi = b.norm_init ? W[i] : fma(W[3 * r + 2], dlt[6 + c], fma(W[3 * r + 1], dlt[3 + c], W[3 * r] * dlt[c]));
		}
		const float Mf = (float)Mi;
		const float m2 = __shfl_sync(FULL_MASK, Mf, 2), m5 = __shfl_sync(FULL_MASK, Mf, 5), m8 = __shfl_sync(FULL_MASK, Mf, 8);
		const float r8 = rcp_approx(m8);
		const float cx = m2 * r8, cy = m5 * r8;
		sane = (m8 > 1e-30f) && (fabsf(cx) < 2e6f) && (fabsf(cy) < 2e6f);          // the magic-number floor needs |coordinates| < 2^22
		if(sane){ X0 = (int)floorf(cx); Y0 = (int)floorf(cy); }
		const double mz = __shfl_sync(FULL_MASK, Mi, 6 + lane % 3);
		if(lane < 6) Mi = fma(-(double)(lane < 3 ? X0 : Y0), mz, Mi);
		const float Mc = fabsf((float)Mi);
		if(lane < 9) cf[C_M + lane] = (float)Mi;
		if(lane >= 9 && lane < 15){
			// a00 - a20 X0, a01 - a21 X0, a10 - a20 Y0, a11 - a21 Y0, a20, a21
			const int q = lane - 9;
			const int ia = q < 2 ? q : (q < 4 ? q + 1 : q + 2), ib = 6 + (q & 1);
			const double sc = q < 2 ? (double)X0 : (q < 4 ? (double)Y0 : 0.0);
			cf[C_A + q] = (float)fma(-W[ib], sc, W[ia]);
		}
		float m[9];
#pragma unroll
		for(int i = 0; i < 9; ++i) m[i] = __shfl_sync(FULL_MASK, Mc, i);
		const float Ax = 0.5f*m[0] + 0.5f*m[1] + m[2], Ay = 0.5f*m[3] + 0.5f*m[4] + m[5];
		const float Dmin = m[8] - 0.5f*m[6] - 0.5f*m[7];
		const bool dmin_ok = (Dmin > 1e-30f) && (Dmin < 1e30f);
		const float rD = rcp_approx(dmin_ok ? Dmin : 1.0f);
		E = fmaxf(Ax, Ay) * rD;
		rho = (m[8] + 0.5f*m[6] + 0.5f*m[7]) * rD;
		sane = sane && dmin_ok;
	} else{
		const double cx = W[0] * x0 + W[1] * y0 + W[2], cy = W[3] * x0 + W[4] * y0 + W[5];
		sane = (fabs(cx) < 2e6) && (fabs(cy) < 2e6);
		if(sane){ X0 = (int)floor(cx); Y0 = (int)floor(cy); }
		const double c[6] = { W[0] * s, W[1] * s, cx - X0, W[3] * s, W[4] * s, cy - Y0 };
		if(lane == 0){
#pragma unroll
			for(int i = 0; i < 6; ++i) cf[C_M + i] = (float)c[i];
			cf[C_M + 6] = 0; cf[C_M + 7] = 0; cf[C_M + 8] = 1;
			// Affine.cc:217-220 reads curr_state(2)+1, (3), (4), (5)+1 with curr_state = getStateFromWarp(curr_warp)
			cf[C_A + 0] = (float)((W[0] - 1) + 1); cf[C_A + 1] = (float)W[1]; cf[C_A + 2] = (float)W[3]; cf[C_A + 3] = (float)((W[4] - 1) + 1);
			cf[C_A + 4] = 0; cf[C_A + 5] = 0;
		}
		E = (float)fmax(fabs(c[0]) + fabs(c[1]) + fabs(c[2]), fabs(c[3]) + fabs(c[4]) + fabs(c[5]));
	}
	// (1 + 1e-5): the bound itself is evaluated in fp32
	float delta = 2.5f * 5.9604644775390625e-8f * (9.0f + 6.0f * rho) * E * 1.00001f + 2e-6f;
	if(!sane || !(delta < 0.25f)) delta = 2.0f;                  // every pixel takes the fp64 path
	if(lane == 0){
		cf[C_DELTA] = delta;
		// fast path only if all four neighbours are inside the image: 0 <= lx, lx + 1 <= w - 1 (same for y)
		cf[C_LOX] = (float)(-X0); cf[C_HIX] = (float)(b.img.w - 2 - X0);
		cf[C_LOY] = (float)(-Y0); cf[C_HIY] = (float)(b.img.h - 2 - Y0);
		ci[0] = X0; ci[1] = Y0;
	}
}

// Frame window in shared memory (one thread, once per pass): does the F32_WIN x F32_WIN window staged at origin
// (wi[0], wi[1]) still hold every neighbour the patch's samples can touch -- the integer hull of its four corners plus
// one?  If not, pick a new origin centred on the patch (clipped to the frame) and ask for a restage; a patch that does
// not fit the window (or a frame smaller than it) samples the frame in global memory for this pass.
// wi: 0 ox, 1 oy, 2 use the window this pass, 3 restage before the pass, 4 window contents valid, 5 the window holds the
// whole hull (+ 1 pixel) of the patch: no per-sample range test needed
__device__ __forceinline__ void window_decide(const DevBatch &b, const double *corners, bool have_window, int *wi){
	wi[2] = 0; wi[3] = 0; wi[5] = 0;
	if(!have_window || b.img.w < F32_WIN || b.img.h < F32_WIN) return;
	// fp32 is enough: the hull is widened by a pixel, and a sample the window misses takes the fp64 path anyway
	float x0 = (float)corners[0], x1 = x0, y0 = (float)corners[4], y1 = y0;
#pragma unroll
	for(int i = 1; i < 4; ++i){
		const float cx = (float)corners[i], cy = (float)corners[4 + i];
		x0 = fminf(x0, cx); x1 = fmaxf(x1, cx); y0 = fminf(y0, cy); y1 = fmaxf(y1, cy);
	}
	if(!(x0 > -1e6f && x1 < 1e6f && y0 > -1e6f && y1 < 1e6f)) return;
	const int ix0 = (int)floorf(x0) - 1, ix1 = (int)floorf(x1) + 2, iy0 = (int)floorf(y0) - 1, iy1 = (int)floorf(y1) + 2;
	if(ix1 - ix0 + 1 > F32_WIN + 2 || iy1 - iy0 + 1 > F32_WIN + 2) return;
	const bool covered = wi[4] && ix0 + 1 >= wi[0] && ix1 - 1 <= wi[0] + F32_WIN - 1 && iy0 + 1 >= wi[1] && iy1 - 1 <= wi[1] + F32_WIN - 1;
	if(!covered){
		int ox = (ix0 + ix1 + 1) / 2 - F32_WIN / 2, oy = (iy0 + iy1 + 1) / 2 - F32_WIN / 2;
		ox = ox < 0 ? 0 : (ox > b.img.w - F32_WIN ? b.img.w - F32_WIN : ox);
		oy = oy < 0 ? 0 : (oy > b.img.h - F32_WIN ? b.img.h - F32_WIN : oy);
		wi[0] = ox; wi[1] = oy; wi[3] = 1; wi[4] = 1;
	}
	wi[2] = 1;
	// (false only when the clipping at the frame border moved the window off the patch)
	wi[5] = (ix0 >= wi[0] && ix1 <= wi[0] + F32_WIN - 1 && iy0 >= wi[1] && iy1 <= wi[1] + F32_WIN - 1) ? 1 : 0;
}

// per-patch setup by warp 0: template frame, basis maps, centred DLT rows, first pass constants
template<int SSM> __device__ __forceinline__ void patch_setup(const DevBatch &b, int lane, const double *s_W, const double *s_dlt,
	const double *s_init_corners, double *s_loc, double *s_T, double *s_Tinv, float *s_dl, float *s_cf, int *s_ci){
	double x0, y0, s;
	template_frame<SSM>(b, s_init_corners, x0, y0, s);
	const double rs = 1.0 / s;
	if(lane == 0){
		s_loc[0] = x0; s_loc[1] = y0; s_loc[2] = s;
		make_basis_map<SSM>(x0, y0, s, s_T);
	}
	if(lane == 1) make_basis_map<SSM>(-x0*rs, -y0*rs, rs, s_Tinv);        // xl = -x0/s + x/s: the inverse map
	if(lane == 2){
		// centred / scaled DLT rows.  Third row constant (rectangles, parallelograms: an affine DLT up to rounding):
		// the division is folded into the coefficients and the pixel loop skips it (s_dl[6] = s_dl[7] = 0 exactly)
		// hom_normalized_init: the template points are the grid itself, (xl, yl) = (u, v) / s with s = 1/2
		const bool unit = (SSM == SSM_HOM) && b.norm_init;
		// (xl, yl) only parametrise the Jacobian basis, in fp32: a relative 1e-9 is invisible there
		const bool aff = unit || (fabs(s_dlt[6]) <= 1e-9*fabs(s_dlt[8]) && fabs(s_dlt[7]) <= 1e-9*fabs(s_dlt[8]));
		const double q = aff ? rs / s_dlt[8] : rs;
#pragma unroll
		for(int c = 0; c < 3; ++c){
			s_dl[c] = unit ? (c == 0 ? 2.0f : 0.0f) : (float)((s_dlt[c] - x0*s_dlt[6 + c])*q);
			s_dl[3 + c] = unit ? (c == 1 ? 2.0f : 0.0f) : (float)((s_dlt[3 + c] - y0*s_dlt[6 + c])*q);
			s_dl[6 + c] = aff ? (c == 2 ? 1.0f : 0.0f) : (float)s_dlt[6 + c];
		}
	}
	pass_constants<SSM>(b, lane, s_W, s_dlt, x0, y0, s, s_cf, s_ci);
}

// pixel iterator with the grid position kept in fp32 (small integers: exact) -- no int -> float conversions
struct PixIterF {
	int pix; float rowf, colf, dcolf, drowf, resxf;
	__device__ __forceinline__ PixIterF(){}
	__device__ __forceinline__ PixIterF(int tid, int step, int resx) : pix(tid), rowf((float)(tid / resx)), colf((float)(tid % resx)),
		dcolf((float)(step % resx)), drowf((float)(step / resx)), resxf((float)resx){}
	__device__ __forceinline__ void next(int step){
		pix += step; colf += dcolf; rowf += drowf;
		if(colf >= resxf){ colf -= resxf; rowf += 1.0f; }
	}
};

// what the fp32 front end hands to the chain rule
struct PixF { float xl, yl, wxl, wyl, invD, val, gx, gy; int lx, ly; bool fast; };

// floor(x) for |x| < 2^22 without the conversion pipe: RN(x - 1/2) through the 1.5 * 2^23 trick.  Ties and
// near-ties (x within an ulp of an integer) may come out one off; the caller's guard band sends those to fp64.
__device__ __forceinline__ void fast_floor(float x, float &fl, int &il){
	const float magic = 12582912.0f;
	const float t = (x - 0.5f) + magic;
	fl = t - magic;
	il = __float_as_int(t) - 0x4B400000;
}

// One pixel, straight-line fp32: template-local coordinates, warped point, and -- where the point is safely inside a
// pixel cell (o.fast) -- the bilinear sample and its gradient.  Loads are issued for every lane (at a safe address
// when !fast) so that two pixels can be interleaved without a branch in between.
//   dl[9]: rows of the centred / scaled DLT (patch_setup) -- unused with normalized_init
// CHK = false: the caller knows that every sample's four neighbours lie inside the array it reads (the frame window covers the
// hull of the patch's corners plus a pixel: window_decide), so the four range comparisons are dropped.
template<int SSM, bool CHK = true> __device__ __forceinline__ void front_fast(const DevBatch &b, const PassConst &k, const float (&dl)[9], bool dlt_affine,
	float rowf, float colf, PixF &o){
	const float u = fmaf(colf, b.gx_step, b.gx_lo), v = fmaf(rowf, b.gy_step, b.gy_lo);
	o.xl = fmaf(dl[0], u, fmaf(dl[1], v, dl[2]));
	o.yl = fmaf(dl[3], u, fmaf(dl[4], v, dl[5]));
	if(!dlt_affine){
		const float rz = rcp_approx(fmaf(dl[6], u, fmaf(dl[7], v, dl[8])));
		o.xl *= rz; o.yl *= rz;
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
	float fx, fy; int ix, iy;
	fast_floor(o.wxl, fx, ix);
	fast_floor(o.wyl, fy, iy);
	const float dx = o.wxl - fx, dy = o.wyl - fy;
	const float hi = 1.0f - k.delta;
	// written so that NaN fails
	o.fast = (dx >= k.delta) && (dx <= hi) && (dy >= k.delta) && (dy <= hi);
	if(CHK) o.fast = o.fast && (fx >= k.lox) && (fx <= k.hix) && (fy >= k.loy) && (fy <= k.hiy);
	o.lx = k.X0 + ix; o.ly = k.Y0 + iy;
	const int off = o.fast ? (k.Yr + iy)*k.pitch + (k.Xr + ix) : 0;
	const float *r0 = k.base + off, *r1 = r0 + k.pitch;
	const float p00 = r0[0], p01 = r0[1], p10 = r1[0], p11 = r1[1];
	const float t0 = p01 - p00, t1 = p11 - p10;
	const float top = fmaf(dx, t0, p00), bot = fmaf(dx, t1, p10);
	o.gy = bot - top;                                                    // (1 - dx)(p10 - p00) + dx (p11 - p01)
	o.val = fmaf(dy, o.gy, top);
	o.gx = fmaf(dy, t1 - t0, t0);                                        // (1 - dy)(p01 - p00) + dy (p11 - p10)
}
// The same pixel through the reference-exact fp64 functions of lk_math.cuh: warped point, indices, value, gradient
// (incl. the literal finite difference where the two samples straddle a cell or leave the image).
// Inlined into the rarely taken branch: an out-of-line call was measured 12 % slower (the live accumulators are saved and
// restored around it: profiles/README.md).
struct ExactOut { float val, gx, gy, wxl, wyl, invD; int lx, ly; };
struct ExactArgs { Image img; const double *xv, *yv; const double *s_dlt, *s_W; double grad_eps, grad_mult; int norm_init, X0, Y0; };
template<int SSM> __device__ __forceinline__ ExactOut exact_pixel(ExactArgs a, int row, int col){
	const PixGeom g = pixel_geometry<SSM>(MemMat3{ a.s_dlt }, MemMat3{ a.s_W }, __ldg(a.xv + col), __ldg(a.yv + row), a.norm_init != 0);
	Sample smp = sample_fast<true>(a.img, g.wx, g.wy, a.grad_eps, 1.0);
	if(smp.lit) sample_literal(a.img, g.wx, g.wy, a.grad_eps, a.grad_mult, smp);
	ExactOut o;
	o.val = (float)smp.val; o.gx = (float)smp.gx; o.gy = (float)smp.gy;
	o.wxl = (float)(g.wx - a.X0); o.wyl = (float)(g.wy - a.Y0);
	o.invD = (SSM == SSM_HOM) ? (float)g.rD : 1.0f;
	const bool inb = !check_overflow(g.wx, g.wy, a.img.hd, a.img.wd);
	o.lx = inb ? (int)g.wx : -1; o.ly = inb ? (int)g.wy : -1;
	return o;
}
template<int SSM> __device__ __forceinline__ void front_exact(const DevBatch &b, const PassConst &k, const double *s_dlt, const double *s_W,
	float rowf, float colf, PixF &o){
	ExactArgs a;
	a.img = b.img; a.xv = b.xv; a.yv = b.yv; a.s_dlt = s_dlt; a.s_W = s_W; a.grad_eps = b.grad_eps; a.grad_mult = b.grad_mult;
	a.norm_init = b.norm_init; a.X0 = k.X0; a.Y0 = k.Y0;
	const ExactOut e = exact_pixel<SSM>(a, (int)rowf, (int)colf);
	o.val = e.val; o.gx = e.gx; o.gy = e.gy; o.wxl = e.wxl; o.wyl = e.wyl; o.invD = e.invD; o.lx = e.lx; o.ly = e.ly;
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

// Per-thread sums in packed fp32x2 registers, updated with fma.rn.f32x2 (SASS FFMA2: two fused multiply-adds per
// issue slot; ptxas folds the {a, a} operand into the instruction's scalar-broadcast form).  Pairing: with
// P_i = (J_i, J_i+1) for even i,   P_i * J_i -> (H_ii, H_i,i+1),   P_i * J_j -> (H_ij, H_i+1,j) for j >= i + 2,
// H_i+1,i+1 alone; (g_i, g_i+1) += w * P_i.  S = 8: 20 FFMA2 + 5 FFMA instead of 45 FFMA per pixel.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi){
	unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi){
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void ffma2(unsigned long long &d, unsigned long long a, unsigned long long b){
	asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
#ifdef MTFB_F32_NO_PACK     // experiment: one FFMA per sum (45 per pixel) instead of the packed form
template<int S> struct PackedAcc {
	typedef AccLayout<S> L;
	float a[L::NA];
	__device__ __forceinline__ void clear(){
#pragma unroll
		for(int i = 0; i < L::NA; ++i) a[i] = 0;
	}
	__device__ __forceinline__ void add(float r, float wj, const float (&Jj)[S], const float (&Jt)[S], bool with_hessian){
		a[0] = fmaf(r, r, a[0]);
#pragma unroll
		for(int s = 0; s < S; ++s) a[1 + s] = fmaf(wj, Jj[s], a[1 + s]);
		if(with_hessian){
#pragma unroll
			for(int i = 0; i < S; ++i)
#pragma unroll
			for(int j = i; j < S; ++j) a[1 + S + L::tri(i, j)] = fmaf(Jt[i], Jt[j], a[1 + S + L::tri(i, j)]);
		}
	}
	__device__ __forceinline__ void unpack(float (&acc)[L::NA]) const{
#pragma unroll
		for(int i = 0; i < L::NA; ++i) acc[i] = a[i];
	}
};
#else
template<int S> struct PackedAcc {
	typedef AccLayout<S> L;
	static constexpr int NP = S*S / 4;                 // pairs of the Hessian: S/2 row pairs x (1 + (S - 2 - i) ...) summed = S^2/4
	unsigned long long h2[NP], g2[S / 2];
	float hd[S / 2], f;
	__device__ __forceinline__ void clear(){
#pragma unroll
		for(int i = 0; i < NP; ++i) h2[i] = 0ull;
#pragma unroll
		for(int i = 0; i < S / 2; ++i){ g2[i] = 0ull; hd[i] = 0; }
		f = 0;
	}
	__device__ __forceinline__ void add(float r, float wj, const float (&Jj)[S], const float (&Jt)[S], bool with_hessian){
		f = fmaf(r, r, f);
		const unsigned long long w2 = pack2(wj, wj);
#pragma unroll
		for(int i = 0; i < S; i += 2) ffma2(g2[i / 2], w2, pack2(Jj[i], Jj[i + 1]));
		if(with_hessian){
			int q = 0;
#pragma unroll
			for(int i = 0; i < S; i += 2){
				const unsigned long long P = pack2(Jt[i], Jt[i + 1]);
				ffma2(h2[q++], pack2(Jt[i], Jt[i]), P);
				hd[i / 2] = fmaf(Jt[i + 1], Jt[i + 1], hd[i / 2]);
#pragma unroll
				for(int j = i + 2; j < S; ++j) ffma2(h2[q++], pack2(Jt[j], Jt[j]), P);
			}
		}
	}
	// -> AccLayout order: sum r^2 | J^T d | upper triangle of J^T J
	__device__ __forceinline__ void unpack(float (&acc)[L::NA]) const{
		acc[0] = f;
#pragma unroll
		for(int i = 0; i < S; i += 2) unpack2(g2[i / 2], acc[1 + i], acc[2 + i]);
		int q = 0;
#pragma unroll
		for(int i = 0; i < S; i += 2){
			unpack2(h2[q++], acc[1 + S + L::tri(i, i)], acc[1 + S + L::tri(i, i + 1)]);
			acc[1 + S + L::tri(i + 1, i + 1)] = hd[i / 2];
#pragma unroll
			for(int j = i + 2; j < S; ++j) unpack2(h2[q++], acc[1 + S + L::tri(i, j)], acc[1 + S + L::tri(i + 1, j)]);
		}
	}
};

#endif

// what one pixel adds to the sums (same cases as lk_ssd_terms.cuh pixel_terms / accumulate_terms)
template<int SSM, int SM> __device__ __forceinline__ void accumulate_pixel(const DevBatch &b, const PassConst &k, const PixF &px, float i0,
	const float *__restrict__ G0, int pix, bool valid, bool need_grad, bool esm_mean, PackedAcc<StateSize<SSM>::value> &acc){
	constexpr int S = StateSize<SSM>::value;
	const float r = valid ? px.val - i0 : 0.0f;                     // I_diff (SSDBase.cc:78)
	float Jt[S], Jj[S], wj;
#pragma unroll
	for(int s = 0; s < S; ++s) Jt[s] = 0;
	if(need_grad){
		float Gx, Gy;
		chain_gradient<SSM>(k, px, Gx, Gy);
		if(!valid){ Gx = 0; Gy = 0; }
		local_row<SSM>(px.xl, px.yl, Gx, Gy, Jt);
	}
	if(SM == SM_ICLK){
		wj = r;                                                     // df_dI0 = I_diff (SSDBase.cc:34)
		local_row<SSM>(px.xl, px.yl, __ldcg(G0 + pix), __ldcg(G0 + b.N + pix), Jj);
	} else if(SM == SM_ESM){
		wj = -r;                                                    // df_dIt = -I_diff (SSDBase.cc:115-121)
		float J0[S];
		float g0x = __ldcg(G0 + pix), g0y = __ldcg(G0 + b.N + pix);
		if(!valid){ g0x = 0; g0y = 0; }
		local_row<SSM>(px.xl, px.yl, g0x, g0y, J0);
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
	acc.add(r, wj, Jj, Jt, need_grad);
}

// Solve (J_loc^T J_loc) x = g for the state update in the LOCAL basis, one column per lane, rhs on lane S; then map to
// the reference's parameters: dp = T^-1 x.  Returns false (uniformly) if a pivot is not safely positive -- the caller then
// takes the reference's rank-revealing QR in the reference basis.
// In exact arithmetic dp equals the reference's -H^-1 J^T (NT/FCLK.cc:298): H = -T^T A T, J^T = T^T g.
//
// The matrix is symmetric positive definite and, in this basis, well conditioned: Gauss-Jordan without pivoting, each
// pivot row normalised by a reciprocal (fp32 seed + one Newton step).
// MTFB_F32_BAREISS (experiment builds) selects the fraction-free form a_ij <- (a_kk a_ij - a_ik a_kj) / a_(k-1)(k-1), whose
// only division is by the PREVIOUS pivot, i.e. off the dependent chain -- measured slower (0.657 vs 0.640 ms per frame):
// three fp64 operations per entry instead of one load the shared fp64 pipe more than the shorter chain saves.
template<int S> __device__ __forceinline__ bool solve_local(int lane, const double *s_sum, const double *s_Tinv, double g_scale, double *s_x,
	double *s_dp){
	typedef AccLayout<S> L;
	double a[S];
	const int j = lane < S ? lane : 0;
#pragma unroll
	for(int i = 0; i < S; ++i) a[i] = (lane == S) ? s_sum[1 + i] * g_scale : s_sum[1 + S + L::tri(i < j ? i : j, i < j ? j : i)];
#ifndef MTFB_F32_BAREISS
	double dmax = 0;
#pragma unroll
	for(int i = 0; i < S; ++i) dmax = fmax(dmax, s_sum[1 + S + L::tri(i, i)]);
	bool ok = dmax > 0 && dmax < 1e300;
#pragma unroll
	for(int kk = 0; kk < S; ++kk){
		const double piv = __shfl_sync(FULL_MASK, a[kk], kk);
		ok = ok && (piv > 1e-11 * dmax);
		double v[S];
#pragma unroll
		for(int i = 0; i < S; ++i) if(i != kk) v[i] = __shfl_sync(FULL_MASK, a[i], kk);
		const double t = a[kk] * rcp_newton1(ok ? piv : 1.0);
		a[kk] = t;
#pragma unroll
		for(int i = 0; i < S; ++i) if(i != kk) a[i] = fma(-v[i], t, a[i]);
	}
	if(!ok) return false;
#else
	// Jacobi scaling: s_i ~ 1 / sqrt(A_ii) (any positive scaling is exact algebra; fp32 rsqrt is plenty)
	float sc[S];
	{
		float mine = 0.0f;
#pragma unroll
		for(int i = 0; i < S; ++i) if(lane == i) mine = (float)a[i];
		mine = rsqrtf(mine);
#pragma unroll
		for(int i = 0; i < S; ++i) sc[i] = __shfl_sync(FULL_MASK, mine, i);
	}
	bool ok = true;
#pragma unroll
	for(int i = 0; i < S; ++i) ok = ok && (sc[i] > 0.0f) && (sc[i] < 1e18f);          // diagonal finite and positive
	{
		double my_s = 1.0;
#pragma unroll
		for(int i = 0; i < S; ++i) if(lane == i) my_s = (double)sc[i];
#pragma unroll
		for(int i = 0; i < S; ++i) a[i] = a[i] * ((double)sc[i] * my_s);              // rhs lane: my_s = 1
	}
	double rprev = 1.0;
#pragma unroll
	for(int kk = 0; kk < S; ++kk){
		const double piv = __shfl_sync(FULL_MASK, a[kk], kk);
		ok = ok && (piv * rprev > 1e-11);                                             // the ordinary pivot of the unit-diagonal matrix
		double v[S];
#pragma unroll
		for(int i = 0; i < S; ++i) if(i != kk) v[i] = __shfl_sync(FULL_MASK, a[i], kk);
		const double akk = a[kk];
#pragma unroll
		for(int i = 0; i < S; ++i) if(i != kk) a[i] = fma(piv, a[i], -(v[i] * akk)) * rprev;
		rprev = rcp_newton1(ok ? piv : 1.0);                                          // needed one step later
	}
	if(!ok) return false;
	// every diagonal entry is now det (= the last pivot): x~ = rhs / det, x = diag(s) x~
	{
		const double det = __shfl_sync(FULL_MASK, a[S - 1], S - 1);
		const double rdet = rcp_newton(det);
#pragma unroll
		for(int i = 0; i < S; ++i) a[i] = a[i] * rdet * (double)sc[i];
	}
#endif
	if(lane == S){
#pragma unroll
		for(int i = 0; i < S; ++i) s_x[i] = a[i];
	}
	__syncwarp();
	if(lane < S){
		double d0 = 0, d1 = 0;
#pragma unroll
		for(int m = 0; m < S; m += 2){ d0 = fma(s_Tinv[lane*S + m], s_x[m], d0); d1 = fma(s_Tinv[lane*S + m + 1], s_x[m + 1], d1); }
		s_dp[lane] = d0 + d1;
	}
	__syncwarp();
	return true;
}

// The reference's solve for the pass-local Hessian, by one warp and without a trip through shared memory: lane j < S builds
// column j of H = -T^T (J_loc^T J_loc) T (the SSD self Hessian in the reference's parameters, SSDBase.h:91-94, fp64), lane S
// the right-hand side J^T = g_scale T^T g_loc, and the warp runs the column-pivoted Householder QR with Eigen's rank rule
// (`-H.colPivHouseholderQr().solve(J^T)`, SM/src/NT/FCLK.cc:298, NT/ESM.cc:266) in its F32-precision variant (lk_warp.cuh
// factor_lean).  s_T row-major: J_ref[k] = sum_m J_loc[m] T[m][k].  Writes s_dp[S]; sets MTFB_PATCH_SINGULAR like
// serial_step does.
template<int S> __device__ __forceinline__ void solve_reference_warp(int lane, const double *s_sum, const double *s_T, double g_scale,
	double *s_dp, int &patch_status, long long *prof = nullptr){
	typedef AccLayout<S> L;
#if MTFB_PROF
	const long long pt0 = clock64();
#endif
	const int j = lane < S ? lane : 0;
	double tj[S], A[S];
#pragma unroll
	for(int n = 0; n < S; ++n) tj[n] = s_T[n*S + j];
#pragma unroll
	for(int m = 0; m < S; ++m){
		double a0 = 0, a1 = 0;
#pragma unroll
		for(int n = 0; n < S; n += 2){
			a0 = fma(s_sum[1 + S + L::tri(m < n ? m : n, m < n ? n : m)], tj[n], a0);
			a1 = fma(s_sum[1 + S + L::tri(m < n + 1 ? m : n + 1, m < n + 1 ? n + 1 : m)], tj[n + 1], a1);
		}
		A[m] = a0 + a1;                                                  // (H_loc T)[m][j]
	}
	if(lane == S){
#pragma unroll
		for(int m = 0; m < S; ++m) A[m] = -g_scale * s_sum[1 + m];       // the sign is undone below
	}
	WarpColPivQR<S, S> qr;
#pragma unroll
	for(int i = 0; i < S; ++i){
		double a0 = 0, a1 = 0;
#pragma unroll
		for(int m = 0; m < S; m += 2){ a0 = fma(s_T[m*S + i], A[m], a0); a1 = fma(s_T[(m + 1)*S + i], A[m + 1], a1); }
		qr.a[i] = -(a0 + a1);
	}
#if MTFB_PROF
	const long long pt1 = clock64();
#endif
	qr.factor_lean(lane, true);
#if MTFB_PROF
	const long long pt2 = clock64();
#endif
	const double x = -qr.solve_fast_cols(lane);                          // state_update = -H^-1 J^T
#if MTFB_PROF
	if(prof && lane == 0){
		atomicAdd((unsigned long long*)prof + 4, (unsigned long long)(pt1 - pt0)); atomicAdd((unsigned long long*)prof + 5, (unsigned long long)(pt2 - pt1));
		atomicAdd((unsigned long long*)prof + 6, (unsigned long long)(clock64() - pt2));
	}
#endif
	if(qr.nonzero_pivots < S) patch_status |= MTFB_PATCH_SINGULAR;
	if(lane < S) s_dp[lane] = x;
	__syncwarp();
}

// entry (k, c) of the update's warp matrix getWarpFromState(dp) (Homography.cc:94-107, Affine.cc:117-131)
template<int SSM> __device__ __forceinline__ double update_entry(const double *dp, int k, int c){
	double u = (k == c) ? 1.0 : 0.0;
	if(SSM == SSM_HOM){
		const int idx = 3 * k + c;
		if(idx < 8) u += dp[idx];
	} else{
		// [[1 + s2, s3, s0], [s4, 1 + s5, s1], [0, 0, 1]]
		const int idx = (k == 0) ? (c == 0 ? 2 : c == 1 ? 3 : 0) : (c == 0 ? 4 : c == 1 ? 5 : 1);
		if(k < 2) u += dp[idx];
	}
	return u;
}
// The tail of a pass once the state update is known, for the forward-compositional searches without
// Levenberg-Marquardt and without an iteration log (what serial_step<.., PRESOLVED> does, spread over the lanes of the
// warp instead of replicated on each): ssm.compositionalUpdate (Homography.cc:73-92, Affine.cc:90-107) with lane i < 9
// on entry i of curr_warp . update, the four corners on lanes 0..3, the corner-change test (NT/FCLK.cc:331-343).
template<int SSM> __device__ __forceinline__ int apply_update_lean(const DevBatch &b, int lane, double f, const double *s_dp,
	double *s_W, double *s_corners, const double *s_init_corners, int &patch_status){
	const int i = lane < 9 ? lane : 0, r = i / 3, c = i - 3 * r;
	double Wn = s_W[3 * r] * update_entry<SSM>(s_dp, 0, c);
	Wn = fma(s_W[3 * r + 1], update_entry<SSM>(s_dp, 1, c), Wn);
	Wn = fma(s_W[3 * r + 2], update_entry<SSM>(s_dp, 2, c), Wn);
	if(SSM == SSM_HOM){
		const double d = __shfl_sync(FULL_MASK, Wn, 8);
		Wn = (lane == 8) ? 1.0 : Wn * rcp_newton(d);
	}
	__syncwarp();
	if(lane < 9) s_W[lane] = Wn;
	__syncwarp();
	double v = 0, nx = 0, ny = 0;
	if(lane < 4){
		const double px = s_init_corners[lane], py = s_init_corners[4 + lane];
		double hx = fma(s_W[1], py, s_W[0] * px) + s_W[2], hy = fma(s_W[4], py, s_W[3] * px) + s_W[5];
		if(SSM == SSM_HOM){
			const double rz = rcp_newton(fma(s_W[7], py, s_W[6] * px) + s_W[8]);
			hx *= rz; hy *= rz;
		}
		nx = hx; ny = hy;
		const double dx = s_corners[lane] - nx, dy = s_corners[4 + lane] - ny;
		v = fma(dx, dx, dy*dy);
	}
	v += __shfl_xor_sync(FULL_MASK, v, 1);
	v += __shfl_xor_sync(FULL_MASK, v, 2);
	const double upd_norm = __shfl_sync(FULL_MASK, v, 0);
	if(lane < 4){ s_corners[lane] = nx; s_corners[4 + lane] = ny; }
	if(!(upd_norm == upd_norm) || !(f == f)) patch_status |= MTFB_PATCH_NAN;
	__syncwarp();
	return upd_norm < b.epsilon ? CTRL_BREAK : CTRL_NEXT;
}

} // namespace f32

} // namespace mtfb
