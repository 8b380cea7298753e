// lk_ssd_mom.cu -- the SSD forward-compositional (FCLK) frame kernel of the F32 precision, second generation:
// COLUMN-FIXED THREADS and MOMENT ACCUMULATORS.
//
// Same path and same numerics contract as lk_ssd_f32.cu (nt::FCLK::update, SM/src/NT/FCLK.cc:171-358: fp32 per-pixel
// arithmetic in patch-local coordinates, bit-exact sampling indices through the fp64 guard band, fp64 reduction / solve /
// compositional update), restructured around two observations:
//
//  * In the frame of the sampling grid, ~u = (u, v, 1) with u, v in [-1, 1], a row of the pixel Jacobian is
//        j_i = p_i(u, v) Gu + q_i(u, v) Gv
//    with (Gu, Gv) the gradient of I(w(~u)) with respect to (u, v) and p_i, q_i signed monomials of degree <= 2
//    (Homography.cc:231-294, Affine.cc:213-242 written in the grid frame; lk_moment_tables.cuh).  J^T J and J^T w are
//    therefore MOMENTS  sum u^a v^b {Gu^2, Gu Gv, Gv^2, w Gu, w Gv}, a + b <= 4.
//  * The grid is a tensor product: u depends on the column only.  A thread that stays in ONE COLUMN for the whole pass
//    accumulates  S_b = sum_rows v^b {..}  -- 18 sums instead of 45 (Homography), 12 fused instructions per pixel instead of
//    ~35, no Jacobian row at all -- and multiplies by its u^a once per pass, before the CTA-wide reduction.  The column
//    also makes three of the six geometry FMAs per-pass constants of the thread.
//
// With 18 accumulators the kernel fits 72 registers: four warps per patch (one per SM sub-partition) at seven patches per
// SM, where the 45-accumulator kernel has two.
//
// THE LOCAL BASIS.  The moments give H_loc = J_loc^T J_loc for J_loc = d I / d q, q the parameters of an update composed in
// the GRID frame: w(~u) = M (I + X(q)) ~u, M = curr_warp . D, D = dlt . N the map from ~u to the reference's homogeneous
// template points (N: [-1, 1]^2 -> the LinSpaced grid).  The reference composes in ITS template frame,
// curr_warp (I + X(p)) D ~u, hence X(p) = D X(q) D^-1 (mod the identity): p = Tinv q with the constant S x S matrix of that
// conjugation, J_loc = J_ref Tinv, and  H_ref = T^T H_loc T, g_ref = T^T g_loc  exactly (fp64), T = Tinv^-1 = the
// conjugation by D^-1.  Both maps are built once per frame by warp 0 (adjoint_maps()).  One reference quirk is reproduced:
// Homography.cc:68 keeps the DLT's un-normalised third row in init_pts_hm, so the reference's dI/dp is the true one divided
// by hz = (D ~u)_2 (identically 1 unless the initial region is a general quadrilateral): the gradient is scaled by 1 / hz.
#include <mutex>
#include <cuda.h>
#include "lk_f32.cuh"
#include "lk_moment_tables.cuh"

namespace mtfb {

using namespace f32;

namespace mom {

// per-pass constants in shared memory (floats)
enum { K_M = 0,            // 9: rows of [M0 - X0 M2; M1 - Y0 M2; M2], M = curr_warp . D
       K_DELTA = 9, K_LOX = 10, K_HIX = 11, K_LOY = 12, K_HIY = 13, K_COUNT = 14 };

// The frame window of this kernel: MOM_WIN_ROWS rows of MOM_WINP floats, loaded by ONE 2-D TMA copy (cp.async.bulk.tensor.2d,
// SASS UTMALDG) from the frame's tensor map with the box's first column rounded down to a multiple of 4 floats (the copy
// faults unless the box starts on a 16-byte boundary: profiles/r02_tma2d_run.log), so the window's column 0 sits at offset
// (ox & 3) of the box and F32_WIN usable columns always fit: 3 + 56 <= 64.  Columns of the box beyond the frame are zero filled
// and never sampled.
constexpr int MOM_WINP = 64, MOM_WIN_ROWS = F32_WIN;
__device__ __forceinline__ bool elect_one(){
	unsigned pred;
	asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
	return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int x, int y, unsigned long long *bar){
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
		:: "r"(smem_u32(smem_dst)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b){
	unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ void fadd2(unsigned long long &d, unsigned long long a){
	asm("add.rn.f32x2 %0, %0, %1;" : "+l"(d) : "l"(a));
}

// position (row, column) in the 3 x 3 update matrix X of state entry k (Homography.cc:94-107, Affine.cc:117-131)
template<int SSM> __device__ __forceinline__ void param_pos(int k, int &r, int &c){
	if(SSM == SSM_HOM){ r = k / 3; c = k - 3 * r; }
	else{ r = (k == 0 || k == 2 || k == 3) ? 0 : 1; c = k < 2 ? 2 : ((k == 2 || k == 4) ? 0 : 1); }
}

// D = dlt . N (N: (u, v) in [-1, 1]^2 -> the LinSpaced sampling grid), and the two conjugation maps.  Warp 0, once per frame.
template<int SSM> __device__ __forceinline__ void adjoint_maps(const DevBatch &b, int lane, const double *s_dlt, double *s_Dt,
	double *s_T, double *s_Tinv){
	constexpr int S = StateSize<SSM>::value;
	const double x_lo = __ldg(b.xv), x_hi = __ldg(b.xv + b.resx - 1), y_lo = __ldg(b.yv), y_hi = __ldg(b.yv + b.resy - 1);
	const double uc = 0.5*(x_lo + x_hi), uh = 0.5*(x_hi - x_lo), vc = 0.5*(y_lo + y_hi), vh = 0.5*(y_hi - y_lo);
	if(lane < 9){
		const int r = lane / 3, c = lane - 3 * r;
		const bool unit = (SSM == SSM_HOM) && b.norm_init;            // hom_normalized_init: init_pts_hm is the grid itself
		const double d0 = unit ? (r == 0 ? 1.0 : 0.0) : s_dlt[3 * r], d1 = unit ? (r == 1 ? 1.0 : 0.0) : s_dlt[3 * r + 1],
			d2 = unit ? (r == 2 ? 1.0 : 0.0) : s_dlt[3 * r + 2];
		double v = c == 0 ? d0*uh : (c == 1 ? d1*vh : fma(d0, uc, fma(d1, vc, d2)));
		if(SSM == SSM_AFF){
			// (launched for parallelograms only: the DLT is affine up to rounding; Affine.cc:81-82 dehomogenises init_pts)
			const double hz = s_dlt[8];
			v = r < 2 ? v / hz : (c == 2 ? 1.0 : 0.0);
		}
		s_Dt[lane] = v;
	}
	__syncwarp();
	Mat3 Dt;
#pragma unroll
	for(int i = 0; i < 9; ++i) Dt.m[i] = s_Dt[i];
	const Mat3 Di = mat3_inverse(Dt);
	if(lane < 2 * S){
		const bool inv = lane >= S;                                   // lanes S .. 2S-1: T, the conjugation by D^-1
		const int k = inv ? lane - S : lane;
		int r, c;
		param_pos<SSM>(k, r, c);
		// L E_k R = (column r of L) (row c of R)
		const Mat3 &Lm = inv ? Di : Dt, &Rm = inv ? Dt : Di;
		double col[3], row[3];
#pragma unroll
		for(int i = 0; i < 3; ++i){ col[i] = 0; row[i] = 0; }
#pragma unroll
		for(int i = 0; i < 3; ++i){
#pragma unroll
			for(int q = 0; q < 3; ++q){ if(q == r) col[i] = Lm.m[3 * i + q]; if(q == c) row[i] = Rm.m[3 * q + i]; }
		}
		const double kappa = (SSM == SSM_HOM) ? col[2] * row[2] : 0.0;
		double *out = inv ? s_T : s_Tinv;
#pragma unroll
		for(int i = 0; i < S; ++i){
			int ri, ci;
			param_pos<SSM>(i, ri, ci);
			double a = 0;
#pragma unroll
			for(int q = 0; q < 3; ++q)
#pragma unroll
			for(int t = 0; t < 3; ++t) if(q == ri && t == ci) a = col[q] * row[t];
			out[i*S + k] = a - (ri == ci ? kappa : 0.0);
		}
	}
}

// Per-pass constants by warp 0 (same error analysis as f32::pass_constants, lk_f32.cuh, with |u|, |v| <= 1 and the grid
// values correctly rounded: the bound there is an upper bound here).
template<int SSM> __device__ __forceinline__ void pass_constants_mom(const DevBatch &b, int lane, const double *W, const double *Dt,
	float *cf, int *ci, double *M64){
	float E, rho = 2.0f;
	int X0 = 0, Y0 = 0;
	bool sane = true;
	double Mi;
	{
		const int i = lane < 9 ? lane : 0, r = i / 3, c = i - 3 * r;
		Mi = fma(W[3 * r + 2], Dt[6 + c], fma(W[3 * r + 1], Dt[3 + c], W[3 * r] * Dt[c]));
		if(SSM == SSM_AFF && r == 2) Mi = (c == 2) ? 1.0 : 0.0;
	}
	if(lane < 9) M64[lane] = Mi;                                 // the fp64 map, for the pixels the fp32 guard band defers
	const float Mf = (float)Mi;
	const float m2 = __shfl_sync(FULL_MASK, Mf, 2), m5 = __shfl_sync(FULL_MASK, Mf, 5), m8 = __shfl_sync(FULL_MASK, Mf, 8);
	const float r8 = rcp_approx(m8);
	const float cx = m2 * r8, cy = m5 * r8;
	sane = (m8 > 1e-30f) && (fabsf(cx) < 2e6f) && (fabsf(cy) < 2e6f);          // the magic-number floor needs |coordinates| < 2^22
	if(sane){ X0 = (int)floorf(cx); Y0 = (int)floorf(cy); }
	const double mz = __shfl_sync(FULL_MASK, Mi, 6 + lane % 3);
	if(lane < 6) Mi = fma(-(double)(lane < 3 ? X0 : Y0), mz, Mi);
	const float Mc = fabsf((float)Mi);
	if(lane < 9) cf[K_M + lane] = (float)Mi;
	float m[9];
#pragma unroll
	for(int i = 0; i < 9; ++i) m[i] = __shfl_sync(FULL_MASK, Mc, i);
	const float Ax = m[0] + m[1] + m[2], Ay = m[3] + m[4] + m[5];
	const float Dmin = m[8] - m[6] - m[7];
	const bool dmin_ok = (Dmin > 1e-30f) && (Dmin < 1e30f);
	const float rD = rcp_approx(dmin_ok ? Dmin : 1.0f);
	E = fmaxf(Ax, Ay) * rD;
	if(SSM == SSM_HOM) rho = (m[8] + m[6] + m[7]) * rD;
	sane = sane && dmin_ok;
	float delta = 2.5f * 5.9604644775390625e-8f * (9.0f + 6.0f * rho) * E * 1.00001f + 2e-6f;
	if(!sane || !(delta < 0.25f)) delta = 2.0f;                  // every pixel takes the fp64 path
	if(lane == 0){
		cf[K_DELTA] = delta;
		cf[K_LOX] = (float)(-X0); cf[K_HIX] = (float)(b.img.w - 2 - X0);
		cf[K_LOY] = (float)(-Y0); cf[K_HIY] = (float)(b.img.h - 2 - Y0);
		ci[0] = X0; ci[1] = Y0;
	}
}

// the thread's sums over its rows: S_b = sum v^b {X = Gu^2, Y = Gu Gv, Z = Gv^2, A = w Gu, B = w Gv}, and sum r^2.
// (X, Y) and (Z, B) live in packed fp32x2 registers (fma.rn.f32x2, SASS FFMA2: one issue slot for two sums).
template<int SSM> struct MomAcc {
	typedef MomentTab<SSM> MT;
	static constexpr int NP2 = MT::NBB;                       // (Z, B) pairs: b < NBB; Z alone above
	unsigned long long p1[3], p2[NP2];
	float y3, zs[MT::NBZ - NP2], a[2], rr;
	__device__ __forceinline__ void clear(){
#pragma unroll
		for(int i = 0; i < 3; ++i) p1[i] = 0ull;
#pragma unroll
		for(int i = 0; i < NP2; ++i) p2[i] = 0ull;
#pragma unroll
		for(int i = 0; i < MT::NBZ - NP2; ++i) zs[i] = 0;
		y3 = 0; a[0] = a[1] = 0; rr = 0;
	}
	// vt = (v, v^2, v^3, v^4)
	__device__ __forceinline__ void add(float Gu, float Gv, float w, float r, const float4 &vt){
		const unsigned long long P1 = fmul2(pack2(Gu, Gu), pack2(Gu, Gv));         // (X, Y)
		const unsigned long long P2 = fmul2(pack2(Gv, Gv), pack2(Gv, w));          // (Z, B)
		const float A = Gu * w;
		float X, Y, Z, B;
		unpack2(P1, X, Y); unpack2(P2, Z, B);
		(void)X; (void)B;
		const float vb[5] = { 1.0f, vt.x, vt.y, vt.z, vt.w };
		fadd2(p1[0], P1); fadd2(p2[0], P2); a[0] += A;
		ffma2(p1[1], pack2(vb[1], vb[1]), P1); ffma2(p2[1], pack2(vb[1], vb[1]), P2); a[1] = fmaf(vb[1], A, a[1]);
		ffma2(p1[2], pack2(vb[2], vb[2]), P1);
		if(NP2 > 2) ffma2(p2[NP2 > 2 ? 2 : 0], pack2(vb[2], vb[2]), P2);
#pragma unroll
		for(int q = NP2; q < MT::NBZ; ++q) zs[q - NP2] = fmaf(vb[q], Z, zs[q - NP2]);
		if(MT::NBY > 3) y3 = fmaf(vb[3], Y, y3);
		rr = fmaf(r, r, rr);
	}
	// S_b of a type: 0 X, 1 Y, 2 Z, 3 A, 4 B
	__device__ __forceinline__ float get(int type, int bb) const{
		float lo, hi;
		if(type == 0){ unpack2(p1[bb], lo, hi); return lo; }
		if(type == 1){ if(bb == 3) return y3; unpack2(p1[bb], lo, hi); return hi; }
		if(type == 2){ if(bb >= NP2) return zs[bb - NP2]; unpack2(p2[bb], lo, hi); return lo; }
		if(type == 3) return a[bb];
		unpack2(p2[bb], lo, hi); return hi;
	}
};

// the registers a thread keeps through a pass
struct PassRegs {
	float m0, m1, m3, m4, m6, m7;     // chain rule: d w / d (u, v) = (M_row - M2 w) / D
	float cx, cy, cD;                 // M_row . (u, ., 1) with this thread's u: the numerators / denominator are one FMA in v away
	float cZ, d7;                     // hz = d7 v + cZ (general quadrilaterals only)
	float delta, hi;
	float lox, hix, loy, hiy;         // range test of the generic variant
	// where the four neighbours are read from.  Window variant: base = element (Y0, X0) of the shared-memory window (the
	// window holds the patch's hull, so the patch centre is inside it); generic variant: base = start of the array (frame
	// or window) and (Xr, Yr) = the integer origin relative to it -- the origin itself may lie outside the array
	const float *base; int pitch, Xr, Yr;
	int X0, Y0;
};

// chain rule of one pixel once its sample is known (shared by the three evaluation tiers): gradient with respect to
// (u, v), scaled as the reference's dI/dp is (Homography.cc:250-262 in the grid frame), and the residual; valid = false gives
// zeros.  cZ: the column's part of hz (general quadrilaterals only)
template<int SSM> __device__ __forceinline__ void mom_pixel_grad(const PassRegs &k, float cZ, float v, bool proj, float wxl, float wyl,
	float invD, float gx, float gy, float val, float i0, bool valid, float &Gu, float &Gv, float &r){
	float gs = valid ? invD : 0.0f;
	if(SSM == SSM_HOM && proj) gs *= rcp_approx(fmaf(k.d7, v, cZ));
	const float gxd = gx * gs, gyd = gy * gs;
	if(SSM == SSM_HOM){
		Gu = fmaf(fmaf(-k.m6, wxl, k.m0), gxd, fmaf(-k.m6, wyl, k.m3) * gyd);
		Gv = fmaf(fmaf(-k.m7, wxl, k.m1), gxd, fmaf(-k.m7, wyl, k.m4) * gyd);
	} else{
		Gu = fmaf(k.m0, gxd, k.m3 * gyd);
		Gv = fmaf(k.m1, gxd, k.m4 * gyd);
	}
	r = valid ? val - i0 : 0.0f;                                                  // I_diff (SSDBase.cc:78)
}
// ... and its sums
template<int SSM> __device__ __forceinline__ void mom_pixel_back(const PassRegs &k, const float4 &vt, bool proj, float wxl, float wyl,
	float invD, float gx, float gy, float val, float i0, bool valid, MomAcc<SSM> &acc){
	float Gu, Gv, r;
	mom_pixel_grad<SSM>(k, k.cZ, vt.x, proj, wxl, wyl, invD, gx, gy, val, i0, valid, Gu, Gv, r);
	acc.add(Gu, Gv, -r, r, vt);                                                   // df_dIt = -I_diff (SSDBase.cc:115-121)
}

// TIER 1.  One pixel of the thread's column, fp32: warped point, guard-band decision, bilinear sample + gradient.
// WIN: the neighbours come from the shared-memory window that covers the patch's whole hull (no range test, compile-time
// pitch); else from whatever k.base points to, range-tested.  Returns false if the pixel must be re-evaluated (it then
// has contributed nothing).
template<int SSM, bool WIN> __device__ __forceinline__ bool mom_pixel(const PassRegs &k, const float *base, const float4 &vt, bool proj,
	float i0, MomAcc<SSM> &acc){
	const float v = vt.x;
	float invD, wxl, wyl;
	if(SSM == SSM_HOM){
		invD = rcp_approx(fmaf(k.m7, v, k.cD));
		wxl = fmaf(k.m1, v, k.cx) * invD;
		wyl = fmaf(k.m4, v, k.cy) * invD;
	} else{
		invD = 1.0f;
		wxl = fmaf(k.m1, v, k.cx);
		wyl = fmaf(k.m4, v, k.cy);
	}
	float fx, fy; int ix, iy;
	fast_floor(wxl, fx, ix);
	fast_floor(wyl, fy, iy);
	const float dx = wxl - fx, dy = wyl - fy;
	bool fast = (dx >= k.delta) && (dx <= k.hi) && (dy >= k.delta) && (dy <= k.hi);          // written so that NaN fails
	if(!WIN) fast = fast && (fx >= k.lox) && (fx <= k.hix) && (fy >= k.loy) && (fy <= k.hiy);
	const int pitch = WIN ? MOM_WINP : k.pitch;
	const int off = fast ? (WIN ? iy*pitch + ix : (k.Yr + iy)*pitch + (k.Xr + ix)) : 0;
	const float *r0 = base + off, *r1 = r0 + pitch;
	const float p00 = r0[0], p01 = r0[1], p10 = r1[0], p11 = r1[1];
	const float t0 = p01 - p00, t1 = p11 - p10;
	const float top = fmaf(dx, t0, p00), bot = fmaf(dx, t1, p10);
	const float gy = bot - top;                                                   // (1 - dx)(p10 - p00) + dx (p11 - p01)
	const float val = fmaf(dy, gy, top);
	const float gx = fmaf(dy, t1 - t0, t0);                                       // (1 - dy)(p01 - p00) + dy (p11 - p10)
	mom_pixel_back<SSM>(k, vt, proj, wxl, wyl, invD, gx, gy, val, i0, fast, acc);
	return fast;
}

// TIER 2.  A pixel whose fp32 coordinate fell inside the guard band of a cell boundary: the same map in fp64,
// x = (M (u, v, 1))_x / (M (u, v, 1))_z with M = curr_warp . D as pass_constants_mom formed it.  This differs from the
// reference's own fp64 evaluation order by a few ulp of the coordinate (< 1e-12 px), so (int)x and the choice between the
// cell's slope and the reference's straddling finite difference (imgUtils.cc:233-254, |frac| < grad_eps) are the reference's
// whenever frac(x), frac(y) are farther than d2 = 2 grad_eps + 1e-9 from 0 and 1; then the sample is the cell's bilinear
// patch like tier 1, from the frame in global memory.  Returns false for the rest (on or next to the pixel lattice, outside
// the frame): tier 3.
struct MidArgs { Image img; const double *xv, *yv, *M; double uc, uh, vc, vh, d2; int X0, Y0; };
template<int SSM> __device__ __forceinline__ bool mom_pixel_mid(const PassRegs &k, const MidArgs &a, float cZ, float v32, bool proj,
	int row, int col, float i0, float &Gu, float &Gv, float &r){
	const double u = (__ldg(a.xv + col) - a.uc) / a.uh, v = (__ldg(a.yv + row) - a.vc) / a.vh;
	const double nx = fma(a.M[0], u, fma(a.M[1], v, a.M[2])), ny = fma(a.M[3], u, fma(a.M[4], v, a.M[5]));
	double rd = 1.0, x = nx, y = ny;
	if(SSM == SSM_HOM){
		rd = 1.0 / fma(a.M[6], u, fma(a.M[7], v, a.M[8]));
		x = nx * rd; y = ny * rd;
	}
	const double fx = floor(x), fy = floor(y), dx = x - fx, dy = y - fy;
	const bool ok = (dx >= a.d2) && (dx <= 1.0 - a.d2) && (dy >= a.d2) && (dy <= 1.0 - a.d2) &&
		(fx >= 0.0) && (fx <= (double)(a.img.w - 2)) && (fy >= 0.0) && (fy <= (double)(a.img.h - 2));
	if(!ok) return false;
	const float *r0 = a.img.data + (size_t)(int)fy*a.img.pitch + (int)fx, *r1 = r0 + a.img.pitch;
	const float p00 = __ldg(r0), p01 = __ldg(r0 + 1), p10 = __ldg(r1), p11 = __ldg(r1 + 1);
	const float dxf = (float)dx, dyf = (float)dy;
	const float t0 = p01 - p00, t1 = p11 - p10;
	const float top = fmaf(dxf, t0, p00), bot = fmaf(dxf, t1, p10);
	const float gy = bot - top;
	const float val = fmaf(dyf, gy, top);
	const float gx = fmaf(dyf, t1 - t0, t0);
	mom_pixel_grad<SSM>(k, cZ, v32, proj, (float)(x - a.X0), (float)(y - a.Y0), (float)rd, gx, gy, val, i0, true, Gu, Gv, r);
	return true;
}

// TIER 3.  The same pixel through the reference-exact fp64 functions (lk_f32.cuh exact_pixel), then the fp32 chain rule and sums.
template<int SSM> __device__ __forceinline__ void mom_pixel_exact(const DevBatch &b, const PassRegs &k, float cZ, float v32, bool proj,
	const double *s_dlt, const double *s_W, int row, int col, float i0, float &Gu, float &Gv, float &r){
	ExactArgs a;
	a.img = b.img; a.xv = b.xv; a.yv = b.yv; a.s_dlt = s_dlt; a.s_W = s_W; a.grad_eps = b.grad_eps; a.grad_mult = b.grad_mult;
	a.norm_init = b.norm_init; a.X0 = k.X0; a.Y0 = k.Y0;
	const ExactOut e = exact_pixel<SSM>(a, row, col);
	mom_pixel_grad<SSM>(k, cZ, v32, proj, e.wxl, e.wyl, e.invD, e.gx, e.gy, e.val, i0, true, Gu, Gv, r);
}

} // namespace mom

using namespace mom;

#if MTFB_PROF
#define MOM_PROF_T(k) const long long prof_t##k = clock64();
#define MOM_PROF_ADD() if(tid == 0){ \
	atomicAdd((unsigned long long*)b.n_iters_prof + 8, (unsigned long long)(prof_t1 - prof_t0)); \
	atomicAdd((unsigned long long*)b.n_iters_prof + 9, (unsigned long long)(prof_t2 - prof_t1)); \
	atomicAdd((unsigned long long*)b.n_iters_prof + 10, (unsigned long long)(prof_t3 - prof_t2)); \
	atomicAdd((unsigned long long*)b.n_iters_prof + 11, (unsigned long long)(prof_t4 - prof_t3)); \
	atomicAdd((unsigned long long*)b.n_iters_prof + 12, (unsigned long long)(prof_t5 - prof_t4)); }
#define MOM_PROF_KERNEL_T(k) const long long prof_k##k = clock64();
#define MOM_PROF_KERNEL_ADD() if(tid == 0){ \
	atomicAdd((unsigned long long*)b.n_iters_prof + 13, (unsigned long long)(prof_k1 - prof_k0)); \
	atomicAdd((unsigned long long*)b.n_iters_prof + 14, (unsigned long long)(prof_k2 - prof_k1)); \
	atomicAdd((unsigned long long*)b.n_iters_prof + 15, (unsigned long long)(clock64() - prof_k2)); \
	atomicMax((unsigned long long*)b.n_iters_prof + 3, (unsigned long long)(clock64() - prof_k0)); }
#else
#define MOM_PROF_T(k)
#define MOM_PROF_ADD()
#define MOM_PROF_KERNEL_T(k)
#define MOM_PROF_KERNEL_ADD()
#endif

// work: per thread (column, first row, rows, -) -- which pixels of the patch the thread owns (mtfb_api.cu builds the table
// from resx, resy and T; a thread never leaves its column)
template<int SSM, int T, int MINB>
__global__ void __launch_bounds__(T, MINB) ssd_fclk_mom_kernel(DevBatch b, const int4 *__restrict__ work, unsigned tmpl_bytes,
	unsigned win_elems, unsigned win_off, const __grid_constant__ CUtensorMap frame_map, int use_tma){
	constexpr int S = StateSize<SSM>::value;
	constexpr int SM = SM_FCLK;
	typedef AccLayout<S> L;
	typedef MomentTab<SSM> MT;
	constexpr int NM = MT::NM, NR = NM + 1;                // moments + sum r^2
	const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	MOM_PROF_KERNEL_T(0)
	constexpr int PART = NR > 48 ? NR : 48;                 // per warp: NR partial sums, or 32 x 3 floats of pooled deferred pixels
	__shared__ double s_part[(T / 32) * PART];
	__shared__ double s_sum[L::NA];
	__shared__ double s_W[9], s_corners[8], s_init_corners[8], s_dlt[9], s_Dt[9];
	__shared__ double s_J[S], s_Hc[S*S], s_Hl[S*S], s_A[S*S], s_T[S*S], s_Tinv[S*S], s_x[S], s_dp[S];
	__shared__ double s_M64[9], s_gc[4];
	__shared__ float s_cf[K_COUNT];
	__shared__ int s_ci[2], s_wi[6];
	__shared__ int s_ctrl;
	extern __shared__ __align__(128) float s_dyn[];
	__shared__ __align__(8) unsigned long long s_bar, s_bar_win;
	const bool use_smem = tmpl_bytes != 0;
	if(use_smem && tid == 0) mbar_init(&s_bar, 1);
	if(tid == 0) mbar_init(&s_bar_win, 1);
	unsigned win_phase = 0;
	if(tid < 9){ s_W[tid] = b.warp[(size_t)p * 9 + tid]; s_dlt[tid] = b.dlt[(size_t)p * 9 + tid]; }
	if(tid < 8){ s_corners[tid] = b.corners[(size_t)p * 8 + tid]; s_init_corners[tid] = b.init_corners[(size_t)p * 8 + tid]; }
	cta_sync<T>();
	if(use_smem && tid == 0){
		mbar_expect_tx(&s_bar, tmpl_bytes);
		bulk_copy_g2s(s_dyn, b.I0f + (size_t)p*b.I0f_stride, tmpl_bytes, &s_bar);      // TMA, SASS UBLKCP
	}
	// dynamic shared memory: template row | per-row powers of v | frame window
	float4 *s_vtab = reinterpret_cast<float4*>(s_dyn + (use_smem ? b.I0f_stride : 0));
	float *s_win = s_dyn + win_off;                          // 128-byte aligned (TMA destination)
	const double y_lo = __ldg(b.yv), y_hi = __ldg(b.yv + b.resy - 1), x_lo = __ldg(b.xv), x_hi = __ldg(b.xv + b.resx - 1);
	for(int r = tid; r < b.resy; r += T){
		const float v = (float)((__ldg(b.yv + r) - 0.5*(y_lo + y_hi)) / (0.5*(y_hi - y_lo)));
		const float v2 = v*v;
		s_vtab[r] = make_float4(v, v2, v2*v, v2*v2);
	}
	if(warp == 0){
		adjoint_maps<SSM>(b, lane, s_dlt, s_Dt, s_T, s_Tinv);
		__syncwarp();
		pass_constants_mom<SSM>(b, lane, s_W, s_Dt, s_cf, s_ci, s_M64);
		if(lane == 0){ s_wi[4] = 0; window_decide(b, s_corners, win_elems != 0, s_wi); }
	}
	// this thread's pixels: rows row0 .. row0 + nrows - 1 of column col
	const int4 wk = __ldg(work + tid);
	const int col = wk.x, row0 = wk.y, nrows = wk.z;
	const float u = (float)((__ldg(b.xv + col) - 0.5*(x_lo + x_hi)) / (0.5*(x_hi - x_lo)));
	if(tid == 0){ s_gc[0] = 0.5*(x_lo + x_hi); s_gc[1] = 0.5*(x_hi - x_lo); s_gc[2] = 0.5*(y_lo + y_hi); s_gc[3] = 0.5*(y_hi - y_lo); }
	cta_sync<T>();
	// general quadrilateral: the DLT's third row varies over the patch (relative 1e-9: invisible in fp32 otherwise)
	const bool proj = (SSM == SSM_HOM) && !(fabs(s_Dt[6]) <= 1e-9*fabs(s_Dt[8]) && fabs(s_Dt[7]) <= 1e-9*fabs(s_Dt[8]));
	const float *I0 = b.I0f + (size_t)p*b.I0f_stride;
	if(use_smem) mbar_wait(&s_bar, 0);
	const float *tmpl = (use_smem ? (const float*)s_dyn : I0) + row0*b.resx + col;
	const bool lean_tail = (hessian_select<SM>(b.hess_type) == 0) && !b.leven_marq && !b.log;
	const bool local_solve = b.f32_local_solve && lean_tail;
	LMState lm = { 0.0, b.lm_delta_init, 0.0, false };
	int iter_id = 0, n_passes = 0, patch_status = 0;
	double f = 0;
	MOM_PROF_KERNEL_T(1)
#if MTFB_PROF == 2
	int prof_slow = 0;
#endif
	while(iter_id < b.max_iters){
		MOM_PROF_T(0)
		PassRegs k;
		{
			const float m2 = s_cf[K_M + 2], m5 = s_cf[K_M + 5], m8 = s_cf[K_M + 8];
			k.m0 = s_cf[K_M + 0]; k.m1 = s_cf[K_M + 1]; k.m3 = s_cf[K_M + 3]; k.m4 = s_cf[K_M + 4];
			k.m6 = s_cf[K_M + 6]; k.m7 = s_cf[K_M + 7];
			k.cx = fmaf(k.m0, u, m2); k.cy = fmaf(k.m3, u, m5); k.cD = fmaf(k.m6, u, m8);
			k.d7 = (float)s_Dt[7]; k.cZ = fmaf((float)s_Dt[6], u, (float)s_Dt[8]);
			k.delta = s_cf[K_DELTA]; k.hi = 1.0f - k.delta;
			k.lox = s_cf[K_LOX]; k.hix = s_cf[K_HIX]; k.loy = s_cf[K_LOY]; k.hiy = s_cf[K_HIY];
			k.X0 = s_ci[0]; k.Y0 = s_ci[1];
			k.base = b.img.data; k.pitch = b.img.pitch; k.Xr = k.X0; k.Yr = k.Y0;
		}
		const bool use_win = s_wi[2] != 0, win_full = use_win && s_wi[5] != 0;
		if(use_win){
			const int ox = s_wi[0], oy = s_wi[1];
			if(s_wi[3]){
				if(use_tma){
					// one elected lane of warp 0 issues the tile copy; everybody waits on its mbarrier
					cta_sync<T>();                                   // the previous pass's readers are done with the window
					if(warp == 0 && elect_one()){
						mbar_expect_tx(&s_bar_win, MOM_WINP*MOM_WIN_ROWS*(unsigned)sizeof(float));
						tma_load_2d(s_win, &frame_map, ox & ~3, oy, &s_bar_win);
					}
					mbar_wait(&s_bar_win, win_phase);
					win_phase ^= 1u;
				} else{
					for(int i = tid; i < F32_WIN*F32_WIN; i += T){
						const int r = i / F32_WIN, c = i - r*F32_WIN;
						s_win[r*MOM_WINP + (ox & 3) + c] = __ldg(b.img.data + (size_t)(oy + r)*b.img.pitch + ox + c);
					}
					cta_sync<T>();
				}
			}
			k.Xr = k.X0 - ox; k.Yr = k.Y0 - oy; k.pitch = MOM_WINP; k.base = s_win + (ox & 3);
			k.lox = (float)(ox - k.X0); k.hix = (float)(ox + F32_WIN - 2 - k.X0);
			k.loy = (float)(oy - k.Y0); k.hiy = (float)(oy + F32_WIN - 2 - k.Y0);
		}
		MomAcc<SSM> acc;
		acc.clear();
		// The loop is branch-free: a pixel that must take the fp64 path (inside the guard band of a cell boundary, on the
		// pixel lattice, neighbours outside the array) contributes nothing and sets a bit; the bits are worked off after
		// each run of <= 32 rows.
		const int resx = b.resx;
		for(int c0 = 0; c0 < nrows; c0 += 32){
			const int n = nrows - c0 < 32 ? nrows - c0 : 32;
			unsigned slow = 0u;
			if(win_full && use_smem){
				// everything the pixel reads comes from shared memory, through pointers the compiler can see are shared
				// (32-bit addresses, LDS): template row, powers of v, frame window at element (Y0, X0)
				const float *tp = s_dyn + (row0 + c0)*resx + col;
				const float4 *vp = reinterpret_cast<const float4*>(s_dyn + b.I0f_stride) + row0 + c0;
				const float *wb = s_dyn + win_off + (s_wi[0] & 3) + k.Yr*MOM_WINP + k.Xr;
				for(int i = 0; i < n; ++i, tp += resx){
					const bool fast = mom_pixel<SSM, true>(k, wb, vp[i], proj, *tp, acc);
					slow |= (fast ? 0u : 1u) << i;
				}
			} else{
				const float *tp = tmpl + c0*resx;
				for(int i = 0; i < n; ++i, tp += resx){
					const bool fast = mom_pixel<SSM, false>(k, k.base, s_vtab[row0 + c0 + i], proj, *tp, acc);
					slow |= (fast ? 0u : 1u) << i;
				}
			}
#if MTFB_PROF == 2
			prof_slow += __popc(slow);
#endif
			// THE DEFERRED PIXELS, SHARED OUT OVER THE WARP.  A sampling-grid row or column that lands within the guard band of
			// the pixel lattice defers all its pixels at once; a column's fall on the two or three threads that own it (up to 25
			// evaluations of ~1 K cycles one after the other: such a patch used to set the duration of the whole launch,
			// profiles/r02_patch_spread.txt).  Any lane can evaluate any pixel of the patch -- only the sums are tied to the
			// column -- so the warp pools its deferred pixels, lane j evaluates the j-th of them, and the owners add the results
			// to their sums in the order they would have used alone.
			if(__any_sync(FULL_MASK, slow != 0u)){
				float *res = reinterpret_cast<float*>(s_part + warp*PART);       // 32 x (Gu, Gv, r): free until the reduction
				const unsigned cnt = (unsigned)__popc(slow);
				unsigned incl = cnt;
#pragma unroll
				for(int off = 1; off < 32; off <<= 1){
					const unsigned t = __shfl_up_sync(FULL_MASK, incl, off);
					if(lane >= off) incl += t;
				}
				const unsigned total = __shfl_sync(FULL_MASK, incl, 31), first = incl - cnt;
				MidArgs ma;
				ma.img = b.img; ma.xv = b.xv; ma.yv = b.yv; ma.M = s_M64; ma.uc = s_gc[0]; ma.uh = s_gc[1]; ma.vc = s_gc[2]; ma.vh = s_gc[3];
				ma.d2 = 2.0*b.grad_eps + 1e-9; ma.X0 = k.X0; ma.Y0 = k.Y0;
				for(unsigned start = 0; start < total; start += 32){
					// whose pixel is number start + lane?  the last lane whose run begins at or before it
					const unsigned want = start + lane;
					int owner = 0;
					for(int l = 0; l < 32; ++l){
						const unsigned fl = __shfl_sync(FULL_MASK, first, l), cl = __shfl_sync(FULL_MASK, cnt, l);
						if(cl != 0u && fl <= want) owner = l;
					}
					const unsigned o_slow = __shfl_sync(FULL_MASK, slow, owner), o_first = __shfl_sync(FULL_MASK, first, owner);
					const int o_row0 = __shfl_sync(FULL_MASK, row0, owner), o_col = __shfl_sync(FULL_MASK, col, owner);
					const float o_cZ = __shfl_sync(FULL_MASK, k.cZ, owner);
					if(want < total){
						const int bit = (int)__fns(o_slow, 0u, (int)(want - o_first) + 1);
						const int row = o_row0 + c0 + bit;
						const float i0 = (use_smem ? (const float*)s_dyn : I0)[row*resx + o_col];
						const float v32 = s_vtab[row].x;
						float Gu, Gv, r;
						if(!mom_pixel_mid<SSM>(k, ma, o_cZ, v32, proj, row, o_col, i0, Gu, Gv, r))
							mom_pixel_exact<SSM>(b, k, o_cZ, v32, proj, s_dlt, s_W, row, o_col, i0, Gu, Gv, r);
						res[3 * lane] = Gu; res[3 * lane + 1] = Gv; res[3 * lane + 2] = r;
					}
					__syncwarp();
					{
						unsigned m = slow, j = first;
						while(m){
							const int i = c0 + __ffs((int)m) - 1;
							m &= m - 1;
							if(j >= start && j < start + 32u){
								const float *q = res + 3 * (j - start);
								acc.add(q[0], q[1], -q[2], q[2], s_vtab[row0 + i]);
							}
							++j;
						}
					}
					__syncwarp();
				}
			}
		}
		// moments: u^a S_b
		float mv[NR];
		{
			const float u2 = u*u;
			const float ub[5] = { 1.0f, u, u2, u2*u, u2*u2 };
#pragma unroll
			for(int m = 0; m < NM; ++m) mv[m] = ub[MT::mom_a(m)] * acc.get(MT::mom_type(m), MT::mom_b(m));
			mv[NM] = acc.rr;
		}
		MOM_PROF_T(1)
		// CTA-wide sums (fp32 butterfly inside a warp, fp64 across warps) and the entries of J_loc^T J_loc, J_loc^T w from them
		{
			int idx[3];
			warp_reduce_scatter<NR>(mv, lane, idx);
			double *dst = s_part + warp*PART;
			if(idx[0] >= 0) dst[idx[0]] = (double)mv[0];
			if(idx[1] >= 0) dst[idx[1]] = (double)mv[1];
			if(NR > 64 && idx[2] >= 0) dst[idx[2]] = (double)mv[2];
			cta_sync<T>();
			for(int e = tid; e < L::NA; e += T){
				double a = 0;
				if(e == 0){
#pragma unroll
					for(int w = 0; w < T / 32; ++w) a += s_part[w*PART + NM];
				} else if(e < 1 + S){
					const signed char (*g)[2][2] = MT::grad_dev();
#pragma unroll
					for(int t = 0; t < 2; ++t){
						const int mi = g[e - 1][t][0];
						if(mi >= 0){
							double s = 0;
#pragma unroll
							for(int w = 0; w < T / 32; ++w) s += s_part[w*PART + mi];
							a = fma((double)g[e - 1][t][1], s, a);
						}
					}
				} else{
					const signed char (*h)[MT::NT][2] = MT::hess_dev();
#pragma unroll
					for(int t = 0; t < MT::NT; ++t){
						const int mi = h[e - 1 - S][t][0];
						if(mi >= 0){
							double s = 0;
#pragma unroll
							for(int w = 0; w < T / 32; ++w) s += s_part[w*PART + mi];
							a = fma((double)h[e - 1 - S][t][1], s, a);
						}
					}
				}
				s_sum[e] = a;
			}
			cta_sync<T>();
		}
		++n_passes;
		MOM_PROF_T(2)
		// pass-local Hessian, no Levenberg-Marquardt, no iteration log: warp 0 solves (local basis if asked for and safe, else
		// the reference's QR in the reference's parameters) and applies the update with the work spread over its lanes
		if(lean_tail){
			if(warp == 0){
				bool solved = false;
				if(local_solve) solved = solve_local<S>(lane, s_sum, s_Tinv, 1.0, s_x, s_dp);
				if(!solved) solve_reference_warp<S>(lane, s_sum, s_T, 1.0, s_dp, patch_status, b.n_iters_prof);
				MOM_PROF_T(3)
				f = -s_sum[0] / 2;
				const int ctrl = apply_update_lean<SSM>(b, lane, f, s_dp, s_W, s_corners, s_init_corners, patch_status);
				if(lane == 0) s_ctrl = ctrl;
				__syncwarp();
				MOM_PROF_T(4)
				if(ctrl != CTRL_BREAK){
					pass_constants_mom<SSM>(b, lane, s_W, s_Dt, s_cf, s_ci, s_M64);
					if(lane == 0) window_decide(b, s_corners, win_elems != 0, s_wi);
				}
				MOM_PROF_T(5)
				MOM_PROF_ADD()
			}
		} else{
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
				s_J[tid] = a;
			}
			cta_sync<T>();
			if(warp == 0){
				MOM_PROF_T(3)
				f = -s_sum[0] / 2;
				const int ctrl = serial_step<SSM, SM, false>(b, p, lane, iter_id, n_passes, f, s_J, s_Hc, s_W, s_corners, s_init_corners,
					lm, patch_status);
				if(lane == 0) s_ctrl = ctrl;
				__syncwarp();
				MOM_PROF_T(4)
				if(ctrl != CTRL_BREAK){
					pass_constants_mom<SSM>(b, lane, s_W, s_Dt, s_cf, s_ci, s_M64);
					if(lane == 0) window_decide(b, s_corners, win_elems != 0, s_wi);
				}
				MOM_PROF_T(5)
				MOM_PROF_ADD()
			}
		}
		cta_sync<T>();
		const int ctrl = s_ctrl;
		if(ctrl == CTRL_BREAK) break;
		if(counts_as_iteration<SM>(ctrl, b.nt_semantics)) ++iter_id;
	}
	MOM_PROF_KERNEL_T(2)
	if(warp == 0) store_patch_state<SSM>(b, p, lane, s_W, s_corners, f, n_passes, patch_status);
	MOM_PROF_KERNEL_ADD()
#if MTFB_PROF == 2
	// per-patch diagnostics through the similarity getter: cycles of this CTA + (pixel evaluations deferred to fp64) / 2^20
	prof_slow = __reduce_add_sync(0xffffffffu, prof_slow);
	__shared__ int s_prof_slow;
	if(tid == 0) s_prof_slow = 0;
	__syncthreads();
	if(lane == 0) atomicAdd(&s_prof_slow, prof_slow);
	__syncthreads();
	if(tid == 0){
		b.f[p] = (double)(clock64() - prof_k0) + (double)s_prof_slow / 1048576.0;
		unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
		unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
		b.n_iters[p] = (int)smid;
		b.am_scal[(size_t)p * 8] = (double)gt;                       // end time, ns
	}
#endif
}

// ------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------
template<int SSM, int T, int MINB> static cudaError_t launch_mom_one(const DevBatch &b, const int4 *d_work, const CUtensorMap *frame_map,
	cudaStream_t st){
	size_t tmpl = (size_t)b.I0f_stride*sizeof(float);
	if(tmpl*MINB > 96 * 1024) tmpl = 0;                   // large templates are read through L2
	const size_t vtab = (size_t)b.resy*sizeof(float4);
	const size_t win_off = (tmpl + vtab + 127) & ~(size_t)127;               // the TMA destination is 128-byte aligned
	size_t win = (size_t)MOM_WIN_ROWS*MOM_WINP*sizeof(float);
	cudaFuncAttributes fa;
	cudaError_t e = cudaFuncGetAttributes(&fa, ssd_fclk_mom_kernel<SSM, T, MINB>);
	if(e != cudaSuccess) return e;
	if((fa.sharedSizeBytes + win_off + win + 1024)*MINB > 227 * 1024) win = 0;
	const size_t dyn = win_off + win;
	// the shared-memory carve-out follows the largest footprint seen so far (attributes are per function and sticky)
	static size_t configured_dev[64] = {};
	static std::mutex mu;                                  // contexts of different host threads may share an instantiation
	std::lock_guard<std::mutex> lock(mu);
	int dev = 0;
	cudaGetDevice(&dev);
	size_t &configured = configured_dev[dev & 63];
	if(dyn > configured){
		e = cudaFuncSetAttribute(ssd_fclk_mom_kernel<SSM, T, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
		if(e != cudaSuccess) return e;
		const size_t per_cta = fa.sharedSizeBytes + dyn + 1024;
		const int carve = (int)((per_cta*MINB * 100 + 228 * 1024 - 1) / (228 * 1024)) + 2;
		e = cudaFuncSetAttribute(ssd_fclk_mom_kernel<SSM, T, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, carve > 100 ? 100 : carve);
		if(e != cudaSuccess) return e;
		configured = dyn;
	}
	static const CUtensorMap no_map = {};
	ssd_fclk_mom_kernel<SSM, T, MINB><<<b.P, T, dyn, st>>>(b, d_work, (unsigned)tmpl, (unsigned)(win / sizeof(float)),
		(unsigned)(win_off / sizeof(float)), frame_map ? *frame_map : no_map, frame_map ? 1 : 0);
	return cudaGetLastError();
}

#ifndef MTFB_MOM_MINB128
#define MTFB_MOM_MINB128 7
#endif
cudaError_t launch_update_ssd_mom(int ssm, int threads, const DevBatch &b, const int4 *d_work, const void *frame_tensor_map, cudaStream_t st){
	const CUtensorMap *fm = static_cast<const CUtensorMap*>(frame_tensor_map);
	if(ssm == SSM_HOM){
		switch(threads){
		case 32: return launch_mom_one<SSM_HOM, 32, 16>(b, d_work, fm, st);
		case 64: return launch_mom_one<SSM_HOM, 64, 8>(b, d_work, fm, st);
		case 128: return launch_mom_one<SSM_HOM, 128, MTFB_MOM_MINB128>(b, d_work, fm, st);
		case 256: return launch_mom_one<SSM_HOM, 256, 3>(b, d_work, fm, st);
		default: return cudaErrorInvalidValue;
		}
	}
	switch(threads){
	case 32: return launch_mom_one<SSM_AFF, 32, 16>(b, d_work, fm, st);
	case 64: return launch_mom_one<SSM_AFF, 64, 8>(b, d_work, fm, st);
	case 128: return launch_mom_one<SSM_AFF, 128, MTFB_MOM_MINB128>(b, d_work, fm, st);
	case 256: return launch_mom_one<SSM_AFF, 256, 3>(b, d_work, fm, st);
	default: return cudaErrorInvalidValue;
	}
}

} // namespace mtfb
