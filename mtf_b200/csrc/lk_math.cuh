// lk_math.cuh -- per-pixel arithmetic of the Lucas-Kanade hot path (host+device inline functions).
//
// Everything in here is plain fp64 written in the reference's operation order and compiled with
// -fmad=false (device) / -ffp-contract=off (host unit tests), so that warped coordinates, sampling
// indices and interpolated pixel values come out bit-identical to the reference's Eigen path.
// Fused multiply-adds are used only where they are spelled fma() explicitly (reductions).
//
// Reference functions restated (paths relative to the MTF tree):
//   sample_pixel        Utilities/include/mtf/Utilities/imgUtils.h:51-113  getPixVal<Linear,Constant>
//   sample_pixel_grad   Utilities/src/imgUtils.cc:233-254 getImgGrad -- the eps -> 0 limit of its central
//                       finite difference, including what happens at integer coordinates and image borders
//   Homography / Affine geometry  SSM/src/ProjectiveBase.cc:20-49, SSM/src/Homography.cc:50-132,231-294,
//                       SSM/src/Affine.cc:64-150,213-242, Utilities/include/mtf/Utilities/warpUtils.h:9-21
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define MTFB_HD __host__ __device__ __forceinline__
#else
#define MTFB_HD inline
#endif

namespace mtfb {

enum { AM_SSD = 0, AM_NCC = 1, AM_MI = 2 };
enum { SSM_HOM = 0, SSM_AFF = 1, SSM_TRANS = 2 };   // Homography, Affine, Translation (SSM/src/Translation.cc: SSD, F64)
enum { SM_ESM = 0, SM_FCLK = 1, SM_ICLK = 2, SM_PF = 3, SM_FALK = 4, SM_IALK = 5 };

struct Image {
	const float *data;   // pitched, row-major
	int h, w, pitch;     // pitch in elements
	double hd, wd;       // (double)h, (double)w: the comparisons of imgUtils.h:51-53 promote the ints every call
};
MTFB_HD Image make_image(const float *data, int h, int w, int pitch){
	Image im; im.data = data; im.h = h; im.w = w; im.pitch = pitch; im.hd = h; im.wd = w; return im;
}

// imgUtils.h:51-53
MTFB_HD bool check_overflow(double x, double y, double h, double w){
	// written so that a NaN coordinate counts as outside: on x86 the reference's (int)NaN is INT_MIN, which
	// its second checkOverflow(lx, ly) call rejects (imgUtils.h:105), with the same result
	return !((x >= 0) && (x < w) && (y >= 0) && (y < h));
}

#if defined(__CUDA_ARCH__)
#define MTFB_LDG(p) __ldg(p)
#else
#define MTFB_LDG(p) (*(p))
#endif

// IEEE-rounded reciprocal and a correctly rounded quotient built on it (Markstein): with rb = RN(1/b),
// q = RN(a*rb), r = a - b*q (exact in one fma), RN(q + r*rb) is the correctly rounded a/b -- the same bits the
// reference's a/b produces (checked on 4e8 random operand pairs, tests/test_host_math.py keeps a smaller run),
// at 3 fp64 instructions per quotient once the reciprocal is shared, instead of a ~20-instruction division.
MTFB_HD double ieee_rcp(double b){
#if defined(__CUDA_ARCH__)
	return __drcp_rn(b);
#else
	return 1.0 / b;
#endif
}
MTFB_HD double div_by(double a, double b, double rb){
	double q = a * rb;
	double r = ::fma(-b, q, a);
	return ::fma(r, rb, q);
}

// getPixVal<Linear, Constant>: imgUtils.h:91-113.  overflow_val = 128 (imgUtils.h:57).
MTFB_HD double sample_pixel(const Image &im, double x, double y){
	if(check_overflow(x, y, im.hd, im.wd)){ return 128.0; }
	int lx = static_cast<int>(x);
	int ly = static_cast<int>(y);
	double dx = x - lx;
	double dy = y - ly;
	int ux = dx == 0 ? lx : lx + 1;
	int uy = dy == 0 ? ly : ly + 1;
	if(ux >= im.w || uy >= im.h){ return 128.0; }
	const float *r0 = im.data + (size_t)ly*im.pitch, *r1 = im.data + (size_t)uy*im.pitch;
	double p00 = MTFB_LDG(r0 + lx), p01 = MTFB_LDG(r0 + ux), p10 = MTFB_LDG(r1 + lx), p11 = MTFB_LDG(r1 + ux);
	return p00 * (1 - dx)*(1 - dy) + p01 * dx*(1 - dy) + p10 * (1 - dx)*dy + p11 * dx*dy;
}

// Pixel value (bit-exact getPixVal) and image gradient with the semantics of utils::getImgGrad
// (imgUtils.cc:233-254): the central difference  (P(x+eps, y) - P(x-eps, y)) * (mult / (2 eps))  of the
// piecewise-bilinear interpolant P (constant 128 outside the image), eps = grad_eps = 1e-8.
//   * FAST PATH -- x-eps and x+eps lie in the same pixel cell as x (all but a 2e-8-wide band around each
//     pixel column): P is linear in x there, so the quotient equals the cell's slope
//     (1-dy)(p01-p00) + dy(p11-p10) up to the quotient's own rounding noise (~1e-5 relative); the slope is
//     returned, computed from the four pixels already loaded for the value.
//   * LITERAL PATH -- the two samples straddle a pixel column (always the case at integer coordinates, e.g. an
//     axis-aligned patch at initialisation, where the quotient becomes the mean of the two one-sided slopes),
//     or one of them leaves the image (the reference then divides (128 - I) by 2e-8): the reference's two
//     getPixVal calls and its quotient are evaluated literally, so these cases match bit for bit.
// grad_mult = mult / (2 eps) as imgUtils.cc:238 forms it; mult = pix_norm_mult (1 except MI).
// The common path (sample_fast) is straight-line code -- validity is a predicate and the indices are clamped so
// that the four loads are always legal -- so the compiler can overlap the load latency with arithmetic that
// does not depend on the pixel values, and two pixels can be interleaved.  The literal path (sample_literal) is
// a rarely taken branch the caller runs afterwards for the components flagged in Sample::lit.
struct Sample { double val, gx, gy; int lit; };     // lit bit 0 / 1: gx / gy must be recomputed literally

// where the four neighbours of a sample come from: global memory (read-only path) ...
struct GlobalFetch {
	MTFB_HD void operator()(const Image &im, int lx, int ux, int ly, int uy, double &p00, double &p01, double &p10, double &p11) const{
		const float *r0 = im.data + (size_t)ly*im.pitch, *r1 = im.data + (size_t)uy*im.pitch;
		p00 = MTFB_LDG(r0 + lx); p01 = MTFB_LDG(r0 + ux); p10 = MTFB_LDG(r1 + lx); p11 = MTFB_LDG(r1 + ux);
	}
};
template<bool UNIT_MULT, class Fetch> MTFB_HD Sample sample_fast(const Image &im, const Fetch &fetch, double x, double y,
	double grad_eps, double mult){
	Sample o;
	const bool inb = !check_overflow(x, y, im.hd, im.wd);
	const double xs = inb ? x : 0.0, ys = inb ? y : 0.0;
	int lx = static_cast<int>(xs);
	int ly = static_cast<int>(ys);
	double dx = xs - lx;
	double dy = ys - ly;
	int ux = dx == 0 ? lx : lx + 1;
	int uy = dy == 0 ? ly : ly + 1;
	const bool ok = inb && ux < im.w && uy < im.h;
	ux = ux < im.w ? ux : im.w - 1;
	uy = uy < im.h ? uy : im.h - 1;
	double p00, p01, p10, p11;
	fetch(im, lx, ux, ly, uy, p00, p01, p10, p11);
	const double v = p00 * (1 - dx)*(1 - dy) + p01 * dx*(1 - dy) + p10 * (1 - dx)*dy + p11 * dx*dy;
	o.val = ok ? v : 128.0;
	// both +-eps samples in the cell of (x, y)?  (dx = x - lx is exact)
	const bool fast_x = ok && (dx >= grad_eps) && (dx <= 1 - grad_eps);
	const bool fast_y = ok && (dy >= grad_eps) && (dy <= 1 - grad_eps);
	o.gx = (1 - dy)*(p01 - p00) + dy*(p11 - p10);
	o.gy = (1 - dx)*(p10 - p00) + dx*(p11 - p01);
	if(!UNIT_MULT){ o.gx = o.gx * mult; o.gy = o.gy * mult; }
	o.lit = (fast_x ? 0 : 1) | (fast_y ? 0 : 2);
	return o;
}
template<bool UNIT_MULT> MTFB_HD Sample sample_fast(const Image &im, double x, double y, double grad_eps, double mult){
	return sample_fast<UNIT_MULT>(im, GlobalFetch(), x, y, grad_eps, mult);
}
MTFB_HD void sample_literal(const Image &im, double x, double y, double grad_eps, double grad_mult, Sample &o){
	if(o.lit & 1){
		double inc = sample_pixel(im, x + grad_eps, y), dec = sample_pixel(im, x - grad_eps, y);
		o.gx = (inc - dec)*grad_mult;
	}
	if(o.lit & 2){
		double inc = sample_pixel(im, x, y + grad_eps), dec = sample_pixel(im, x, y - grad_eps);
		o.gy = (inc - dec)*grad_mult;
	}
}
template<bool UNIT_MULT> MTFB_HD void sample_pixel_grad(const Image &im, double x, double y, double grad_eps,
	double grad_mult, double mult, double &val, double &gx, double &gy){
	Sample o = sample_fast<UNIT_MULT>(im, x, y, grad_eps, mult);
	if(o.lit) sample_literal(im, x, y, grad_eps, grad_mult, o);
	val = o.val; gx = o.gx; gy = o.gy;
}

// ------------------------------------------------------------------------------------------------
// 3x3 helpers, row-major m[3*r + c]; arithmetic order = Eigen's lazy coefficient-wise product
// ------------------------------------------------------------------------------------------------
struct Mat3 {
	double m[9];
	MTFB_HD double operator[](int i) const{ return m[i]; }
};
// 3x3 matrix read from memory at every use (volatile: keeps a warp-uniform matrix OUT of the register file;
// each read is one broadcast shared-memory load)
struct MemMat3 {
	const volatile double *p;
	MTFB_HD double operator[](int i) const{ return p[i]; }
};

MTFB_HD Mat3 mat3_identity(){
	Mat3 I;
#pragma unroll
	for(int i = 0; i < 9; ++i) I.m[i] = 0;
	I.m[0] = I.m[4] = I.m[8] = 1;
	return I;
}
MTFB_HD Mat3 mat3_mul(const Mat3 &a, const Mat3 &b){
	Mat3 c;
#pragma unroll
	for(int i = 0; i < 3; ++i)
#pragma unroll
	for(int j = 0; j < 3; ++j){
		double s = a.m[3 * i] * b.m[j];
		s = s + a.m[3 * i + 1] * b.m[3 + j];
		s = s + a.m[3 * i + 2] * b.m[6 + j];
		c.m[3 * i + j] = s;
	}
	return c;
}
// Matrix3d::inverse(): cofactors times 1/det (Eigen compute_inverse_size3_helper)
#define MTFB_COF(m, i, j) ((m).m[3 * (((i) + 1) % 3) + (((j) + 1) % 3)] * (m).m[3 * (((i) + 2) % 3) + (((j) + 2) % 3)] - \
	(m).m[3 * (((i) + 1) % 3) + (((j) + 2) % 3)] * (m).m[3 * (((i) + 2) % 3) + (((j) + 1) % 3)])
MTFB_HD Mat3 mat3_inverse(const Mat3 &m){
	double c0 = MTFB_COF(m, 0, 0), c1 = MTFB_COF(m, 1, 0), c2 = MTFB_COF(m, 2, 0);
	double det = c0 * m.m[0];
	det = det + c1 * m.m[3];
	det = det + c2 * m.m[6];
	double invdet = 1.0 / det;
	Mat3 r;
	r.m[0] = c0 * invdet; r.m[1] = c1 * invdet; r.m[2] = c2 * invdet;
	r.m[3] = MTFB_COF(m, 0, 1)*invdet; r.m[4] = MTFB_COF(m, 1, 1)*invdet; r.m[5] = MTFB_COF(m, 2, 1)*invdet;
	r.m[6] = MTFB_COF(m, 0, 2)*invdet; r.m[7] = MTFB_COF(m, 1, 2)*invdet; r.m[8] = MTFB_COF(m, 2, 2)*invdet;
	return r;
}

// Homography::getWarpFromState Homography.cc:94-107 ; Affine::getWarpFromState Affine.cc:117-131
template<int SSM> MTFB_HD Mat3 warp_from_state(const double *s){
	Mat3 w;
	if(SSM == SSM_HOM){
		w.m[0] = 1 + s[0]; w.m[1] = s[1]; w.m[2] = s[2];
		w.m[3] = s[3]; w.m[4] = 1 + s[4]; w.m[5] = s[5];
		w.m[6] = s[6]; w.m[7] = s[7]; w.m[8] = 1;
	} else if(SSM == SSM_TRANS){
		// Translation::getWarpFromState Translation.cc:93-101
		w.m[0] = 1; w.m[1] = 0; w.m[2] = s[0];
		w.m[3] = 0; w.m[4] = 1; w.m[5] = s[1];
		w.m[6] = 0; w.m[7] = 0; w.m[8] = 1;
	} else{
		w.m[0] = 1 + s[2]; w.m[1] = s[3]; w.m[2] = s[0];
		w.m[3] = s[4]; w.m[4] = 1 + s[5]; w.m[5] = s[1];
		w.m[6] = 0; w.m[7] = 0; w.m[8] = 1;
	}
	return w;
}
// Homography::getStateFromWarp Homography.cc:116-132 ; Affine::getStateFromWarp Affine.cc:133-143
template<int SSM> MTFB_HD void state_from_warp(double *s, const Mat3 &w){
	if(SSM == SSM_HOM){
		s[0] = w.m[0] - 1; s[1] = w.m[1]; s[2] = w.m[2];
		s[3] = w.m[3]; s[4] = w.m[4] - 1; s[5] = w.m[5];
		s[6] = w.m[6]; s[7] = w.m[7];
	} else if(SSM == SSM_TRANS){
		s[0] = w.m[2]; s[1] = w.m[5];                              // Translation.cc:103-109
	} else{
		s[0] = w.m[2]; s[1] = w.m[5]; s[2] = w.m[0] - 1;
		s[3] = w.m[1]; s[4] = w.m[3]; s[5] = w.m[4] - 1;
	}
}
// Homography::compositionalUpdate Homography.cc:73-92 ; Affine::compositionalUpdate Affine.cc:90-107
// (the re-warp of the points is done by the pixel loop)
template<int SSM> MTFB_HD Mat3 compose_update(const Mat3 &curr_warp, const double *state_update){
	Mat3 upd = warp_from_state<SSM>(state_update);
	Mat3 w = mat3_mul(curr_warp, upd);
	if(SSM == SSM_HOM){
		const double d = w.m[8], rd = ieee_rcp(d);
#pragma unroll
		for(int i = 0; i < 9; ++i) w.m[i] = div_by(w.m[i], d, rd);
	}
	return w;
}
// Homography::invertState Homography.cc:109-114 ; Affine::invertState Affine.cc:145-150
template<int SSM> MTFB_HD void invert_state(double *inv_state, const double *state){
	Mat3 w = warp_from_state<SSM>(state);
	Mat3 inv = mat3_inverse(w);
	const double d = inv.m[8], rd = ieee_rcp(d);
#pragma unroll
	for(int i = 0; i < 9; ++i) inv.m[i] = div_by(inv.m[i], d, rd);
	state_from_warp<SSM>(inv_state, inv);
}
// corners of the current region: W . init_corners_hm, dehomogenised for the homography
// (Homography.cc:85-91, Affine.cc:103-105).  init_corners: 8 doubles x0..x3,y0..y3 (hm z = 1)
template<int SSM> MTFB_HD void warp_corners(const Mat3 &w, const double *init_corners, double *out){
#pragma unroll
	for(int i = 0; i < 4; ++i){
		double px = init_corners[i], py = init_corners[4 + i];
		double hx = w.m[0] * px; hx = hx + w.m[1] * py; hx = hx + w.m[2] * 1.0;
		double hy = w.m[3] * px; hy = hy + w.m[4] * py; hy = hy + w.m[5] * 1.0;
		if(SSM == SSM_HOM){
			double hz = w.m[6] * px; hz = hz + w.m[7] * py; hz = hz + w.m[8] * 1.0;
			const double rhz = ieee_rcp(hz);
			out[i] = div_by(hx, hz, rhz); out[4 + i] = div_by(hy, hz, rhz);
		} else{
			out[i] = hx; out[4 + i] = hy;
		}
	}
}

// Geometry of one grid point under the current warp.
//   (u, v)         the normalised grid point (warpUtils.cc:15-33)
//   dlt            4-point DLT warp of ssm.setCorners (ProjectiveBase.cc:20-25)
//   ix, iy         init_pts      = dehomogenize(dlt . (u,v,1))
//   Homography (normalized_init = 0): init_pts_hm keeps the DLT's third row (Homography.cc:68), so
//                  curr_pts_hm = W . (dlt . (u,v,1)) and D = curr_pts_hm(2)
//   Affine:        init_pts_hm is re-homogenised (Affine.cc:81-82), curr_pts = W.topRows(2) . (ix,iy,1)
struct PixGeom { double ix, iy, wx, wy, cx, cy, D, rD; };   // (cx, cy, D) = curr_pts_hm; rD = RN(1 / D) = the reference's inv_det (Homography.cc:250)

// the iteration-invariant half: init_pts_hm = dlt . (u, v, 1) and init_pts = its dehomogenisation
template<class MD> MTFB_HD void template_point(const MD &dlt, double u, double v, double &hx, double &hy, double &hz,
	double &ix, double &iy){
	hx = dlt[0] * u; hx = hx + dlt[1] * v; hx = hx + dlt[2] * 1.0;
	hy = dlt[3] * u; hy = hy + dlt[4] * v; hy = hy + dlt[5] * 1.0;
	hz = dlt[6] * u; hz = hz + dlt[7] * v; hz = hz + dlt[8] * 1.0;
	const double rhz = ieee_rcp(hz);
	ix = div_by(hx, hz, rhz); iy = div_by(hy, hz, rhz);
}
// the per-pass half: the warped point (and its homogeneous denominator)
template<int SSM, class MW> MTFB_HD PixGeom warp_template_point(const MW &W, double hx, double hy, double hz, double ix, double iy){
	PixGeom g;
	g.ix = ix; g.iy = iy;
	if(SSM == SSM_HOM){
		double cx = W[0] * hx; cx = cx + W[1] * hy; cx = cx + W[2] * hz;
		double cy = W[3] * hx; cy = cy + W[4] * hy; cy = cy + W[5] * hz;
		double cz = W[6] * hx; cz = cz + W[7] * hy; cz = cz + W[8] * hz;
		g.cx = cx; g.cy = cy;
		g.D = cz; g.rD = ieee_rcp(cz); g.wx = div_by(cx, cz, g.rD); g.wy = div_by(cy, cz, g.rD);
	} else{
		double cx = W[0] * ix; cx = cx + W[1] * iy; cx = cx + W[2] * 1.0;
		double cy = W[3] * ix; cy = cy + W[4] * iy; cy = cy + W[5] * 1.0;
		g.cx = cx; g.cy = cy; g.D = 1.0; g.rD = 1.0; g.wx = cx; g.wy = cy;
	}
	return g;
}
// norm_init: hom_normalized_init = 1 (Homography.cc:57-62, shipped in Config/modules.cfg): the template points ARE the
// normalised grid (init_pts_hm = (u, v, 1)) and the DLT warp is part of curr_warp instead
template<int SSM, class MD, class MW> MTFB_HD PixGeom pixel_geometry(const MD &dlt, const MW &W, double u, double v, bool norm_init = false){
	double hx, hy, hz, ix, iy;
	if(norm_init){ hx = u; hy = v; hz = 1.0; ix = u; iy = v; }
	else template_point(dlt, u, v, hx, hy, hz, ix, iy);
	return warp_template_point<SSM>(W, hx, hy, hz, ix, iy);
}

// ssm.cmptWarpedPixJacobian: Homography.cc:231-294, Affine.cc:213-242.  gx, gy = dI/dx at the warped point.
// aff_abcd = (curr_state[2]+1, curr_state[3], curr_state[4], curr_state[5]+1) for the affine SSM.
template<int SSM, class MW> MTFB_HD void warped_pix_jacobian(const MW &W, const double *aff_abcd, const PixGeom &g,
	double gx, double gy, double *J){
	double x = g.ix, y = g.iy;
	if(SSM == SSM_HOM){
		double inv_det = g.rD;
		const double a00 = W[0], a01 = W[1], a10 = W[3], a11 = W[4], a20 = W[6], a21 = W[7];
		double dwx_dx = (a00 - a20 * g.wx), dwx_dy = (a01 - a21 * g.wx);
		double dwy_dx = (a10 - a20 * g.wy), dwy_dy = (a11 - a21 * g.wy);
		double Ix = (dwx_dx*gx + dwy_dx*gy)*inv_det;
		double Iy = (dwx_dy*gx + dwy_dy*gy)*inv_det;
		double Ixx = Ix*x, Ixy = Ix*y, Iyy = Iy*y, Iyx = Iy*x;
		J[0] = Ixx; J[1] = Ixy; J[2] = Ix; J[3] = Iyx; J[4] = Iyy; J[5] = Iy;
		J[6] = -x*Ixx - y*Iyx;
		J[7] = -x*Ixy - y*Iyy;
	} else if(SSM == SSM_TRANS){
		J[0] = gx; J[1] = gy;                                      // Translation.h:51-54: the pixel Jacobian is the gradient
	} else{
		double a = aff_abcd[0], b = aff_abcd[1], c = aff_abcd[2], d = aff_abcd[3];
		double Ix = gx, Iy = gy;
		double Ixx = Ix*x, Ixy = Ix*y, Iyy = Iy*y, Iyx = Iy*x;
		J[0] = Ix*a + Iy*c; J[1] = Ix*b + Iy*d;
		J[2] = Ixx*a + Iyx*c; J[3] = Ixy*a + Iyy*c;
		J[4] = Ixx*b + Iyx*d; J[5] = Ixy*b + Iyy*d;
	}
}
// The NON-chained gradient ({esm,fc,ic}_chained_warp = 0, the factory default of parameters.h:174,192): the gradient of
// the WARPED image with respect to the template coordinates, i.e. central differences over the four points
// ssm.updateGradPts builds (Homography.cc:803-827: curr_pts_hm +- eps * curr_warp.col(0|1), dehomogenised;
// Affine.cc:293-312: curr_pts +- eps * the 2x2 block's columns) fed to utils::getWarpedImgGrad (imgUtils.cc:177-202).
// Evaluated literally -- four more getPixVal's per pixel -- so it matches the reference bit for bit.
template<int SSM, class MW> MTFB_HD void warped_image_gradient(const Image &im, const MW &W, const PixGeom &g, double grad_eps,
	double grad_mult, double &gx, double &gy){
	double px[4], py[4];
	if(SSM == SSM_HOM){
		const double dx0 = W[0] * grad_eps, dx1 = W[3] * grad_eps, dx2 = W[6] * grad_eps;
		const double dy0 = W[1] * grad_eps, dy1 = W[4] * grad_eps, dy2 = W[7] * grad_eps;
		const double ax[4] = { g.cx + dx0, g.cx - dx0, g.cx + dy0, g.cx - dy0 };
		const double ay[4] = { g.cy + dx1, g.cy - dx1, g.cy + dy1, g.cy - dy1 };
		const double az[4] = { g.D + dx2, g.D - dx2, g.D + dy2, g.D - dy2 };
		for(int k = 0; k < 4; ++k){
			const double r = ieee_rcp(az[k]);
			px[k] = div_by(ax[k], az[k], r); py[k] = div_by(ay[k], az[k], r);
		}
	} else{
		const double dx0 = W[0] * grad_eps, dx1 = W[3] * grad_eps, dy0 = W[1] * grad_eps, dy1 = W[4] * grad_eps;
		px[0] = g.wx + dx0; py[0] = g.wy + dx1; px[1] = g.wx - dx0; py[1] = g.wy - dx1;
		px[2] = g.wx + dy0; py[2] = g.wy + dy1; px[3] = g.wx - dy0; py[3] = g.wy - dy1;
	}
	gx = (sample_pixel(im, px[0], py[0]) - sample_pixel(im, px[1], py[1]))*grad_mult;
	gy = (sample_pixel(im, px[2], py[2]) - sample_pixel(im, px[3], py[3]))*grad_mult;
}

// pixel Jacobian of the TEMPLATE from its stored, pre-chained gradient (Ix, Iy): what
// cmptWarpedPixJacobian produced at initialize() time, when curr_warp was the identity.
template<int SSM> MTFB_HD void init_pix_jacobian(double x, double y, double Ix, double Iy, double *J){
	double Ixx = Ix*x, Ixy = Ix*y, Iyy = Iy*y, Iyx = Iy*x;
	if(SSM == SSM_HOM){
		J[0] = Ixx; J[1] = Ixy; J[2] = Ix; J[3] = Iyx; J[4] = Iyy; J[5] = Iy;
		J[6] = -x*Ixx - y*Iyx;
		J[7] = -x*Ixy - y*Iyy;
	} else if(SSM == SSM_TRANS){
		J[0] = Ix; J[1] = Iy;                                      // Translation.h:45-49
	} else{
		J[0] = Ix; J[1] = Iy; J[2] = Ixx; J[3] = Ixy; J[4] = Iyx; J[5] = Iyy;
	}
}

// ssm.cmptPixJacobian (the forward ADDITIVE search, nt::FALK): Homography.cc:193-229; Affine.h:35-37 = cmptInitPixJacobian.
// gx, gy = dI/dx at the warped point.
template<int SSM> MTFB_HD void additive_pix_jacobian(const PixGeom &g, double gx, double gy, double *J){
	const double x = g.ix, y = g.iy;
	if(SSM == SSM_HOM){
		const double inv_d = g.rD;                                  // 1.0 / curr_pts_hm(2, pt_id)
		const double Ix = gx * inv_d, Iy = gy * inv_d;
		const double Ixx = Ix*x, Iyy = Iy*y, Ixy = Ix*y, Iyx = Iy*x;
		J[0] = Ixx; J[1] = Ixy; J[2] = Ix; J[3] = Iyx; J[4] = Iyy; J[5] = Iy;
		J[6] = (-g.wx*Ixx - g.wy*Iyx);
		J[7] = (-g.wx*Ixy - g.wy*Iyy);
	} else if(SSM == SSM_TRANS){
		J[0] = gx; J[1] = gy;                                      // Translation.h:60-63
	} else{
		const double Ixx = gx*x, Ixy = gx*y, Iyy = gy*y, Iyx = gy*x;
		J[0] = gx; J[1] = gy; J[2] = Ixx; J[3] = Ixy; J[4] = Iyx; J[5] = Iyy;
	}
}
// ssm.cmptApproxPixJacobian (the inverse ADDITIVE search, nt::IALK): Homography.cc:296-358, Affine.cc:183-211.
// g0x, g0y = the TEMPLATE's gradient (am.getInitPixGrad()); aff_abcd as for warped_pix_jacobian.
template<int SSM, class MW> MTFB_HD void approx_pix_jacobian(const MW &W, const double *aff_abcd, const PixGeom &g,
	double g0x, double g0y, double *J){
	const double x = g.ix, y = g.iy;
	if(SSM == SSM_HOM){
		const double a = (W[0] - W[6] * g.wx), b = (W[1] - W[7] * g.wx);
		const double c = (W[3] - W[6] * g.wy), d = (W[4] - W[7] * g.wy);
		const double inv_factor = ieee_rcp(a*d - b*c);
		const double Ix = (d*g0x - c*g0y)*inv_factor;
		const double Iy = (a*g0y - b*g0x)*inv_factor;
		const double Ixx = Ix*x, Ixy = Ix*y, Iyy = Iy*y, Iyx = Iy*x;
		J[0] = Ixx; J[1] = Ixy; J[2] = Ix; J[3] = Iyx; J[4] = Iyy; J[5] = Iy;
		J[6] = (-g.wx*Ixx - g.wy*Iyx);
		J[7] = (-g.wx*Ixy - g.wy*Iyy);
	} else if(SSM == SSM_TRANS){
		J[0] = g0x; J[1] = g0y;                                    // Translation.h:55-58
	} else{
		const double a = aff_abcd[0], b = aff_abcd[1], c = aff_abcd[2], d = aff_abcd[3];
		const double inv_det = ieee_rcp(a*d - b*c);
		const double Ix = g0x, Iy = g0y;
		const double Ixx = Ix*x, Ixy = Ix*y, Iyy = Iy*y, Iyx = Iy*x;
		J[0] = (Ix*d - Iy*c) * inv_det; J[1] = (Iy*a - Ix*b) * inv_det;
		J[2] = (Ixx*d - Iyx*c) * inv_det; J[3] = (Ixy*d - Iyy*c) * inv_det;
		J[4] = (Iyx*a - Ixx*b) * inv_det; J[5] = (Iyy*a - Ixy*b) * inv_det;
	}
}

template<int SSM> struct StateSize { static const int value = (SSM == SSM_HOM) ? 8 : (SSM == SSM_AFF ? 6 : 2); };

} // namespace mtfb
