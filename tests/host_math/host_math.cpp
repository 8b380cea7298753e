// TEST INFRASTRUCTURE ONLY.  Compiles mtf_b200/csrc/lk_math.cuh (the per-pixel arithmetic the CUDA kernels
// inline) for the HOST with g++ -ffp-contract=off, so that the CPU test suite can compare it with the oracle
// without a GPU.  Nothing in the product links or loads this file; the product has no CPU path.
#include <cstddef>
#include "../../mtf_b200/csrc/lk_math.cuh"
using namespace mtfb;

extern "C" {

// pts: n x 2 ; out: val[n], gx[n], gy[n]
void hm_sample_grad(const float *img, int h, int w, int pitch, const double *pts, int n, double grad_eps,
	double *val, double *gx, double *gy){
	Image im = make_image(img, h, w, pitch);
	for(int i = 0; i < n; ++i) sample_pixel_grad<true>(im, pts[2 * i], pts[2 * i + 1], grad_eps, 1.0 / (2 * grad_eps), 1.0, val[i], gx[i], gy[i]);
}
void hm_sample(const float *img, int h, int w, int pitch, const double *pts, int n, double *val){
	Image im = make_image(img, h, w, pitch);
	for(int i = 0; i < n; ++i) val[i] = sample_pixel(im, pts[2 * i], pts[2 * i + 1]);
}
// one full per-pixel stage for a patch: grid (xv, yv), dlt, W -> init pts, pts, It, grad, pixel Jacobian
void hm_stage(int ssm, const float *img, int h, int w, int pitch, const double *xv, int resx, const double *yv, int resy,
	const double *dlt9, const double *W9, double grad_eps,
	double *init_pts /* N x 2 */, double *pts /* N x 2 */, double *It, double *grad /* 2 x N */, double *jac /* S x N */){
	Image im = make_image(img, h, w, pitch);
	Mat3 dlt, W;
	for(int i = 0; i < 9; ++i){ dlt.m[i] = dlt9[i]; W.m[i] = W9[i]; }
	double abcd[4] = { (W.m[0] - 1) + 1, W.m[1], W.m[3], (W.m[4] - 1) + 1 };
	const int N = resx*resy;
	for(int r = 0; r < resy; ++r) for(int c = 0; c < resx; ++c){
		int i = r*resx + c;
		double val, gx, gy, J[8];
		if(ssm == SSM_HOM){
			PixGeom g = pixel_geometry<SSM_HOM>(dlt, W, xv[c], yv[r]);
			sample_pixel_grad<true>(im, g.wx, g.wy, grad_eps, 1.0 / (2 * grad_eps), 1.0, val, gx, gy);
			warped_pix_jacobian<SSM_HOM>(W, abcd, g, gx, gy, J);
			init_pts[2 * i] = g.ix; init_pts[2 * i + 1] = g.iy; pts[2 * i] = g.wx; pts[2 * i + 1] = g.wy;
			for(int s = 0; s < 8; ++s) jac[(size_t)s*N + i] = J[s];
		} else{
			PixGeom g = pixel_geometry<SSM_AFF>(dlt, W, xv[c], yv[r]);
			sample_pixel_grad<true>(im, g.wx, g.wy, grad_eps, 1.0 / (2 * grad_eps), 1.0, val, gx, gy);
			warped_pix_jacobian<SSM_AFF>(W, abcd, g, gx, gy, J);
			init_pts[2 * i] = g.ix; init_pts[2 * i + 1] = g.iy; pts[2 * i] = g.wx; pts[2 * i + 1] = g.wy;
			for(int s = 0; s < 6; ++s) jac[(size_t)s*N + i] = J[s];
		}
		It[i] = val; grad[i] = gx; grad[N + i] = gy;
	}
}
void hm_compose(int ssm, const double *W9, const double *dp, double *out9){
	Mat3 W; for(int i = 0; i < 9; ++i) W.m[i] = W9[i];
	Mat3 r = ssm == SSM_HOM ? compose_update<SSM_HOM>(W, dp) : compose_update<SSM_AFF>(W, dp);
	for(int i = 0; i < 9; ++i) out9[i] = r.m[i];
}
void hm_invert_state(int ssm, const double *s, double *inv){
	if(ssm == SSM_HOM) invert_state<SSM_HOM>(inv, s); else invert_state<SSM_AFF>(inv, s);
}
void hm_warp_corners(int ssm, const double *W9, const double *init_corners, double *out8){
	Mat3 W; for(int i = 0; i < 9; ++i) W.m[i] = W9[i];
	if(ssm == SSM_HOM) warp_corners<SSM_HOM>(W, init_corners, out8); else warp_corners<SSM_AFF>(W, init_corners, out8);
}

}
