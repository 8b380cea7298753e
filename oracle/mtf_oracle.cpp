/*
 * mtf_oracle.cpp -- CPU oracle: fp64 restatement of MTF's Lucas-Kanade hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see mtf_oracle.h).  PARITY UNPINNED: the reference has
 * no golden vectors for this path and cannot be built here; every function cites
 * the reference file:line it follows (paths relative to /root/reference).
 *
 * Third-party arithmetic restated (not vendored by the reference, version unpinned;
 * ReadMe.md:111 asks for Eigen >= 3.2.10, we follow the published Eigen 3.3 algorithms):
 *   - ColPivHouseholderQR::compute/_solve_impl  (LAPACK dgeqp3-style norm down-dating)
 *   - JacobiSVD<8x9, FullV>: only matrixV().col(8) is consumed; for cols > rows Eigen's
 *     QR preconditioner makes that column the last column of Q of the (col-pivoted)
 *     Householder QR of A^T, which is what dlt_null_vector() computes
 *   - Matrix3d::inverse (cofactor formula), DenseBase::LinSpaced (3.3 semantics)
 *   - fixed/dynamic matrix products: k-ordered mul+add, no FMA (reference builds
 *     with plain -O3, makefile:104); this file must be compiled with -ffp-contract=off.
 */
#include "mtf_oracle.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <cfloat>
#include <vector>
#include <algorithm>
#include <limits>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef std::vector<double> vec;

static inline double now_s(){
	struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// ---------------------------------------------------------------------------------------------
// Utilities/include/mtf/Utilities/imgUtils.h:51-113
// ---------------------------------------------------------------------------------------------
static inline bool checkOverflow(double x, double y, unsigned int h, unsigned int w){
	// imgUtils.h:51-53 is ((x < 0) || (x >= w) || (y < 0) || (y >= h)).  Written here so that a NaN coordinate counts as
	// outside: the reference then goes on to static_cast<int>(NaN), undefined behaviour that on x86 yields INT_MIN and is
	// caught by its second checkOverflow(lx, ly) -- unless the optimiser, entitled to assume lx >= 0 after the first check,
	// removes that test (gcc 13 -O3 does: out-of-bounds read).  Same result for every non-NaN input.
	return !((x >= 0) && (x < w) && (y >= 0) && (y < h));
}
// getPixVal<InterpType::Linear, BorderType::Constant>                          imgUtils.h:91-113
static inline double getPixVal(const float *img, double x, double y,
	unsigned int h, unsigned int w, double overflow_val = 128.0){
	if(checkOverflow(x, y, h, w)){ return overflow_val; }
	int lx = static_cast<int>(x);
	int ly = static_cast<int>(y);
	double dx = x - lx;
	double dy = y - ly;
	int ux = dx == 0 ? lx : lx + 1;
	int uy = dy == 0 ? ly : ly + 1;
	if(checkOverflow(lx, ly, h, w) || checkOverflow(ux, uy, h, w)){ return overflow_val; }
	// EigImgT is row-major float (common.h:190); float * double promotes to double
	return img[(size_t)ly*w + lx] * (1 - dx)*(1 - dy) +
		img[(size_t)ly*w + ux] * dx*(1 - dy) +
		img[(size_t)uy*w + lx] * (1 - dx)*dy +
		img[(size_t)uy*w + ux] * dx*dy;
}
// utils::getPixVals                                                            imgUtils.cc:163-173
static void getPixVals(double *pix_vals, const float *img, const double *pts, unsigned int n_pix,
	unsigned int h, unsigned int w, double norm_mult, double norm_add){
	for(unsigned int i = 0; i < n_pix; i++){
		pix_vals[i] = norm_mult * getPixVal(img, pts[2 * i], pts[2 * i + 1], h, w) + norm_add;
	}
}
// utils::getImgGrad                                                            imgUtils.cc:233-254
static void getImgGrad(double *img_grad, const float *img, const double *pts, double grad_eps,
	unsigned int n_pix, unsigned int h, unsigned int w, double pix_mult_factor){
	double grad_mult_factor = pix_mult_factor / (2 * grad_eps);
	for(unsigned int pix_id = 0; pix_id < n_pix; ++pix_id){
		double curr_x = pts[2 * pix_id], curr_y = pts[2 * pix_id + 1];
		double pix_val_inc = getPixVal(img, curr_x + grad_eps, curr_y, h, w);
		double pix_val_dec = getPixVal(img, curr_x - grad_eps, curr_y, h, w);
		img_grad[pix_id] = (pix_val_inc - pix_val_dec)*grad_mult_factor;
		pix_val_inc = getPixVal(img, curr_x, curr_y + grad_eps, h, w);
		pix_val_dec = getPixVal(img, curr_x, curr_y - grad_eps, h, w);
		img_grad[n_pix + pix_id] = (pix_val_inc - pix_val_dec)*grad_mult_factor;
	}
}
// NOT a reference function (grad_mode = 1): getImgGrad above with the difference quotient replaced by the
// cell's slope wherever x - eps and x + eps fall into the same pixel cell as x.  There the interpolant is
// linear, so getImgGrad's quotient equals that slope up to its own rounding noise (~1e-5 relative); in every
// other case (the samples straddle a pixel column or row -- always so at integer coordinates -- or one of
// them leaves the image) the reference's quotient is evaluated literally.  The product computes its
// gradients this way, so tests compare it (i) against this mode tightly and (ii) against the reference
// mode within the quotient's noise.
static void getImgGradAnalytic(double *img_grad, const float *img, const double *pts, double grad_eps,
	unsigned int n_pix, unsigned int h, unsigned int w, double pix_mult_factor){
	double grad_mult_factor = pix_mult_factor / (2 * grad_eps);
	for(unsigned int pix_id = 0; pix_id < n_pix; ++pix_id){
		double x = pts[2 * pix_id], y = pts[2 * pix_id + 1];
		bool same_cell_x = false, same_cell_y = false;
		double slope_x = 0, slope_y = 0;
		if(!checkOverflow(x, y, h, w)){
			int lx = static_cast<int>(x), ly = static_cast<int>(y);
			double dx = x - lx, dy = y - ly;
			int ux = dx == 0 ? lx : lx + 1, uy = dy == 0 ? ly : ly + 1;
			if(ux < (int)w && uy < (int)h){
				double p00 = img[(size_t)ly*w + lx], p01 = img[(size_t)ly*w + ux];
				double p10 = img[(size_t)uy*w + lx], p11 = img[(size_t)uy*w + ux];
				same_cell_x = (dx >= grad_eps) && (dx <= 1 - grad_eps);       // dx = x - lx is exact
				same_cell_y = (dy >= grad_eps) && (dy <= 1 - grad_eps);
				slope_x = (1 - dy)*(p01 - p00) + dy*(p11 - p10);
				slope_y = (1 - dx)*(p10 - p00) + dx*(p11 - p01);
				if(pix_mult_factor != 1){ slope_x = slope_x * pix_mult_factor; slope_y = slope_y * pix_mult_factor; }
			}
		}
		if(same_cell_x){ img_grad[pix_id] = slope_x; }
		else{
			double pix_val_inc = getPixVal(img, x + grad_eps, y, h, w);
			double pix_val_dec = getPixVal(img, x - grad_eps, y, h, w);
			img_grad[pix_id] = (pix_val_inc - pix_val_dec)*grad_mult_factor;
		}
		if(same_cell_y){ img_grad[n_pix + pix_id] = slope_y; }
		else{
			double pix_val_inc = getPixVal(img, x, y + grad_eps, h, w);
			double pix_val_dec = getPixVal(img, x, y - grad_eps, h, w);
			img_grad[n_pix + pix_id] = (pix_val_inc - pix_val_dec)*grad_mult_factor;
		}
	}
}
// utils::getWarpedImgGrad                                                      imgUtils.cc:177-202
static void getWarpedImgGrad(double *warped_img_grad, const float *img, const double *wop /*8xN*/,
	double grad_eps, unsigned int n_pix, unsigned int h, unsigned int w, double pix_mult_factor){
	double grad_mult_factor = pix_mult_factor / (2 * grad_eps);
	for(unsigned int pix_id = 0; pix_id < n_pix; ++pix_id){
		const double *p = wop + 8 * (size_t)pix_id;
		double pix_val_inc = getPixVal(img, p[0], p[1], h, w);
		double pix_val_dec = getPixVal(img, p[2], p[3], h, w);
		warped_img_grad[pix_id] = (pix_val_inc - pix_val_dec)*grad_mult_factor;
		pix_val_inc = getPixVal(img, p[4], p[5], h, w);
		pix_val_dec = getPixVal(img, p[6], p[7], h, w);
		warped_img_grad[n_pix + pix_id] = (pix_val_inc - pix_val_dec)*grad_mult_factor;
	}
}

// ---------------------------------------------------------------------------------------------
// Eigen restatements
// ---------------------------------------------------------------------------------------------
// DenseBase::LinSpaced(size, low, high), Eigen 3.3 linspaced_op_impl<double,false>
static void linSpaced(double *out, int size, double low, double high){
	int size1 = size == 1 ? 1 : size - 1;
	double step = size == 1 ? 0.0 : (high - low) / double(size - 1);
	bool flip = std::fabs(high) < std::fabs(low);
	for(int i = 0; i < size; ++i){
		if(flip){ out[i] = (i == 0) ? low : (high - (size1 - i)*step); }
		else{ out[i] = (i == size1) ? high : (low + i*step); }
	}
}
// 3x3 row-major helpers (the reference's Matrix3d is column-major; only the arithmetic order matters)
struct Mat3{ double m[9]; double& operator()(int r, int c){ return m[3 * r + c]; }
	double operator()(int r, int c) const{ return m[3 * r + c]; } };
static Mat3 identity3(){ Mat3 I; for(int i = 0; i < 9; ++i) I.m[i] = 0; I(0, 0) = I(1, 1) = I(2, 2) = 1; return I; }
// fixed-size lazy product: coefficient-wise, k ascending
static Mat3 mul3(const Mat3 &a, const Mat3 &b){
	Mat3 c;
	for(int i = 0; i < 3; ++i) for(int j = 0; j < 3; ++j){
		double s = a(i, 0)*b(0, j);
		s = s + a(i, 1)*b(1, j);
		s = s + a(i, 2)*b(2, j);
		c(i, j) = s;
	}
	return c;
}
// Matrix3d::inverse(): Eigen compute_inverse_size3_helper (cofactors * 1/det)
static inline double cofactor3(const Mat3 &m, int i, int j){
	int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
	return m(i1, j1)*m(i2, j2) - m(i1, j2)*m(i2, j1);
}
static Mat3 inverse3(const Mat3 &m){
	double c0 = cofactor3(m, 0, 0), c1 = cofactor3(m, 1, 0), c2 = cofactor3(m, 2, 0);
	double det = c0*m(0, 0);
	det = det + c1*m(1, 0);
	det = det + c2*m(2, 0);
	double invdet = 1.0 / det;
	Mat3 r;
	r(0, 0) = c0*invdet; r(0, 1) = c1*invdet; r(0, 2) = c2*invdet;
	r(1, 0) = cofactor3(m, 0, 1)*invdet; r(1, 1) = cofactor3(m, 1, 1)*invdet; r(1, 2) = cofactor3(m, 2, 1)*invdet;
	r(2, 0) = cofactor3(m, 0, 2)*invdet; r(2, 1) = cofactor3(m, 1, 2)*invdet; r(2, 2) = cofactor3(m, 2, 2)*invdet;
	return r;
}

// Column-pivoted Householder QR on a column-major rows x cols matrix, Eigen 3.3
// ColPivHouseholderQR::computeInPlace.  Returns number of nonzero pivots.
struct ColPivQR{
	int rows, cols, size, nonzero_pivots;
	vec qr, hCoeffs, colNormsUpdated, colNormsDirect;
	std::vector<int> perm;       // perm[i] = original column at position i
	double& A(int r, int c){ return qr[(size_t)c*rows + r]; }
	void compute(const double *mat, int _rows, int _cols){
		rows = _rows; cols = _cols; size = std::min(rows, cols);
		qr.assign(mat, mat + (size_t)rows*cols);
		hCoeffs.assign(size, 0); colNormsUpdated.assign(cols, 0); colNormsDirect.assign(cols, 0);
		std::vector<int> transp(size);
		for(int k = 0; k < cols; ++k){
			double s = 0; for(int r = 0; r < rows; ++r) s += A(r, k)*A(r, k);
			colNormsDirect[k] = std::sqrt(s); colNormsUpdated[k] = colNormsDirect[k];
		}
		const double eps = DBL_EPSILON;
		double maxn = *std::max_element(colNormsUpdated.begin(), colNormsUpdated.end());
		// Eigen 3.3 ColPivHouseholderQR.h computeInPlace():
		//   RealScalar threshold_helper = numext::abs2<RealScalar>(m_colNormsUpdated.maxCoeff() * NumTraits<RealScalar>::epsilon()) / RealScalar(rows);
		// i.e. (max norm * eps)^2 / rows -- the division by rows is OUTSIDE the square (Eigen 3.2 has the same value:
		// m_colSqNorms.maxCoeff() * abs2(epsilon) / rows)
		double me = maxn*eps;
		double threshold_helper = me*me / double(rows);
		double norm_downdate_threshold = std::sqrt(eps);
		nonzero_pivots = size;
		vec temp(cols);
		for(int k = 0; k < size; ++k){
			int biggest = k; double bmax = colNormsUpdated[k];
			for(int j = k + 1; j < cols; ++j) if(colNormsUpdated[j] > bmax){ bmax = colNormsUpdated[j]; biggest = j; }
			double biggest_col_sq_norm = bmax*bmax;
			if(nonzero_pivots == size && biggest_col_sq_norm < threshold_helper * double(rows - k)) nonzero_pivots = k;
			transp[k] = biggest;
			if(k != biggest){
				for(int r = 0; r < rows; ++r) std::swap(A(r, k), A(r, biggest));
				std::swap(colNormsUpdated[k], colNormsUpdated[biggest]);
				std::swap(colNormsDirect[k], colNormsDirect[biggest]);
			}
			// makeHouseholderInPlace on col(k).tail(rows-k)
			double c0 = A(k, k), tailSqNorm = 0;
			for(int r = k + 1; r < rows; ++r) tailSqNorm += A(r, k)*A(r, k);
			double tau, beta;
			if(tailSqNorm <= std::numeric_limits<double>::min()){
				tau = 0; beta = c0;
				for(int r = k + 1; r < rows; ++r) A(r, k) = 0;
			} else{
				beta = std::sqrt(c0*c0 + tailSqNorm);
				if(c0 >= 0) beta = -beta;
				for(int r = k + 1; r < rows; ++r) A(r, k) = A(r, k) / (c0 - beta);
				tau = (beta - c0) / beta;
			}
			hCoeffs[k] = tau;
			A(k, k) = beta;
			// applyHouseholderOnTheLeft to bottomRightCorner(rows-k, cols-k-1)
			int brows = rows - k;
			for(int j = k + 1; j < cols; ++j) temp[j] = 0;
			if(brows == 1){
				for(int j = k + 1; j < cols; ++j) A(k, j) *= (1 - tau);
			} else if(tau != 0){
				for(int j = k + 1; j < cols; ++j){
					double t = 0;
					for(int r = k + 1; r < rows; ++r) t += A(r, k)*A(r, j);
					t += A(k, j);
					A(k, j) -= tau*t;
					for(int r = k + 1; r < rows; ++r) A(r, j) -= tau*A(r, k)*t;
				}
			}
			// norm down-dating
			for(int j = k + 1; j < cols; ++j){
				if(colNormsUpdated[j] != 0){
					double t = std::fabs(A(k, j)) / colNormsUpdated[j];
					t = (1 + t)*(1 - t);
					t = t < 0 ? 0 : t;
					double r = colNormsUpdated[j] / colNormsDirect[j];
					double t2 = t*(r*r);
					if(t2 <= norm_downdate_threshold){
						double s = 0; for(int rr = k + 1; rr < rows; ++rr) s += A(rr, j)*A(rr, j);
						colNormsDirect[j] = std::sqrt(s);
						colNormsUpdated[j] = colNormsDirect[j];
					} else{
						colNormsUpdated[j] *= std::sqrt(t);
					}
				}
			}
		}
		perm.resize(cols);
		for(int i = 0; i < cols; ++i) perm[i] = i;
		for(int k = 0; k < size; ++k) std::swap(perm[k], perm[transp[k]]);
	}
	// apply H_k (k-th reflector) to a length-rows vector
	void applyReflector(int k, double *c){
		double tau = hCoeffs[k];
		if(rows - k == 1){ c[k] *= (1 - tau); return; }
		if(tau == 0) return;
		double t = 0;
		for(int r = k + 1; r < rows; ++r) t += A(r, k)*c[r];
		t += c[k];
		c[k] -= tau*t;
		for(int r = k + 1; r < rows; ++r) c[r] -= tau*A(r, k)*t;
	}
	// ColPivHouseholderQR::_solve_impl (square or tall system, single rhs)
	void solve(const double *b, double *x){
		if(nonzero_pivots == 0){ for(int i = 0; i < cols; ++i) x[i] = 0; return; }
		vec c(b, b + rows);
		for(int k = 0; k < nonzero_pivots; ++k) applyReflector(k, c.data());   // Q^T b
		for(int i = nonzero_pivots - 1; i >= 0; --i){                          // upper-triangular solve
			double s = c[i];
			for(int j = i + 1; j < nonzero_pivots; ++j) s -= A(i, j)*c[j];
			c[i] = s / A(i, i);
		}
		for(int i = 0; i < nonzero_pivots; ++i) x[perm[i]] = c[i];
		for(int i = nonzero_pivots; i < cols; ++i) x[perm[i]] = 0;
	}
};

// utils::computeHomographyDLT(in_corners, out_corners)                         warpUtils.cc:171-223
// corners are 2x4 row-major here: c[0..3] = x, c[4..7] = y
static Mat3 computeHomographyDLT(const double *in_c, const double *out_c){
	double cm[8 * 9];                 // constraint_matrix, row-major 8x9
	for(int i = 0; i < 4; i++){
		double ix = in_c[i], iy = in_c[4 + i], ox = out_c[i], oy = out_c[4 + i];
		double *r1 = cm + 9 * (2 * i), *r2 = cm + 9 * (2 * i + 1);
		r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = -ix; r1[4] = -iy; r1[5] = -1;
		r1[6] = oy*ix; r1[7] = oy*iy; r1[8] = oy;
		r2[0] = ix; r2[1] = iy; r2[2] = 1; r2[3] = 0; r2[4] = 0; r2[5] = 0;
		r2[6] = -ox*ix; r2[7] = -ox*iy; r2[8] = -ox;
	}
	// JacobiSVD<Matrix89d>(ComputeFullV).matrixV().col(8): last column of Q of the
	// column-pivoted Householder QR of the adjoint (9x8).  cm row-major 8x9 == A^T col-major 9x8.
	// Eigen 3.3 JacobiSVD::compute first divides by scale = cwiseAbs().maxCoeff()
	double scale = 0; for(int i = 0; i < 72; ++i) scale = std::max(scale, std::fabs(cm[i]));
	if(scale == 0) scale = 1;
	for(int i = 0; i < 72; ++i) cm[i] = cm[i] / scale;
	ColPivQR qr; qr.compute(cm, 9, 8);
	double h[9]; for(int i = 0; i < 9; ++i) h[i] = 0; h[8] = 1;
	for(int k = 7; k >= 0; --k) qr.applyReflector(k, h);                       // Q e_9 = H_0 ... H_7 e_9
	Mat3 H;
	for(int i = 0; i < 9; ++i) H.m[i] = h[i] / h[8];                            // hom_mat /= h(8)
	return H;
}

// utils::getNormUnitSquarePts                                                  warpUtils.cc:15-33
static void getNormUnitSquarePts(double *basis_pts, double *basis_corners, int resx, int resy,
	double min_x = -0.5, double min_y = -0.5, double max_x = 0.5, double max_y = 0.5){
	vec x_vals(resx), y_vals(resy);
	linSpaced(x_vals.data(), resx, min_x, max_x);
	linSpaced(y_vals.data(), resy, min_y, max_y);
	int pt_id = 0;
	for(int row_id = 0; row_id < resy; row_id++){
		for(int col_id = 0; col_id < resx; col_id++){
			basis_pts[2 * pt_id] = x_vals[col_id];
			basis_pts[2 * pt_id + 1] = y_vals[row_id];
			++pt_id;
		}
	}
	basis_corners[0] = min_x; basis_corners[1] = max_x; basis_corners[2] = max_x; basis_corners[3] = min_x;
	basis_corners[4] = min_y; basis_corners[5] = min_y; basis_corners[6] = max_y; basis_corners[7] = max_y;
}

// utils::computeAffineNDLT(in_corners, out_corners)                           warpUtils.cc:378-386
//   = normalizePts(out_corners) (warpUtils.cc:802-832: centroid, mean distance, scale sqrt(2) / mean_dist)
//   + computeAffineDLT(in_corners, norm_corners) (warpUtils.cc:276-343: the least-squares solution of the 8 x 6 system
//     [x y 1 0 0 0; 0 0 0 x y 1] a = (ox, oy) through JacobiSVD's pseudo-inverse)
//   + inv_norm_mat * warp.
// The system is block diagonal with the same 4 x 3 block P = [x y 1] for both rows of the affine warp and has full column rank,
// so its pseudo-inverse solution is the pair of least-squares solutions of P a = ox, P a = oy; they are computed here by
// Householder QR (no SVD: Eigen's Jacobi sweeps are not restated -- the two agree to rounding, ~1e-15 relative, and
// tests/test_oracle.py pins this function against numpy.linalg.lstsq, which is SVD based like the reference).
static void lstsq_4x3(const double P[4][3], const double *b0, const double *b1, double *a0, double *a1){
	double A[4][3], c0[4], c1[4];
	for(int i = 0; i < 4; ++i){ for(int j = 0; j < 3; ++j) A[i][j] = P[i][j]; c0[i] = b0[i]; c1[i] = b1[i]; }
	for(int k = 0; k < 3; ++k){
		double nrm = 0;
		for(int i = k; i < 4; ++i) nrm += A[i][k] * A[i][k];
		nrm = std::sqrt(nrm);
		if(nrm == 0) continue;
		const double alpha = A[k][k] >= 0 ? -nrm : nrm;
		double v[4] = { 0, 0, 0, 0 };
		for(int i = k; i < 4; ++i) v[i] = A[i][k];
		v[k] -= alpha;
		double vv = 0;
		for(int i = k; i < 4; ++i) vv += v[i] * v[i];
		if(vv == 0) continue;
		for(int j = k; j < 3; ++j){
			double d = 0;
			for(int i = k; i < 4; ++i) d += v[i] * A[i][j];
			d = 2 * d / vv;
			for(int i = k; i < 4; ++i) A[i][j] -= d * v[i];
		}
		double d0 = 0, d1 = 0;
		for(int i = k; i < 4; ++i){ d0 += v[i] * c0[i]; d1 += v[i] * c1[i]; }
		d0 = 2 * d0 / vv; d1 = 2 * d1 / vv;
		for(int i = k; i < 4; ++i){ c0[i] -= d0 * v[i]; c1[i] -= d1 * v[i]; }
	}
	for(int k = 2; k >= 0; --k){
		double s0 = c0[k], s1 = c1[k];
		for(int j = k + 1; j < 3; ++j){ s0 -= A[k][j] * a0[j]; s1 -= A[k][j] * a1[j]; }
		a0[k] = s0 / A[k][k]; a1[k] = s1 / A[k][k];
	}
}
static Mat3 computeAffineNDLT(const double *in_c, const double *out_c){
	// normalizePts
	double cx = 0, cy = 0;
	for(int i = 0; i < 4; ++i){ cx += out_c[i]; cy += out_c[4 + i]; }
	cx /= 4; cy /= 4;
	double tx[4], ty[4], mean_dist = 0;
	for(int i = 0; i < 4; ++i){ tx[i] = out_c[i] - cx; ty[i] = out_c[4 + i] - cy; mean_dist += std::sqrt(tx[i] * tx[i] + ty[i] * ty[i]); }
	mean_dist /= 4;
	const double norm_scale = std::sqrt(2.0) / mean_dist;
	double nx[4], ny[4];
	for(int i = 0; i < 4; ++i){ nx[i] = tx[i] * norm_scale; ny[i] = ty[i] * norm_scale; }
	// computeAffineDLT
	double P[4][3], a0[3], a1[3];
	for(int i = 0; i < 4; ++i){ P[i][0] = in_c[i]; P[i][1] = in_c[4 + i]; P[i][2] = 1; }
	lstsq_4x3(P, nx, ny, a0, a1);
	Mat3 aff, inv_norm = identity3();
	aff(0, 0) = a0[0]; aff(0, 1) = a0[1]; aff(0, 2) = a0[2];
	aff(1, 0) = a1[0]; aff(1, 1) = a1[1]; aff(1, 2) = a1[2];
	aff(2, 0) = 0; aff(2, 1) = 0; aff(2, 2) = 1;
	inv_norm(0, 0) = 1.0 / norm_scale; inv_norm(1, 1) = 1.0 / norm_scale; inv_norm(0, 2) = cx; inv_norm(1, 2) = cy;
	return mul3(inv_norm, aff);
}

// ---------------------------------------------------------------------------------------------
// SSM: ProjectiveBase + Homography + Affine
// ---------------------------------------------------------------------------------------------
struct SSM{
	int type, resx, resy, n_pts, state_size;
	bool normalized_init;
	vec norm_pts, norm_pts_hm, init_pts, curr_pts, init_pts_hm, curr_pts_hm, grad_pts;
	double norm_corners[8], norm_corners_hm[12], init_corners[8], curr_corners[8];
	double init_corners_hm[12], curr_corners_hm[12];   // 3x4 column-major: hm[3*i + r]
	Mat3 curr_warp, dlt_warp;
	vec curr_state;

	// ProjectiveBase::ProjectiveBase SSM/src/ProjectiveBase.cc:9-18, Homography.cc:32-48, Affine.cc:37-62
	SSM(int _type, int _resx, int _resy, bool _norm_init) : type(_type), resx(_resx), resy(_resy),
		n_pts(_resx*_resy), normalized_init(_norm_init){
		state_size = type == ORC_SSM_HOMOGRAPHY ? 8 : (type == ORC_SSM_AFFINE ? 6 : 2);
		norm_pts.resize(2 * n_pts); norm_pts_hm.resize(3 * n_pts);
		init_pts.resize(2 * n_pts); curr_pts.resize(2 * n_pts);
		init_pts_hm.resize(3 * n_pts); curr_pts_hm.resize(3 * n_pts);
		curr_state.assign(state_size, 0);
		if(type == ORC_SSM_AFFINE){
			getNormUnitSquarePts(norm_pts.data(), norm_corners, resx, resy,
				1 - resx / 2.0, 1 - resy / 2.0, resx / 2.0, resy / 2.0);               // Affine.cc:53-54
		} else{
			getNormUnitSquarePts(norm_pts.data(), norm_corners, resx, resy);
		}
		homogenize(norm_pts.data(), norm_pts_hm.data(), n_pts);
		homogenize_corners(norm_corners, norm_corners_hm);
		curr_warp = identity3(); dlt_warp = identity3();
	}
	static void homogenize(const double *p, double *hm, int n){                    // warpUtils.h:9-14
		for(int i = 0; i < n; ++i){ hm[3 * i] = p[2 * i]; hm[3 * i + 1] = p[2 * i + 1]; hm[3 * i + 2] = 1; }
	}
	static void homogenize_corners(const double *c, double *hm){
		for(int i = 0; i < 4; ++i){ hm[3 * i] = c[i]; hm[3 * i + 1] = c[4 + i]; hm[3 * i + 2] = 1; }
	}
	// warp * hm (3x3 . 3xN, k-ascending mul+add) then utils::dehomogenize (warpUtils.h:16-21)
	static void warp_hm(const Mat3 &W, const double *in_hm, double *out_hm, int n){
		for(int i = 0; i < n; ++i){
			const double *p = in_hm + 3 * i;
			for(int r = 0; r < 3; ++r){
				double s = W(r, 0)*p[0];
				s = s + W(r, 1)*p[1];
				s = s + W(r, 2)*p[2];
				out_hm[3 * i + r] = s;
			}
		}
	}
	static void dehomogenize(const double *hm, double *p, int n){
		for(int i = 0; i < n; ++i){ p[2 * i] = hm[3 * i] / hm[3 * i + 2]; p[2 * i + 1] = hm[3 * i + 1] / hm[3 * i + 2]; }
	}
	static void dehomogenize_corners(const double *hm, double *c){
		for(int i = 0; i < 4; ++i){ c[i] = hm[3 * i] / hm[3 * i + 2]; c[4 + i] = hm[3 * i + 1] / hm[3 * i + 2]; }
	}
	// ProjectiveBase::getPtsFromCorners ProjectiveBase.cc:20-25
	void getPtsFromCorners(Mat3 &warp, double *pts, double *pts_hm, const double *corners){
		warp = computeHomographyDLT(norm_corners, corners);
		warp_hm(warp, norm_pts_hm.data(), pts_hm, n_pts);
		dehomogenize(pts_hm, pts, n_pts);
	}
	// Homography::setCorners Homography.cc:50-71 ; Affine::setCorners Affine.cc:64-88 (normalized_init=0)
	bool setCorners(const double *corners){
		if(type == ORC_SSM_AFFINE && normalized_init){
			// Affine::setCorners with normalized_init (Affine.cc:65-74): the template stays the pixel-scaled square, the least-squares
			// affine warp onto the corners becomes curr_warp, and the corners are what that warp makes of the square's
			curr_warp = computeAffineNDLT(norm_corners, corners);
			dlt_warp = curr_warp;
			std::memcpy(init_corners, norm_corners, sizeof(init_corners));
			std::memcpy(init_corners_hm, norm_corners_hm, sizeof(init_corners_hm));
			init_pts = norm_pts; init_pts_hm = norm_pts_hm;
			getStateFromWarp(curr_state.data(), curr_warp);
			affine_pts();
			homogenize(curr_pts.data(), curr_pts_hm.data(), n_pts);
			homogenize_corners(curr_corners, curr_corners_hm);
			return true;
		}
		std::memcpy(curr_corners, corners, sizeof(curr_corners));
		homogenize_corners(curr_corners, curr_corners_hm);
		getPtsFromCorners(curr_warp, curr_pts.data(), curr_pts_hm.data(), curr_corners);
		dlt_warp = curr_warp;
		if(type == ORC_SSM_HOMOGRAPHY){
			if(normalized_init){
				std::memcpy(init_corners, norm_corners, sizeof(init_corners));
				std::memcpy(init_corners_hm, norm_corners_hm, sizeof(init_corners_hm));
				init_pts = norm_pts; init_pts_hm = norm_pts_hm;
				getStateFromWarp(curr_state.data(), curr_warp);
			} else{
				std::memcpy(init_corners, curr_corners, sizeof(init_corners));
				std::memcpy(init_corners_hm, curr_corners_hm, sizeof(init_corners_hm));
				init_pts = curr_pts; init_pts_hm = curr_pts_hm;        // NB: not re-homogenised (Homography.cc:68)
				curr_warp = identity3();
				std::fill(curr_state.begin(), curr_state.end(), 0.0);
			}
		} else if(type == ORC_SSM_TRANSLATION){
			// Translation::setCorners Translation.cc:56-66
			std::memcpy(init_corners, curr_corners, sizeof(init_corners));
			init_pts = curr_pts;
			homogenize_corners(init_corners, init_corners_hm);
			homogenize(init_pts.data(), init_pts_hm.data(), n_pts);
			curr_warp = identity3();
			std::fill(curr_state.begin(), curr_state.end(), 0.0);
		} else{
			std::memcpy(init_corners, curr_corners, sizeof(init_corners));
			init_pts = curr_pts;
			homogenize_corners(init_corners, init_corners_hm);         // Affine.cc:81-82: re-homogenised
			homogenize(init_pts.data(), init_pts_hm.data(), n_pts);
			curr_warp = identity3();
			std::fill(curr_state.begin(), curr_state.end(), 0.0);
		}
		return true;
	}
	// Homography::getWarpFromState Homography.cc:94-107 ; Affine::getWarpFromState Affine.cc:117-131
	void getWarpFromState(Mat3 &w, const double *s) const{
		if(type == ORC_SSM_HOMOGRAPHY){
			w(0, 0) = 1 + s[0]; w(0, 1) = s[1]; w(0, 2) = s[2];
			w(1, 0) = s[3]; w(1, 1) = 1 + s[4]; w(1, 2) = s[5];
			w(2, 0) = s[6]; w(2, 1) = s[7]; w(2, 2) = 1;
		} else if(type == ORC_SSM_TRANSLATION){                              // Translation.cc:93-101
			w = identity3(); w(0, 2) = s[0]; w(1, 2) = s[1];
		} else{
			w(0, 0) = 1 + s[2]; w(0, 1) = s[3]; w(0, 2) = s[0];
			w(1, 0) = s[4]; w(1, 1) = 1 + s[5]; w(1, 2) = s[1];
			w(2, 0) = 0; w(2, 1) = 0; w(2, 2) = 1;
		}
	}
	// Homography::getStateFromWarp Homography.cc:116-132 ; Affine::getStateFromWarp Affine.cc:133-143
	void getStateFromWarp(double *s, const Mat3 &w) const{
		if(type == ORC_SSM_HOMOGRAPHY){
			s[0] = w(0, 0) - 1; s[1] = w(0, 1); s[2] = w(0, 2);
			s[3] = w(1, 0); s[4] = w(1, 1) - 1; s[5] = w(1, 2);
			s[6] = w(2, 0); s[7] = w(2, 1);
		} else if(type == ORC_SSM_TRANSLATION){                              // Translation.cc:103-109
			s[0] = w(0, 2); s[1] = w(1, 2);
		} else{
			s[0] = w(0, 2); s[1] = w(1, 2); s[2] = w(0, 0) - 1;
			s[3] = w(0, 1); s[4] = w(1, 0); s[5] = w(1, 1) - 1;
		}
	}
	// Affine: curr_pts = curr_warp.topRows<2>() * init_pts_hm  (Affine.cc:103-104)
	void affine_pts(){
		for(int i = 0; i < n_pts; ++i){
			const double *p = &init_pts_hm[3 * i];
			for(int r = 0; r < 2; ++r){
				double s = curr_warp(r, 0)*p[0];
				s = s + curr_warp(r, 1)*p[1];
				s = s + curr_warp(r, 2)*p[2];
				curr_pts[2 * i + r] = s;
			}
		}
		for(int i = 0; i < 4; ++i){
			const double *p = &init_corners_hm[3 * i];
			for(int r = 0; r < 2; ++r){
				double s = curr_warp(r, 0)*p[0];
				s = s + curr_warp(r, 1)*p[1];
				s = s + curr_warp(r, 2)*p[2];
				curr_corners[4 * r + i] = s;
			}
		}
	}
	void projective_pts(){
		warp_hm(curr_warp, init_pts_hm.data(), curr_pts_hm.data(), n_pts);
		warp_hm(curr_warp, init_corners_hm, curr_corners_hm, 4);
		dehomogenize(curr_pts_hm.data(), curr_pts.data(), n_pts);
		dehomogenize_corners(curr_corners_hm, curr_corners);
	}
	// Homography::compositionalUpdate Homography.cc:73-92 ; Affine::compositionalUpdate Affine.cc:90-107
	void compositionalUpdate(const double *state_update){
		if(type == ORC_SSM_TRANSLATION){
			// Translation::compositionalUpdate Translation.cc:80-91: the points are moved by the UPDATE (they accumulate their
			// own rounding, one addition per iteration), not recomputed from the state
			curr_state[0] += state_update[0]; curr_state[1] += state_update[1];
			curr_warp(0, 2) = curr_state[0]; curr_warp(1, 2) = curr_state[1];
			for(int i = 0; i < n_pts; ++i){ curr_pts[2 * i] += state_update[0]; curr_pts[2 * i + 1] += state_update[1]; }
			for(int i = 0; i < 4; ++i){ curr_corners[i] += state_update[0]; curr_corners[4 + i] += state_update[1]; }
			return;
		}
		Mat3 warp_update_mat;
		getWarpFromState(warp_update_mat, state_update);
		curr_warp = mul3(curr_warp, warp_update_mat);
		if(type == ORC_SSM_HOMOGRAPHY){
			double d = curr_warp(2, 2);
			for(int i = 0; i < 9; ++i) curr_warp.m[i] = curr_warp.m[i] / d;    // Eigen 3.3 operator/=: true division
			getStateFromWarp(curr_state.data(), curr_warp);
			projective_pts();
		} else{
			getStateFromWarp(curr_state.data(), curr_warp);
			affine_pts();
		}
	}
	// ProjectiveBase::setState ProjectiveBase.cc:41-49 ; Affine::setState Affine.cc:109-115
	void setState(const double *ssm_state){
		for(int i = 0; i < state_size; ++i) curr_state[i] = ssm_state[i];
		getWarpFromState(curr_warp, curr_state.data());
		if(type == ORC_SSM_TRANSLATION){                                       // Translation::setState Translation.cc:68-78
			for(int i = 0; i < n_pts; ++i){ curr_pts[2 * i] = init_pts[2 * i] + curr_state[0]; curr_pts[2 * i + 1] = init_pts[2 * i + 1] + curr_state[1]; }
			for(int i = 0; i < 4; ++i){ curr_corners[i] = init_corners[i] + curr_state[0]; curr_corners[4 + i] = init_corners[4 + i] + curr_state[1]; }
			return;
		}
		if(type == ORC_SSM_HOMOGRAPHY) projective_pts(); else affine_pts();
	}
	// Homography::invertState Homography.cc:109-114 ; Affine::invertState Affine.cc:145-150
	void invertState(double *inv_state, const double *state) const{
		if(type == ORC_SSM_TRANSLATION){ inv_state[0] = -state[0]; inv_state[1] = -state[1]; return; }   // Translation.cc:111-113
		Mat3 warp_mat, inv_warp_mat;
		getWarpFromState(warp_mat, state);
		inv_warp_mat = inverse3(warp_mat);
		double d = inv_warp_mat(2, 2);
		for(int i = 0; i < 9; ++i) inv_warp_mat.m[i] = inv_warp_mat.m[i] / d;
		getStateFromWarp(inv_state, inv_warp_mat);
	}
	// Homography::updateGradPts Homography.cc:803-827 ; Affine::updateGradPts Affine.cc:293-312
	void updateGradPts(double grad_eps){
		grad_pts.resize(8 * (size_t)n_pts);
		if(type == ORC_SSM_TRANSLATION){                                       // Translation.cc:115-130
			for(int i = 0; i < n_pts; ++i){
				double x = curr_pts[2 * i], y = curr_pts[2 * i + 1]; double *g = &grad_pts[8 * (size_t)i];
				g[0] = x + grad_eps; g[1] = y; g[2] = x - grad_eps; g[3] = y;
				g[4] = x; g[5] = y + grad_eps; g[6] = x; g[7] = y - grad_eps;
			}
			return;
		}
		if(type == ORC_SSM_HOMOGRAPHY){
			double dx[3], dy[3];
			for(int r = 0; r < 3; ++r){ dx[r] = curr_warp(r, 0)*grad_eps; dy[r] = curr_warp(r, 1)*grad_eps; }
			for(int i = 0; i < n_pts; ++i){
				const double *h = &curr_pts_hm[3 * i]; double *g = &grad_pts[8 * (size_t)i];
				double a[3];
				for(int r = 0; r < 3; ++r) a[r] = h[r] + dx[r];
				g[0] = a[0] / a[2]; g[1] = a[1] / a[2];
				for(int r = 0; r < 3; ++r) a[r] = h[r] - dx[r];
				g[2] = a[0] / a[2]; g[3] = a[1] / a[2];
				for(int r = 0; r < 3; ++r) a[r] = h[r] + dy[r];
				g[4] = a[0] / a[2]; g[5] = a[1] / a[2];
				for(int r = 0; r < 3; ++r) a[r] = h[r] - dy[r];
				g[6] = a[0] / a[2]; g[7] = a[1] / a[2];
			}
		} else{
			double dx0 = curr_warp(0, 0)*grad_eps, dx1 = curr_warp(1, 0)*grad_eps;
			double dy0 = curr_warp(0, 1)*grad_eps, dy1 = curr_warp(1, 1)*grad_eps;
			for(int i = 0; i < n_pts; ++i){
				double x = curr_pts[2 * i], y = curr_pts[2 * i + 1]; double *g = &grad_pts[8 * (size_t)i];
				g[0] = x + dx0; g[1] = y + dx1; g[2] = x - dx0; g[3] = y - dx1;
				g[4] = x + dy0; g[5] = y + dy1; g[6] = x - dy0; g[7] = y - dy1;
			}
		}
	}
	// Homography::cmptInitPixJacobian Homography.cc:157-191 ; Affine::cmptInitPixJacobian Affine.cc:160-181
	void cmptInitPixJacobian(double *dI_dp, const double *dI_dw) const{
		const int N = n_pts;
		if(type == ORC_SSM_TRANSLATION){                                       // Translation.h:45-63: dI_dp = dI_dx
			for(int i = 0; i < 2 * N; ++i) dI_dp[i] = dI_dw[i];
			return;
		}
		for(int i = 0; i < N; ++i){
			double x = init_pts[2 * i], y = init_pts[2 * i + 1];
			double Ix = dI_dw[i], Iy = dI_dw[N + i];
			if(type == ORC_SSM_HOMOGRAPHY){
				double Ixx = Ix*x, Iyy = Iy*y, Ixy = Ix*y, Iyx = Iy*x;
				dI_dp[0 * N + i] = Ixx; dI_dp[1 * N + i] = Ixy; dI_dp[2 * N + i] = Ix;
				dI_dp[3 * N + i] = Iyx; dI_dp[4 * N + i] = Iyy; dI_dp[5 * N + i] = Iy;
				dI_dp[6 * N + i] = -x*Ixx - y*Iyx;
				dI_dp[7 * N + i] = -x*Ixy - y*Iyy;
			} else{
				dI_dp[0 * N + i] = Ix; dI_dp[1 * N + i] = Iy;
				dI_dp[2 * N + i] = Ix*x; dI_dp[3 * N + i] = Ix*y;
				dI_dp[4 * N + i] = Iy*x; dI_dp[5 * N + i] = Iy*y;
			}
		}
	}
	// ProjectiveBase::additiveUpdate ProjectiveBase.cc:51-55 ; Affine::additiveUpdate (same: curr_state += update; setState)
	void additiveUpdate(const double *state_update){
		vec ns(state_size);
		for(int i = 0; i < state_size; ++i) ns[i] = curr_state[i] + state_update[i];
		setState(ns.data());
	}
	// Homography::cmptPixJacobian Homography.cc:193-229 ; Affine::cmptPixJacobian = cmptInitPixJacobian (Affine.h:35-37)
	void cmptPixJacobian(double *dI_dp, const double *dI_dw) const{
		if(type != ORC_SSM_HOMOGRAPHY){ cmptInitPixJacobian(dI_dp, dI_dw); return; }
		const int N = n_pts;
		for(int i = 0; i < N; ++i){
			double x = init_pts[2 * i], y = init_pts[2 * i + 1];
			double curr_x = curr_pts[2 * i], curr_y = curr_pts[2 * i + 1];
			double inv_d = 1.0 / curr_pts_hm[3 * i + 2];
			double Ix = dI_dw[i] * inv_d, Iy = dI_dw[N + i] * inv_d;
			double Ixx = Ix*x, Iyy = Iy*y, Ixy = Ix*y, Iyx = Iy*x;
			dI_dp[0 * N + i] = Ixx; dI_dp[1 * N + i] = Ixy; dI_dp[2 * N + i] = Ix;
			dI_dp[3 * N + i] = Iyx; dI_dp[4 * N + i] = Iyy; dI_dp[5 * N + i] = Iy;
			dI_dp[6 * N + i] = (-curr_x*Ixx - curr_y*Iyx);
			dI_dp[7 * N + i] = (-curr_x*Ixy - curr_y*Iyy);
		}
	}
	// Homography::cmptApproxPixJacobian Homography.cc:296-358 ; Affine::cmptApproxPixJacobian Affine.cc:183-211
	void cmptApproxPixJacobian(double *dI_dp, const double *dI_dw) const{
		const int N = n_pts;
		if(type == ORC_SSM_TRANSLATION){ cmptInitPixJacobian(dI_dp, dI_dw); return; }
		if(type == ORC_SSM_HOMOGRAPHY){
			Mat3 W = curr_warp;
			double h00_plus_1 = W(0, 0), h01 = W(0, 1), h10 = W(1, 0), h11_plus_1 = W(1, 1), h20 = W(2, 0), h21 = W(2, 1);
			for(int i = 0; i < N; ++i){
				double curr_x = curr_pts[2 * i], curr_y = curr_pts[2 * i + 1];
				double a = (h00_plus_1 - h20*curr_x), b = (h01 - h21*curr_x);
				double c = (h10 - h20*curr_y), d = (h11_plus_1 - h21*curr_y);
				double inv_factor = 1.0 / (a*d - b*c);
				double x = init_pts[2 * i], y = init_pts[2 * i + 1];
				double Ix = (d*dI_dw[i] - c*dI_dw[N + i])*inv_factor;
				double Iy = (a*dI_dw[N + i] - b*dI_dw[i])*inv_factor;
				double Ixx = Ix*x, Ixy = Ix*y, Iyy = Iy*y, Iyx = Iy*x;
				dI_dp[0 * N + i] = Ixx; dI_dp[1 * N + i] = Ixy; dI_dp[2 * N + i] = Ix;
				dI_dp[3 * N + i] = Iyx; dI_dp[4 * N + i] = Iyy; dI_dp[5 * N + i] = Iy;
				dI_dp[6 * N + i] = (-curr_x*Ixx - curr_y*Iyx);
				dI_dp[7 * N + i] = (-curr_x*Ixy - curr_y*Iyy);
			}
		} else{
			double a = curr_state[2] + 1, b = curr_state[3], c = curr_state[4], d = curr_state[5] + 1;
			double inv_det = 1.0 / (a*d - b*c);
			for(int i = 0; i < N; ++i){
				double x = init_pts[2 * i], y = init_pts[2 * i + 1];
				double Ix = dI_dw[i], Iy = dI_dw[N + i];
				double Ixx = Ix*x, Ixy = Ix*y, Iyy = Iy*y, Iyx = Iy*x;
				dI_dp[0 * N + i] = (Ix*d - Iy*c) * inv_det; dI_dp[1 * N + i] = (Iy*a - Ix*b) * inv_det;
				dI_dp[2 * N + i] = (Ixx*d - Iyx*c) * inv_det; dI_dp[3 * N + i] = (Ixy*d - Iyy*c) * inv_det;
				dI_dp[4 * N + i] = (Iyx*a - Ixx*b) * inv_det; dI_dp[5 * N + i] = (Iyy*a - Ixy*b) * inv_det;
			}
		}
	}
	// Homography::cmptWarpedPixJacobian Homography.cc:231-294 ; Affine::cmptWarpedPixJacobian Affine.cc:213-242
	void cmptWarpedPixJacobian(double *dI_dp, const double *dI_dw) const{
		const int N = n_pts;
		if(type == ORC_SSM_TRANSLATION){ cmptInitPixJacobian(dI_dp, dI_dw); return; }
		if(type == ORC_SSM_HOMOGRAPHY){
			double a00 = curr_warp(0, 0), a01 = curr_warp(0, 1), a10 = curr_warp(1, 0);
			double a11 = curr_warp(1, 1), a20 = curr_warp(2, 0), a21 = curr_warp(2, 1);
			for(int i = 0; i < N; ++i){
				double w_x = curr_pts[2 * i], w_y = curr_pts[2 * i + 1];
				double D = curr_pts_hm[3 * i + 2];
				double inv_det = 1.0 / D;
				double dwx_dx = (a00 - a20*w_x), dwx_dy = (a01 - a21*w_x);
				double dwy_dx = (a10 - a20*w_y), dwy_dy = (a11 - a21*w_y);
				double x = init_pts[2 * i], y = init_pts[2 * i + 1];
				double Ix = (dwx_dx*dI_dw[i] + dwy_dx*dI_dw[N + i])*inv_det;
				double Iy = (dwx_dy*dI_dw[i] + dwy_dy*dI_dw[N + i])*inv_det;
				double Ixx = Ix*x, Ixy = Ix*y, Iyy = Iy*y, Iyx = Iy*x;
				dI_dp[0 * N + i] = Ixx; dI_dp[1 * N + i] = Ixy; dI_dp[2 * N + i] = Ix;
				dI_dp[3 * N + i] = Iyx; dI_dp[4 * N + i] = Iyy; dI_dp[5 * N + i] = Iy;
				dI_dp[6 * N + i] = -x*Ixx - y*Iyx;
				dI_dp[7 * N + i] = -x*Ixy - y*Iyy;
			}
		} else{
			double a = curr_state[2] + 1, b = curr_state[3], c = curr_state[4], d = curr_state[5] + 1;
			for(int i = 0; i < N; ++i){
				double x = init_pts[2 * i], y = init_pts[2 * i + 1];
				double Ix = dI_dw[i], Iy = dI_dw[N + i];
				double Ixx = Ix*x, Ixy = Ix*y, Iyy = Iy*y, Iyx = Iy*x;
				dI_dp[0 * N + i] = Ix*a + Iy*c; dI_dp[1 * N + i] = Ix*b + Iy*d;
				dI_dp[2 * N + i] = Ixx*a + Iyx*c; dI_dp[3 * N + i] = Ixy*a + Iyy*c;
				dI_dp[4 * N + i] = Ixx*b + Iyx*d; dI_dp[5 * N + i] = Ixy*b + Iyy*d;
			}
		}
	}
};

// ---------------------------------------------------------------------------------------------
// AM: ImageBase + SSDBase/SSD + NCC + MI
// ---------------------------------------------------------------------------------------------
// Utilities/include/mtf/Utilities/histUtils.h:206-224
static inline void bSpl3WithGrad(double &val, double &diff, double x){
	const double _2_BY_3 = 2.0 / 3.0;
	if((x > -2) && (x <= -1)){
		double temp = 2 + x; diff = (temp * temp) / 2; val = (diff * temp) / 3;
	} else if((x > -1) && (x <= 0)){
		double temp = x / 2; val = _2_BY_3 - x*x*(1 + temp); diff = -x * (temp + x + 2);
	} else if((x > 0) && (x <= 1)){
		double temp = x / 2; val = _2_BY_3 - x*x*(1 - temp); diff = x * (temp + x - 2);
	} else if((x > 1) && (x < 2)){
		double temp = 2 - x; diff = -(temp * temp) / 2; val = -(diff * temp) / 3;
	}
}
// histUtils.h:269-280
static inline double bSpl3Hess(double x){
	if((x > -2) && (x <= -1)){ return 2 + x; }
	else if((x > -1) && (x <= 0)){ return -(3 * x + 2); }
	else if((x > 0) && (x <= 1)){ return 3 * x - 2; }
	else if((x > 1) && (x < 2)){ return 2 - x; }
	return 0;
}

struct AM{
	int type, resx, resy, n_pix, patch_size;
	double grad_eps, pix_norm_mult, pix_norm_add, likelihood_alpha; int grad_mode; bool fast_sums; bool blocked_syrk;
	const float *img; unsigned int img_height, img_width;
	vec I0, It, dI0_dx, dIt_dx, df_dI0, df_dIt;
	double f;
	bool init_pix_vals, init_pix_grad, init_sim, init_grad, init_hess;
	// NCC (AM/include/mtf/AM/NCC.h:99-106)
	double I0_mean, It_mean, a, b, c, bc, b2c, df_dI0_ncntr_mean, df_dIt_ncntr_mean;
	vec I0_cntr, It_cntr, I0_cntr_c, It_cntr_b, df_dI0_ncntr, df_dIt_ncntr;
	// MI
	int n_bins; double pre_seed; bool pou;
	double hist_pre_seed, hist_norm_mult, max_similarity;
	int joint_hist_size;
	std::vector<int> std_bspl_ids, init_bspl_ids, curr_bspl_ids;   // [2*i + {0,1}]
	vec init_hist, curr_hist, init_hist_log, curr_hist_log, joint_hist, joint_hist_log;
	vec init_hist_mat, curr_hist_mat, init_hist_grad, curr_hist_grad, init_hist_hess, curr_hist_hess;
	vec init_grad_factor, curr_grad_factor, init_joint_hist_grad, curr_joint_hist_grad;
	vec self_joint_hist, self_joint_hist_log, self_grad_factor;

	AM(const orc_params &p) : type(p.am), resx(p.resx), resy(p.resy), n_pix(p.resx*p.resy),
		patch_size(p.resx*p.resy), grad_eps(p.grad_eps), pix_norm_mult(1), pix_norm_add(0),
		likelihood_alpha(p.likelihood_alpha), grad_mode(p.grad_mode), fast_sums(p.fast_sums != 0), blocked_syrk(p.fast_sums == 2), img(nullptr), img_height(0), img_width(0), f(0),
		init_pix_vals(false), init_pix_grad(false), init_sim(false), init_grad(false), init_hess(false),
		n_bins(p.mi_n_bins), pre_seed(p.mi_pre_seed), pou(p.mi_pou != 0){
		if(type == ORC_AM_MI){                                                 // MI::MI AM/src/MI.cc:55-123
			double norm_pix_min = 0, norm_pix_max = n_bins - 1;
			if(pou){ norm_pix_min = 1; norm_pix_max = n_bins - 2; }
			pix_norm_mult = (norm_pix_max - norm_pix_min) / (255.0 - 0.0 + 1);   // PIX_MAX 255, PIX_MIN 0
			pix_norm_add = norm_pix_min;
			joint_hist_size = n_bins*n_bins;
			hist_pre_seed = n_bins*pre_seed;
			hist_norm_mult = 1.0 / (static_cast<double>(patch_size) + hist_pre_seed*n_bins);
			std_bspl_ids.resize(2 * n_bins);
			for(int i = 0; i < n_bins; i++){
				std_bspl_ids[2 * i] = std::max(0, i - 1);
				std_bspl_ids[2 * i + 1] = std::min(n_bins - 1, i + 2);
			}
		}
	}
	void clearInitStatus(){ init_pix_vals = init_pix_grad = init_sim = init_grad = init_hess = false; }
	void setCurrImg(const float *_img, int h, int w){ img = _img; img_height = h; img_width = w; }     // ImageBase.cc:38-60
	// ImageBase::initializePixVals ImageBase.cc:62-99
	void initializePixVals(const double *init_pts){
		if(!init_pix_vals){ I0.resize(patch_size); It.resize(patch_size); }
		getPixVals(I0.data(), img, init_pts, n_pix, img_height, img_width, pix_norm_mult, pix_norm_add);
		if(!init_pix_vals){ It = I0; init_pix_vals = true; }
	}
	// ImageBase::initializePixGrad(Matrix2Xd) ImageBase.cc:101-132
	void initializePixGrad(const double *init_pts){
		if(!init_pix_grad){ dI0_dx.resize(2 * (size_t)patch_size); dIt_dx.resize(2 * (size_t)patch_size); }
		if(grad_mode) getImgGradAnalytic(dI0_dx.data(), img, init_pts, grad_eps, n_pix, img_height, img_width, pix_norm_mult);
		else getImgGrad(dI0_dx.data(), img, init_pts, grad_eps, n_pix, img_height, img_width, pix_norm_mult);
		if(!init_pix_grad){ dIt_dx = dI0_dx; init_pix_grad = true; }
	}
	// ImageBase::initializePixGrad(Matrix8Xd) ImageBase.cc:134-166
	void initializePixGradWarped(const double *warped_offset_pts){
		if(!init_pix_grad){ dI0_dx.resize(2 * (size_t)patch_size); dIt_dx.resize(2 * (size_t)patch_size); }
		getWarpedImgGrad(dI0_dx.data(), img, warped_offset_pts, grad_eps, n_pix, img_height, img_width, pix_norm_mult);
		if(!init_pix_grad){ dIt_dx = dI0_dx; init_pix_grad = true; }
	}
	// ImageBase::updatePixVals ImageBase.cc:268-290 (MI::updatePixVals MI.cc:166-192 is identical for 32FC1)
	void updatePixVals(const double *curr_pts){
		getPixVals(It.data(), img, curr_pts, n_pix, img_height, img_width, pix_norm_mult, pix_norm_add);
	}
	// ImageBase::updatePixGrad(Matrix2Xd) ImageBase.cc:292-314
	void updatePixGrad(const double *curr_pts){
		if(grad_mode) getImgGradAnalytic(dIt_dx.data(), img, curr_pts, grad_eps, n_pix, img_height, img_width, pix_norm_mult);
		else getImgGrad(dIt_dx.data(), img, curr_pts, grad_eps, n_pix, img_height, img_width, pix_norm_mult);
	}
	// ImageBase::updatePixGrad(Matrix8Xd) ImageBase.cc:340-362
	void updatePixGradWarped(const double *warped_offset_pts){
		getWarpedImgGrad(dIt_dx.data(), img, warped_offset_pts, grad_eps, n_pix, img_height, img_width, pix_norm_mult);
	}

	// ------------------------------------------------------------------ similarity
	void initializeSimilarity(){
		switch(type){
		case ORC_AM_SSD:                                                     // SSDBase.cc:29-45
			if(init_sim) return;
			df_dI0.assign(patch_size, 0.0);                                    // I_diff aliases df_dI0 (SSDBase.cc:34)
			f = 0; init_sim = true; break;
		case ORC_AM_NCC:                                                     // NCC.cc:55-95
			if(!init_sim){ I0_cntr.resize(patch_size); It_cntr.resize(patch_size); }
			I0_mean = mean(I0);
			for(int i = 0; i < patch_size; ++i) I0_cntr[i] = I0[i] - I0_mean;
			c = norm(I0_cntr);
			if(!init_sim){ f = 1; It_mean = I0_mean; It_cntr = I0_cntr; b = c; init_sim = true; }
			break;
		case ORC_AM_MI: mi_initializeSimilarity(); break;
		}
	}
	void initializeGrad(){
		switch(type){
		case ORC_AM_SSD:                                                     // SSDBase.cc:47-63
			if(init_grad) return;
			df_dIt = df_dI0; init_grad = true; break;
		case ORC_AM_NCC:                                                     // NCC.cc:97-122
			if(!init_grad){
				df_dIt.assign(patch_size, 0.0); df_dI0.assign(patch_size, 0.0);
				df_dI0_ncntr.assign(patch_size, 0.0); df_dIt_ncntr.assign(patch_size, 0.0);
				I0_cntr_c.resize(patch_size); It_cntr_b.resize(patch_size);
				df_dI0_ncntr_mean = df_dIt_ncntr_mean = 0;
			}
			for(int i = 0; i < patch_size; ++i) I0_cntr_c[i] = I0_cntr[i] / c;
			if(!init_grad){ It_cntr_b = I0_cntr_c; init_grad = true; }
			break;
		case ORC_AM_MI: mi_initializeGrad(); break;
		}
	}
	void initializeHess(){
		if(type == ORC_AM_MI) mi_initializeHess();
	}
	static double mean(const vec &v){ double s = 0; for(size_t i = 0; i < v.size(); ++i) s += v[i]; return s / double(v.size()); }
	static double norm(const vec &v){ double s = 0; for(size_t i = 0; i < v.size(); ++i) s += v[i] * v[i]; return std::sqrt(s); }

	void updateSimilarity(bool prereq_only){
		switch(type){
		case ORC_AM_SSD:{                                                    // SSDBase.cc:75-96
			for(int i = 0; i < patch_size; ++i) df_dI0[i] = It[i] - I0[i];     // I_diff = It - I0
			if(prereq_only) return;
			double s = 0; for(int i = 0; i < patch_size; ++i) s += df_dI0[i] * df_dI0[i];
			f = -s / 2; break; }
		case ORC_AM_NCC:{                                                    // NCC.cc:124-161
			It_mean = mean(It);
			for(int i = 0; i < patch_size; ++i) It_cntr[i] = It[i] - It_mean;
			double s = 0; for(int i = 0; i < patch_size; ++i) s += I0_cntr[i] * It_cntr[i];
			a = s; b = norm(It_cntr);
			bc = b*c; b2c = bc*b; f = a / bc; break; }
		case ORC_AM_MI: mi_updateSimilarity(prereq_only); break;
		}
	}
	void updateInitGrad(){
		switch(type){
		case ORC_AM_SSD: break;                                              // SSDBase.h:67-72 (no ILM: nothing)
		case ORC_AM_NCC:                                                     // NCC.cc:163-194
			df_dI0_ncntr_mean = 0;
			for(int i = 0; i < patch_size; ++i){
				It_cntr_b[i] = It_cntr[i] / b;
				df_dI0_ncntr[i] = (It_cntr_b[i] - f*I0_cntr_c[i]) / c;
				df_dI0_ncntr_mean += df_dI0_ncntr[i];
			}
			df_dI0_ncntr_mean /= patch_size;
			for(int i = 0; i < patch_size; ++i) df_dI0[i] = df_dI0_ncntr[i] - df_dI0_ncntr_mean;
			break;
		case ORC_AM_MI: mi_updateInitGrad(); break;
		}
	}
	void updateCurrGrad(){
		switch(type){
		case ORC_AM_SSD:                                                     // SSDBase.cc:115-121
			for(int i = 0; i < patch_size; ++i) { df_dIt[i] = -df_dI0[i]; } break;
		case ORC_AM_NCC:                                                     // NCC.cc:196-234
			df_dIt_ncntr_mean = 0;
			for(int i = 0; i < patch_size; ++i){
				It_cntr_b[i] = It_cntr[i] / b;
				df_dIt_ncntr[i] = (I0_cntr_c[i] - f*It_cntr_b[i]) / b;
				df_dIt_ncntr_mean += df_dIt_ncntr[i];
			}
			df_dIt_ncntr_mean /= patch_size;
			for(int i = 0; i < patch_size; ++i) df_dIt[i] = df_dIt_ncntr[i] - df_dIt_ncntr_mean;
			break;
		case ORC_AM_MI: mi_updateCurrGrad(); break;
		}
	}
	// SSD.h:41-43 ; NCC.cc:50-53 ; MI.cc:384-387
	double getLikelihood() const{
		if(type == ORC_AM_SSD) return std::exp(-likelihood_alpha * std::sqrt(-f / (static_cast<double>(patch_size))));
		double d = (1.0 / f) - 1;
		return std::exp(-likelihood_alpha * d*d);
	}

	// ------------------------------------------------------------------ Jacobians: (1xN).(NxS), sequential dot per column
	// fast_sums (timing only): lets the compiler vectorise the length-N dot products the way Eigen's product kernels
	// do, instead of one sequential add chain; changes the summation order, so parity tests keep it off
	static double dot(const double *a, const double *b, int N, bool fast){
		double acc = 0;
		if(fast){
			#pragma omp simd reduction(+:acc)
			for(int i = 0; i < N; ++i) acc += a[i] * b[i];
		} else{
			for(int i = 0; i < N; ++i) acc += a[i] * b[i];
		}
		return acc;
	}
	void gemv(double *out, const vec &v, const double *M, int N, int S) const{
		for(int s = 0; s < S; ++s) out[s] = dot(v.data(), M + (size_t)s*N, N, fast_sums);
	}
	// SSDBase.cc:123-143 / NCC.cc:236-250 / AppearanceModel.h:146-149
	void cmptInitJacobian(double *df_dp, const double *dI0_dp, int S) const{ gemv(df_dp, df_dI0, dI0_dp, patch_size, S); }
	// SSDBase.cc:144-168 / NCC.cc:252-266 / AppearanceModel.h:150-153
	void cmptCurrJacobian(double *df_dp, const double *dIt_dp, int S) const{ gemv(df_dp, df_dIt, dIt_dp, patch_size, S); }
	void cmptDifferenceOfJacobians(double *df_dp_diff, const double *dI0_dp, const double *dIt_dp, int S) const{
		const int N = patch_size;
		if(type == ORC_AM_SSD){                                               // SSDBase.cc:186: df_dIt * (dI0_dp + dIt_dp)
			for(int s = 0; s < S; ++s){
				double acc = 0;
				for(int i = 0; i < N; ++i) acc += df_dIt[i] * (dI0_dp[(size_t)s*N + i] + dIt_dp[(size_t)s*N + i]);
				df_dp_diff[s] = acc;
			}
		} else{                                                              // NCC.cc:276 ; AppearanceModel.h:162-166
			vec t(S), u(S);
			gemv(t.data(), df_dIt, dIt_dp, N, S); gemv(u.data(), df_dI0, dI0_dp, N, S);
			for(int s = 0; s < S; ++s) df_dp_diff[s] = t[s] - u[s];
		}
	}
	// ------------------------------------------------------------------ Hessians (column-major SxS out)
	void neg_JtJ(double *H, const double *J, int N, int S) const{
		if(blocked_syrk){
			// fast_sums = 2 (timing only): a cache-blocked SYRK on the upper triangle -- 256-pixel blocks of the S columns stay in
			// L1 while their S (S + 1) / 2 vectorised dot products are taken.  Does LESS work than Eigen's general product
			// `dI_dp.transpose() * dI_dp` (SSDBase.cc:262-280 computes all S^2 entries): an upper bound on what the reference's
			// Hessian stage can reach on this CPU
			double acc[8][8];
			for(int i = 0; i < S; ++i) for(int j = 0; j < S; ++j) acc[i][j] = 0;
			const int B = 256;
			for(int n0 = 0; n0 < N; n0 += B){
				const int nb = N - n0 < B ? N - n0 : B;
				for(int i = 0; i < S; ++i){
					const double *a = J + (size_t)i*N + n0;
					for(int j = i; j < S; ++j){
						const double *b = J + (size_t)j*N + n0;
						double sum = 0;
						#pragma omp simd reduction(+:sum)
						for(int n = 0; n < nb; ++n) sum += a[n] * b[n];
						acc[i][j] += sum;
					}
				}
			}
			for(int i = 0; i < S; ++i) for(int j = i; j < S; ++j){ H[(size_t)j*S + i] = -acc[i][j]; H[(size_t)i*S + j] = -acc[i][j]; }
			return;
		}
		for(int i = 0; i < S; ++i) for(int j = 0; j < S; ++j)
			H[(size_t)j*S + i] = -dot(J + (size_t)i*N, J + (size_t)j*N, N, fast_sums);
	}
	// NCC helper: (dI_dp.rowwise() - dI_dp.colwise().mean()).array() / b
	void ncc_center(vec &out, const double *dI_dp, int S) const{
		const int N = patch_size; out.resize((size_t)N*S);
		for(int s = 0; s < S; ++s){
			double m = 0; for(int i = 0; i < N; ++i) m += dI_dp[(size_t)s*N + i];
			m /= double(N);
			for(int i = 0; i < N; ++i) out[(size_t)s*N + i] = (dI_dp[(size_t)s*N + i] - m) / b;
		}
	}
	static void mat_t_vec(double *out, const vec &M, const vec &v, int N, int S){
		for(int s = 0; s < S; ++s){ double acc = 0; for(int i = 0; i < N; ++i) acc += M[(size_t)s*N + i] * v[i]; out[s] = acc; }
	}
	void cmptSelfHessian(double *H, const double *dIt_dp, int S){
		const int N = patch_size;
		switch(type){
		case ORC_AM_SSD: neg_JtJ(H, dIt_dp, N, S); break;                    // SSDBase.h:91-94 -> SSDBase.cc:268-285
		case ORC_AM_NCC:{                                                    // NCC.cc:337-389 (fast_hess = 0)
			vec Jc; ncc_center(Jc, dIt_dp, S);
			neg_JtJ(H, Jc.data(), N, S);
			vec v(S); mat_t_vec(v.data(), Jc, It_cntr_b, N, S);
			for(int i = 0; i < S; ++i) for(int j = 0; j < S; ++j) H[(size_t)j*S + i] += v[i] * v[j];
			break; }
		case ORC_AM_MI: mi_cmptSelfHessian(H, dIt_dp, S); break;
		}
	}
	void cmptCurrHessian(double *H, const double *dIt_dp, int S){
		const int N = patch_size;
		switch(type){
		case ORC_AM_SSD: neg_JtJ(H, dIt_dp, N, S); break;                    // SSDBase.cc:268-285
		case ORC_AM_NCC:{                                                    // NCC.cc:307-336
			vec Jc; ncc_center(Jc, dIt_dp, S);
			neg_JtJ(H, Jc.data(), N, S);
			for(int i = 0; i < S*S; ++i) H[i] *= f;                           // -f * Jc^T Jc
			vec vb(S), vc(S); mat_t_vec(vb.data(), Jc, It_cntr_b, N, S); mat_t_vec(vc.data(), Jc, I0_cntr_c, N, S);
			for(int i = 0; i < S; ++i) for(int j = 0; j < S; ++j)
				H[(size_t)j*S + i] += -vb[i] * vc[j] - vc[i] * vb[j] + 3 * vb[i] * vb[j];
			break; }
		case ORC_AM_MI: mi_cmptCurrHessian(H, dIt_dp, S); break;
		}
	}
	void cmptInitHessian(double *H, const double *dI0_dp, int S){
		const int N = patch_size;
		switch(type){
		case ORC_AM_SSD: neg_JtJ(H, dI0_dp, N, S); break;                    // SSDBase.cc:251-267
		case ORC_AM_NCC:{                                                    // NCC.cc:282-306 (divides by b, not c: quirk kept)
			vec Jc; ncc_center(Jc, dI0_dp, S);
			neg_JtJ(H, Jc.data(), N, S);
			for(int i = 0; i < S*S; ++i) H[i] *= f;
			vec vb(S), vc(S); mat_t_vec(vb.data(), Jc, It_cntr_b, N, S); mat_t_vec(vc.data(), Jc, I0_cntr_c, N, S);
			for(int i = 0; i < S; ++i) for(int j = 0; j < S; ++j)
				H[(size_t)j*S + i] += -vb[i] * vc[j] - vc[i] * vb[j] + 3 * vc[i] * vc[j];
			break; }
		case ORC_AM_MI: mi_cmptInitHessian(H, dI0_dp, S); break;
		}
	}
	// SSDBase.cc:287-311 ; AppearanceModel.h default = init + curr
	void cmptSumOfHessians(double *H, const double *dI0_dp, const double *dIt_dp, int S){
		if(type == ORC_AM_SSD){
			const int N = patch_size;
			for(int i = 0; i < S; ++i) for(int j = 0; j < S; ++j){
				double a0 = 0, at = 0;
				for(int k = 0; k < N; ++k) a0 += dI0_dp[(size_t)i*N + k] * dI0_dp[(size_t)j*N + k];
				for(int k = 0; k < N; ++k) at += dIt_dp[(size_t)i*N + k] * dIt_dp[(size_t)j*N + k];
				H[(size_t)j*S + i] = -(a0 + at);
			}
		} else{
			vec H0((size_t)S*S), Ht((size_t)S*S);
			cmptInitHessian(H0.data(), dI0_dp, S); cmptCurrHessian(Ht.data(), dIt_dp, S);
			for(int i = 0; i < S*S; ++i) H[i] = H0[i] + Ht[i];
		}
	}

	// ------------------------------------------------------------------ MI (AM/src/MI.cc)
	double& HM(vec &m, int id, int pix){ return m[(size_t)pix*n_bins + id]; }    // n_bins x N column-major
	double& JH(vec &m, int r, int c){ return m[(size_t)c*n_bins + r]; }          // n_bins x n_bins column-major
	void mi_initializeSimilarity(){                                              // MI.cc:207-290
		const int B = n_bins, N = patch_size;
		if(!init_sim){
			init_hist.resize(B); init_hist_mat.resize((size_t)B*N); joint_hist.resize(B*B);
			init_hist_log.resize(B); joint_hist_log.resize(B*B);
			curr_hist.resize(B); curr_hist_mat.resize((size_t)B*N); curr_hist_log.resize(B);
			init_bspl_ids.resize(2 * N); curr_bspl_ids.resize(2 * N);
			init_hist_grad.resize((size_t)B*N); curr_hist_grad.resize((size_t)B*N);
		}
		std::fill(init_hist.begin(), init_hist.end(), hist_pre_seed);
		std::fill(init_hist_mat.begin(), init_hist_mat.end(), 0.0);
		std::fill(init_hist_grad.begin(), init_hist_grad.end(), 0.0);
		for(int pix_id = 0; pix_id < N; pix_id++){
			int bin = static_cast<int>(I0[pix_id]);
			init_bspl_ids[2 * pix_id] = std_bspl_ids[2 * bin]; init_bspl_ids[2 * pix_id + 1] = std_bspl_ids[2 * bin + 1];
			double curr_diff = init_bspl_ids[2 * pix_id] - I0[pix_id];
			for(int id1 = init_bspl_ids[2 * pix_id]; id1 <= init_bspl_ids[2 * pix_id + 1]; id1++){
				bSpl3WithGrad(HM(init_hist_mat, id1, pix_id), HM(init_hist_grad, id1, pix_id), curr_diff);
				HM(init_hist_grad, id1, pix_id) *= -hist_norm_mult;
				init_hist[id1] += HM(init_hist_mat, id1, pix_id);
				++curr_diff;
			}
		}
		for(int i = 0; i < B; ++i){ init_hist[i] *= hist_norm_mult; init_hist_log[i] = std::log(init_hist[i]); }
		if(!init_sim){
			std::fill(joint_hist.begin(), joint_hist.end(), pre_seed);
			for(int pix_id = 0; pix_id < N; pix_id++){
				for(int id1 = init_bspl_ids[2 * pix_id]; id1 <= init_bspl_ids[2 * pix_id + 1]; id1++){
					for(int id2 = init_bspl_ids[2 * pix_id]; id2 <= init_bspl_ids[2 * pix_id + 1]; id2++){
						JH(joint_hist, id1, id2) += HM(init_hist_mat, id1, pix_id) * HM(init_hist_mat, id2, pix_id);
					}
				}
			}
			for(int i = 0; i < B*B; ++i){ joint_hist[i] *= hist_norm_mult; joint_hist_log[i] = std::log(joint_hist[i]); }
			f = 0;
			for(int curr_id = 0; curr_id < B; curr_id++){
				for(int init_id = 0; init_id < B; init_id++){
					f += JH(joint_hist, curr_id, init_id) * (JH(joint_hist_log, curr_id, init_id) -
						init_hist_log[curr_id] - init_hist_log[init_id]);
				}
			}
			max_similarity = f;
			curr_bspl_ids = init_bspl_ids; curr_hist = init_hist; curr_hist_mat = init_hist_mat;
			curr_hist_log = init_hist_log; curr_hist_grad = init_hist_grad;
			init_sim = true;
		}
	}
	void mi_initializeGrad(){                                                    // MI.cc:302-335
		const int B = n_bins, N = patch_size;
		if(!init_grad){
			init_joint_hist_grad.assign((size_t)B*B*N, 0.0); curr_joint_hist_grad.resize((size_t)B*B*N);
			init_grad_factor.resize(B*B); curr_grad_factor.resize(B*B);
			df_dIt.resize(N); df_dI0.assign(N, 0.0);
			for(int curr_id = 0; curr_id < B; curr_id++)
				for(int init_id = 0; init_id < B; init_id++)
					JH(init_grad_factor, curr_id, init_id) = 1 + JH(joint_hist_log, curr_id, init_id) - init_hist_log[curr_id];
			for(int pix_id = 0; pix_id < N; pix_id++){
				for(int curr_id = init_bspl_ids[2 * pix_id]; curr_id <= init_bspl_ids[2 * pix_id + 1]; curr_id++){
					for(int init_id = init_bspl_ids[2 * pix_id]; init_id <= init_bspl_ids[2 * pix_id + 1]; init_id++){
						int joint_id = curr_id*B + init_id;
						double g = HM(init_hist_grad, curr_id, pix_id) * HM(init_hist_mat, init_id, pix_id);
						init_joint_hist_grad[(size_t)pix_id*B*B + joint_id] = g;
						df_dI0[pix_id] += g * JH(init_grad_factor, curr_id, init_id);
					}
				}
			}
			curr_grad_factor = init_grad_factor; curr_joint_hist_grad = init_joint_hist_grad;
			df_dIt = df_dI0; init_grad = true;
		}
	}
	void mi_initializeHess(){                                                    // MI.cc:444-460 + histUtils.cc:234-256
		const int B = n_bins, N = patch_size;
		if(!init_hess){
			init_hist_hess.resize((size_t)B*N); curr_hist_hess.resize((size_t)B*N);
			self_joint_hist.resize(B*B); self_joint_hist_log.resize(B*B); self_grad_factor.resize(B*B);
		}
		std::fill(init_hist_hess.begin(), init_hist_hess.end(), 0.0);
		for(int pix = 0; pix < N; pix++){
			double curr_diff = init_bspl_ids[2 * pix] - I0[pix];
			for(int id = init_bspl_ids[2 * pix]; id <= init_bspl_ids[2 * pix + 1]; id++){
				HM(init_hist_hess, id, pix) = hist_norm_mult*bSpl3Hess(curr_diff++);
			}
		}
		if(!init_hess){ curr_hist_hess = init_hist_hess; init_hess = true; }
	}
	void mi_updateSimilarity(bool prereq_only){                                  // MI.cc:346-382
		const int B = n_bins, N = patch_size;
		std::fill(curr_hist.begin(), curr_hist.end(), hist_pre_seed);
		std::fill(joint_hist.begin(), joint_hist.end(), pre_seed);
		std::fill(curr_hist_mat.begin(), curr_hist_mat.end(), 0.0);
		std::fill(curr_hist_grad.begin(), curr_hist_grad.end(), 0.0);
		for(int pix_id = 0; pix_id < N; pix_id++){
			int bin = static_cast<int>(It[pix_id]);
			curr_bspl_ids[2 * pix_id] = std_bspl_ids[2 * bin]; curr_bspl_ids[2 * pix_id + 1] = std_bspl_ids[2 * bin + 1];
			double curr_diff = curr_bspl_ids[2 * pix_id] - It[pix_id];
			for(int curr_id = curr_bspl_ids[2 * pix_id]; curr_id <= curr_bspl_ids[2 * pix_id + 1]; curr_id++){
				bSpl3WithGrad(HM(curr_hist_mat, curr_id, pix_id), HM(curr_hist_grad, curr_id, pix_id), curr_diff);
				++curr_diff;
				HM(curr_hist_grad, curr_id, pix_id) *= -hist_norm_mult;
				curr_hist[curr_id] += HM(curr_hist_mat, curr_id, pix_id);
				for(int init_id = init_bspl_ids[2 * pix_id]; init_id <= init_bspl_ids[2 * pix_id + 1]; init_id++){
					JH(joint_hist, curr_id, init_id) += HM(curr_hist_mat, curr_id, pix_id) * HM(init_hist_mat, init_id, pix_id);
				}
			}
		}
		for(int i = 0; i < B; ++i){ curr_hist[i] *= hist_norm_mult; }
		for(int i = 0; i < B*B; ++i){ joint_hist[i] *= hist_norm_mult; }
		for(int i = 0; i < B; ++i){ curr_hist_log[i] = std::log(curr_hist[i]); }
		for(int i = 0; i < B*B; ++i){ joint_hist_log[i] = std::log(joint_hist[i]); }
		if(prereq_only){ return; }
		f = 0;
		for(int curr_id = 0; curr_id < B; curr_id++){
			for(int init_id = 0; init_id < B; init_id++){
				f += JH(joint_hist, curr_id, init_id) * (JH(joint_hist_log, curr_id, init_id) - curr_hist_log[curr_id] - init_hist_log[init_id]);
			}
		}
	}
	void mi_updateInitGrad(){                                                    // MI.cc:398-417
		const int B = n_bins, N = patch_size;
		for(int init_id = 0; init_id < B; init_id++)
			for(int curr_id = 0; curr_id < B; curr_id++)
				JH(init_grad_factor, init_id, curr_id) = 1 + JH(joint_hist_log, curr_id, init_id) - init_hist_log[init_id];
		std::fill(init_joint_hist_grad.begin(), init_joint_hist_grad.end(), 0.0);
		for(int pix_id = 0; pix_id < N; pix_id++){
			df_dI0[pix_id] = 0;
			for(int init_id = init_bspl_ids[2 * pix_id]; init_id <= init_bspl_ids[2 * pix_id + 1]; init_id++){
				for(int curr_id = curr_bspl_ids[2 * pix_id]; curr_id <= curr_bspl_ids[2 * pix_id + 1]; curr_id++){
					int joint_id = init_id*B + curr_id;
					double g = HM(init_hist_grad, init_id, pix_id) * HM(curr_hist_mat, curr_id, pix_id);
					init_joint_hist_grad[(size_t)pix_id*B*B + joint_id] = g;
					df_dI0[pix_id] += g * JH(init_grad_factor, init_id, curr_id);
				}
			}
		}
	}
	void mi_updateCurrGrad(){                                                    // MI.cc:426-443
		const int B = n_bins, N = patch_size;
		for(int curr_id = 0; curr_id < B; curr_id++)
			for(int init_id = 0; init_id < B; init_id++)
				JH(curr_grad_factor, curr_id, init_id) = 1 + JH(joint_hist_log, curr_id, init_id) - curr_hist_log[curr_id];
		std::fill(curr_joint_hist_grad.begin(), curr_joint_hist_grad.end(), 0.0);
		for(int pix_id = 0; pix_id < N; pix_id++){
			df_dIt[pix_id] = 0;
			for(int curr_id = curr_bspl_ids[2 * pix_id]; curr_id <= curr_bspl_ids[2 * pix_id + 1]; curr_id++){
				for(int init_id = init_bspl_ids[2 * pix_id]; init_id <= init_bspl_ids[2 * pix_id + 1]; init_id++){
					int joint_id = curr_id*B + init_id;
					double g = HM(curr_hist_grad, curr_id, pix_id) * HM(init_hist_mat, init_id, pix_id);
					curr_joint_hist_grad[(size_t)pix_id*B*B + joint_id] = g;
					df_dIt[pix_id] += g * JH(curr_grad_factor, curr_id, init_id);
				}
			}
		}
	}
	// add hist_factor * row^T row for every joint bin; rows of joint_hist_jacobian are (B*B) x S row-major here
	void mi_add_joint_terms(double *H, const vec &jhj, const vec &jh, const vec &hist, bool denom_is_row, int S){
		const int B = n_bins;
		for(int curr_id = 0; curr_id < B; curr_id++){
			for(int init_id = 0; init_id < B; init_id++){
				int joint_id = curr_id*B + init_id;
				double hist_factor = (1.0 / jh[(size_t)init_id*B + curr_id]) - (1.0 / hist[denom_is_row ? curr_id : init_id]);
				const double *row = &jhj[(size_t)joint_id*S];
				for(int i = 0; i < S; ++i) for(int j = 0; j < S; ++j) H[(size_t)j*S + i] += row[i] * row[j] * hist_factor;
			}
		}
	}
	void mi_cmptInitHessian(double *H, const double *J, int S){                    // MI.cc:461-514
		const int B = n_bins, N = patch_size;
		vec jhj((size_t)B*B*S, 0.0);
		std::fill(H, H + S*S, 0.0);
		for(int pix_id = 0; pix_id < N; pix_id++){
			double hist_hess_term = 0;
			for(int init_id = init_bspl_ids[2 * pix_id]; init_id <= init_bspl_ids[2 * pix_id + 1]; init_id++){
				double inner_term = 0;
				for(int curr_id = curr_bspl_ids[2 * pix_id]; curr_id <= curr_bspl_ids[2 * pix_id + 1]; curr_id++){
					int joint_id = curr_id*B + init_id;
					double g = init_joint_hist_grad[(size_t)pix_id*B*B + init_id*B + curr_id];
					for(int s = 0; s < S; ++s) jhj[(size_t)joint_id*S + s] += g*J[(size_t)s*N + pix_id];
					inner_term += HM(curr_hist_mat, curr_id, pix_id) * JH(init_grad_factor, init_id, curr_id);
				}
				hist_hess_term += HM(init_hist_hess, init_id, pix_id)*inner_term;
			}
			for(int i = 0; i < S; ++i) for(int j = 0; j < S; ++j)
				H[(size_t)j*S + i] += hist_hess_term * J[(size_t)i*N + pix_id] * J[(size_t)j*N + pix_id];
		}
		mi_add_joint_terms(H, jhj, joint_hist, init_hist, false, S);            // 1/joint_hist(curr,init) - 1/init_hist(init)
	}
	void mi_cmptSelfHist(){                                                      // MI.cc:639-658
		const int B = n_bins, N = patch_size;
		std::fill(self_joint_hist.begin(), self_joint_hist.end(), pre_seed);
		for(int pix_id = 0; pix_id < N; pix_id++){
			for(int id1 = curr_bspl_ids[2 * pix_id]; id1 <= curr_bspl_ids[2 * pix_id + 1]; id1++){
				for(int id2 = curr_bspl_ids[2 * pix_id]; id2 <= curr_bspl_ids[2 * pix_id + 1]; id2++){
					JH(self_joint_hist, id1, id2) += HM(curr_hist_mat, id1, pix_id) * HM(curr_hist_mat, id2, pix_id);
				}
			}
		}
		for(int i = 0; i < B*B; ++i){ self_joint_hist[i] *= hist_norm_mult; self_joint_hist_log[i] = std::log(self_joint_hist[i]); }
		for(int curr_id = 0; curr_id < B; curr_id++)
			for(int init_id = 0; init_id < B; init_id++)
				JH(self_grad_factor, curr_id, init_id) = 1 + JH(self_joint_hist_log, curr_id, init_id) - curr_hist_log[curr_id];
	}
	void mi_cmptSelfHessian(double *H, const double *J, int S){                    // MI.cc:515-594 (the returned half, :567-590)
		const int B = n_bins, N = patch_size;
		mi_cmptSelfHist();
		vec jhj((size_t)B*B*S, 0.0);
		std::fill(H, H + S*S, 0.0);
		for(int pix_id = 0; pix_id < N; pix_id++){
			double curr_diff = curr_bspl_ids[2 * pix_id] - It[pix_id];
			double hist_hess_term = 0;
			for(int curr_id = curr_bspl_ids[2 * pix_id]; curr_id <= curr_bspl_ids[2 * pix_id + 1]; curr_id++){
				HM(curr_hist_hess, curr_id, pix_id) = hist_norm_mult*bSpl3Hess(curr_diff);
				++curr_diff;
				double inner_term = 0;
				for(int init_id = curr_bspl_ids[2 * pix_id]; init_id <= curr_bspl_ids[2 * pix_id + 1]; init_id++){
					int idx = curr_id*B + init_id;
					double g = HM(curr_hist_grad, curr_id, pix_id)*HM(curr_hist_mat, init_id, pix_id);
					for(int s = 0; s < S; ++s) jhj[(size_t)idx*S + s] += g*J[(size_t)s*N + pix_id];
					inner_term += HM(curr_hist_mat, init_id, pix_id) * JH(self_grad_factor, curr_id, init_id);
				}
				hist_hess_term += HM(curr_hist_hess, curr_id, pix_id)*inner_term;
			}
			for(int i = 0; i < S; ++i) for(int j = 0; j < S; ++j)
				H[(size_t)j*S + i] += hist_hess_term * J[(size_t)i*N + pix_id] * J[(size_t)j*N + pix_id];
		}
		mi_add_joint_terms(H, jhj, self_joint_hist, curr_hist, true, S);        // 1/self_joint(curr,init) - 1/curr_hist(curr)
	}
	void mi_cmptCurrHessian(double *H, const double *J, int S){                    // MI.cc:603-637
		const int B = n_bins, N = patch_size;
		vec jhj((size_t)B*B*S, 0.0);
		std::fill(H, H + S*S, 0.0);
		for(int pix_id = 0; pix_id < N; pix_id++){
			double curr_diff = curr_bspl_ids[2 * pix_id] - It[pix_id];
			double hist_hess_term = 0;
			for(int curr_id = curr_bspl_ids[2 * pix_id]; curr_id <= curr_bspl_ids[2 * pix_id + 1]; curr_id++){
				HM(curr_hist_hess, curr_id, pix_id) = hist_norm_mult*bSpl3Hess(curr_diff);
				++curr_diff;
				double inner_term = 0;
				for(int init_id = init_bspl_ids[2 * pix_id]; init_id <= init_bspl_ids[2 * pix_id + 1]; init_id++){
					int idx = curr_id*B + init_id;
					double g = curr_joint_hist_grad[(size_t)pix_id*B*B + idx];
					for(int s = 0; s < S; ++s) jhj[(size_t)idx*S + s] += g*J[(size_t)s*N + pix_id];
					inner_term += HM(init_hist_mat, init_id, pix_id) * JH(curr_grad_factor, curr_id, init_id);
				}
				hist_hess_term += HM(curr_hist_hess, curr_id, pix_id)*inner_term;
			}
			for(int i = 0; i < S; ++i) for(int j = 0; j < S; ++j)
				H[(size_t)j*S + i] += hist_hess_term * J[(size_t)i*N + pix_id] * J[(size_t)j*N + pix_id];
		}
		mi_add_joint_terms(H, jhj, joint_hist, curr_hist, true, S);             // 1/joint_hist(curr,init) - 1/curr_hist(curr)
	}
};

} // namespace

// ---------------------------------------------------------------------------------------------
// SM: nt::ESM / nt::FCLK / nt::ICLK (and the templated twins' iteration counting)
// ---------------------------------------------------------------------------------------------
struct orc_tracker{
	orc_params params;
	SSM ssm; AM am;
	int S, N;
	vec init_pix_jacobian, curr_pix_jacobian, mean_pix_jacobian;
	vec jacobian, hessian, init_self_hessian, state_update, ssm_update, inv_ssm_update;
	double prev_corners[8];
	int n_iters;
	std::vector<orc_iter_log> log;
	double stage_t[9];

	orc_tracker(const orc_params &p) : params(p), ssm(p.ssm, p.resx, p.resy, p.hom_normalized_init != 0),
		am(p), n_iters(0){
		S = ssm.state_size; N = am.patch_size;
		init_pix_jacobian.assign((size_t)N*S, 0); curr_pix_jacobian.assign((size_t)N*S, 0);
		mean_pix_jacobian.assign((size_t)N*S, 0);
		jacobian.assign(S, 0); hessian.assign((size_t)S*S, 0); init_self_hessian.assign((size_t)S*S, 0);
		state_update.assign(S, 0); ssm_update.assign(S, 0); inv_ssm_update.assign(S, 0);
		for(int i = 0; i < 9; ++i) stage_t[i] = 0;
	}
	// the templated SMs have no chained_warp switch: always chained (ESM.cc:94-95, FCLK.cc:82-86, ICLK.cc:80-90)
	bool chained() const{ return params.chained_warp || !params.nt_semantics; }
	// nt::ESM::initializePixJacobian NT/ESM.cc:376-385 ; templated ESM::initialize ESM.cc:94-95
	void initPixJacobian(vec &out){
		if(chained()){
			am.initializePixGrad(ssm.curr_pts.data());
			ssm.cmptWarpedPixJacobian(out.data(), am.dI0_dx.data());
		} else{
			ssm.updateGradPts(am.grad_eps);
			am.initializePixGradWarped(ssm.grad_pts.data());
			ssm.cmptInitPixJacobian(out.data(), am.dI0_dx.data());
		}
	}
	// nt::ESM::updatePixJacobian NT/ESM.cc:387-404 ; FCLK NT/FCLK.cc:222-236 ; ICLK NT/ICLK.cc:205-219
	void updatePixJacobian(vec &out){
		double t0 = now_s();
		if(chained()){
			am.updatePixGrad(ssm.curr_pts.data());
			double t1 = now_s(); stage_t[3] += t1 - t0;
			ssm.cmptWarpedPixJacobian(out.data(), am.dIt_dx.data());
			stage_t[4] += now_s() - t1;
		} else{
			ssm.updateGradPts(am.grad_eps);
			am.updatePixGradWarped(ssm.grad_pts.data());
			double t1 = now_s(); stage_t[3] += t1 - t0;
			ssm.cmptInitPixJacobian(out.data(), am.dIt_dx.data());
			stage_t[4] += now_s() - t1;
		}
	}
	int initialize(const double *corners){
		am.clearInitStatus();
		if(!ssm.setCorners(corners)) return 1;
		am.initializePixVals(ssm.curr_pts.data());
		switch(params.sm){
		case ORC_SM_ESM:                                                       // NT/ESM.cc:110-146 (== ESM.cc:78-117)
			if(params.nt_semantics){
				initPixJacobian(init_pix_jacobian);
				am.initializeSimilarity(); am.initializeGrad(); am.initializeHess();
			} else{
				am.initializeSimilarity(); am.initializeGrad(); am.initializeHess();
				initPixJacobian(init_pix_jacobian);
			}
			if(params.hess_type == ORC_ESM_HESS_INITIAL_SELF || params.hess_type == ORC_ESM_HESS_SUM_OF_SELF){
				am.cmptSelfHessian(hessian.data(), init_pix_jacobian.data(), S);
				init_self_hessian = hessian;
			}
			break;
		case ORC_SM_FCLK:                                                      // NT/FCLK.cc:102-169
			am.initializeSimilarity(); am.initializeGrad(); am.initializeHess();
			if(chained()){ am.initializePixGrad(ssm.curr_pts.data()); }
			else{ ssm.updateGradPts(am.grad_eps); am.initializePixGradWarped(ssm.grad_pts.data()); }
			if(params.hess_type == ORC_LK_HESS_INITIAL_SELF){
				if(chained()){ ssm.cmptWarpedPixJacobian(init_pix_jacobian.data(), am.dI0_dx.data()); }
				else{ ssm.cmptInitPixJacobian(init_pix_jacobian.data(), am.dI0_dx.data()); }
				am.cmptSelfHessian(hessian.data(), init_pix_jacobian.data(), S);
				if(params.leven_marq){ init_self_hessian = hessian; }
			}
			break;
		case ORC_SM_ICLK:                                                      // NT/ICLK.cc:71-127
			if(chained()){ am.initializePixGrad(ssm.curr_pts.data()); }
			else{ ssm.updateGradPts(am.grad_eps); am.initializePixGradWarped(ssm.grad_pts.data()); }
			am.initializeSimilarity(); am.initializeGrad(); am.initializeHess();
			if(chained()){ ssm.cmptWarpedPixJacobian(init_pix_jacobian.data(), am.dI0_dx.data()); }
			else{ ssm.cmptInitPixJacobian(init_pix_jacobian.data(), am.dI0_dx.data()); }
			am.cmptInitJacobian(jacobian.data(), init_pix_jacobian.data(), S);
			if(params.hess_type == ORC_LK_HESS_INITIAL_SELF){
				am.cmptSelfHessian(hessian.data(), init_pix_jacobian.data(), S);
				if(params.leven_marq){ init_self_hessian = hessian; }
			}
			break;
		case ORC_SM_FALK:                                                      // NT/FALK.cc:93-130
		case ORC_SM_IALK:                                                      // NT/IALK.cc:56-86
			am.initializePixGrad(ssm.curr_pts.data());
			am.initializeSimilarity(); am.initializeGrad(); am.initializeHess();
			if(params.hess_type == ORC_LK_HESS_INITIAL_SELF){
				ssm.cmptPixJacobian(init_pix_jacobian.data(), am.dI0_dx.data());
				am.cmptSelfHessian(hessian.data(), init_pix_jacobian.data(), S);
				if(params.leven_marq){ init_self_hessian = hessian; }
			}
			break;
		default: return 1;
		}
		return 0;
	}
	// FCLK::setRegion NT/FCLK.cc:360-376 ; ESM::setRegion NT/ESM.cc:148-167 ; ICLK (update_ssm=0 default): SSM only
	int setRegion(const double *corners){
		if(!ssm.setCorners(corners)) return 1;
		if(params.sm == ORC_SM_ESM){
			ssm.cmptInitPixJacobian(init_pix_jacobian.data(), am.dI0_dx.data());
			if(params.hess_type == ORC_ESM_HESS_INITIAL_SELF || params.hess_type == ORC_ESM_HESS_SUM_OF_SELF){
				am.cmptSelfHessian(hessian.data(), init_pix_jacobian.data(), S);
				init_self_hessian = hessian;
			}
		} else if(params.sm == ORC_SM_FCLK && params.hess_type == ORC_LK_HESS_INITIAL_SELF){
			ssm.cmptInitPixJacobian(init_pix_jacobian.data(), am.dI0_dx.data());
			am.cmptSelfHessian(hessian.data(), init_pix_jacobian.data(), S);
		}
		return 0;
	}
	void solve(){
		// state_update = -hessian.colPivHouseholderQr().solve(jacobian.transpose())   NT/FCLK.cc:298
		double t0 = now_s();
		ColPivQR qr; qr.compute(hessian.data(), S, S);
		qr.solve(jacobian.data(), state_update.data());
		for(int i = 0; i < S; ++i){ state_update[i] = -state_update[i]; ssm_update[i] = state_update[i]; }
		stage_t[7] += now_s() - t0;
	}
	void lm_damp(double delta){                                                  // NT/FCLK.cc:289-296
		for(int i = 0; i < S; ++i) hessian[(size_t)i*S + i] += delta*hessian[(size_t)i*S + i];
	}
	void record(bool rejected, double update_norm){
		orc_iter_log e; std::memset(&e, 0, sizeof(e));
		e.f = am.f; e.rejected = rejected; e.update_norm = update_norm;
		for(int i = 0; i < S; ++i){ e.jacobian[i] = jacobian[i]; e.state_update[i] = state_update[i]; }
		for(int i = 0; i < S*S; ++i) e.hessian[i] = hessian[i];
		std::memcpy(e.corners, ssm.curr_corners, sizeof(e.corners));
		log.push_back(e);
	}
	double cornerDiffSqNorm() const{
		double s = 0; for(int i = 0; i < 8; ++i){ double d = prev_corners[i] - ssm.curr_corners[i]; s += d*d; } return s;
	}
	int update(){
		log.clear(); for(int i = 0; i < 9; ++i) stage_t[i] = 0;
		switch(params.sm){
		case ORC_SM_FCLK: return update_fclk();
		case ORC_SM_ESM: return update_esm();
		case ORC_SM_ICLK: return update_iclk();
		case ORC_SM_FALK: return update_additive(true);
		case ORC_SM_IALK: return update_additive(false);
		}
		return 1;
	}
	// nt::FALK::update NT/FALK.cc:132-258 (forward = true) ; nt::IALK::update NT/IALK.cc:88-215 (forward = false): the additive
	// searches -- same loop, the pixel Jacobian from the current image's gradient (cmptPixJacobian) or from the template's
	// (cmptApproxPixJacobian), the state update added to the state (ssm->additiveUpdate)
	int update_additive(bool forward){
		double prev_similarity = 0, leven_marq_delta = params.lm_delta_init;
		bool state_reset = false; n_iters = 0;
		vec neg(S);
		for(int iter_id = 0; iter_id < params.max_iters; ++iter_id){
			++n_iters;
			am.updatePixVals(ssm.curr_pts.data());
			am.updateSimilarity(false);
			if(params.leven_marq && !state_reset){
				double curr_similarity = am.f;
				if(iter_id > 0){
					if(curr_similarity < prev_similarity){
						leven_marq_delta *= params.lm_delta_update;
						for(int i = 0; i < S; ++i) neg[i] = -ssm_update[i];
						ssm.additiveUpdate(neg.data());
						state_reset = true;
						record(true, 0);
						continue;
					}
					if(curr_similarity > prev_similarity){ leven_marq_delta /= params.lm_delta_update; }
				}
				prev_similarity = curr_similarity;
			}
			state_reset = false;
			if(forward){
				am.updatePixGrad(ssm.curr_pts.data());
				ssm.cmptPixJacobian(curr_pix_jacobian.data(), am.dIt_dx.data());
			} else{
				ssm.cmptApproxPixJacobian(curr_pix_jacobian.data(), am.dI0_dx.data());
			}
			am.updateCurrGrad();
			am.cmptCurrJacobian(jacobian.data(), curr_pix_jacobian.data(), S);
			switch(params.hess_type){
			case ORC_LK_HESS_INITIAL_SELF: if(params.leven_marq){ hessian = init_self_hessian; } break;
			case ORC_LK_HESS_CURRENT_SELF: am.cmptSelfHessian(hessian.data(), curr_pix_jacobian.data(), S); break;
			case ORC_LK_HESS_STD: am.cmptCurrHessian(hessian.data(), curr_pix_jacobian.data(), S); break;
			}
			if(params.leven_marq){ lm_damp(leven_marq_delta); }
			solve();
			std::memcpy(prev_corners, ssm.curr_corners, sizeof(prev_corners));
			ssm.additiveUpdate(ssm_update.data());
			double update_norm = cornerDiffSqNorm();
			record(false, update_norm);
			if(update_norm < params.epsilon){ break; }
		}
		return 0;
	}
	// nt::FCLK::update NT/FCLK.cc:171-358 ; FCLK<AM,SSM>::update FCLK.cc:106-224
	int update_fclk(){
		int iter_id = 0; n_iters = 0;
		double prev_similarity = 0, leven_marq_delta = params.lm_delta_init;
		bool state_reset = false;
		while(iter_id < params.max_iters){
			++n_iters;
			double t0 = now_s();
			am.updatePixVals(ssm.curr_pts.data()); double t1 = now_s(); stage_t[0] += t1 - t0;
			am.updateSimilarity(false); double t2 = now_s(); stage_t[1] += t2 - t1;
			if(params.leven_marq && !state_reset){
				double curr_similarity = am.f;
				if(iter_id > 0){
					if(curr_similarity < prev_similarity){
						leven_marq_delta *= params.lm_delta_update;
						ssm.invertState(inv_ssm_update.data(), ssm_update.data());
						ssm.compositionalUpdate(inv_ssm_update.data());
						state_reset = true;
						record(true, 0);
						// nt::FCLK: while + continue without ++iter_id (NT/FCLK.cc:187,210);
						// templated: for(...; ++iter_id) (FCLK.cc:117,135)
						if(!params.nt_semantics) ++iter_id;
						continue;
					}
					if(curr_similarity > prev_similarity){ leven_marq_delta /= params.lm_delta_update; }
				}
				prev_similarity = curr_similarity;
			}
			state_reset = false;
			t2 = now_s();
			am.updateCurrGrad(); stage_t[2] += now_s() - t2;
			updatePixJacobian(curr_pix_jacobian);
			double t3 = now_s();
			am.cmptCurrJacobian(jacobian.data(), curr_pix_jacobian.data(), S); double t4 = now_s(); stage_t[5] += t4 - t3;
			switch(params.hess_type){
			case ORC_LK_HESS_INITIAL_SELF: if(params.leven_marq){ hessian = init_self_hessian; } break;
			case ORC_LK_HESS_CURRENT_SELF: am.cmptSelfHessian(hessian.data(), curr_pix_jacobian.data(), S); break;
			case ORC_LK_HESS_STD: am.cmptCurrHessian(hessian.data(), curr_pix_jacobian.data(), S); break;
			}
			stage_t[6] += now_s() - t4;
			if(params.leven_marq){ lm_damp(leven_marq_delta); }
			solve();
			std::memcpy(prev_corners, ssm.curr_corners, sizeof(prev_corners));
			double t5 = now_s();
			ssm.compositionalUpdate(ssm_update.data()); stage_t[8] += now_s() - t5;
			double update_norm = cornerDiffSqNorm();
			record(false, update_norm);
			if(update_norm < params.epsilon){ break; }
			++iter_id;
		}
		return 0;
	}
	// nt::ESM::update NT/ESM.cc:170-297 (SPI disabled at run time, not compile time) ; ESM<AM,SSM>::update ESM.cc:119-292
	int update_esm(){
		double prev_similarity = 0, leven_marq_delta = params.lm_delta_init;
		bool state_reset = false; n_iters = 0;
		for(int iter_id = 0; iter_id < params.max_iters; ++iter_id){
			++n_iters;
			double t0 = now_s();
			am.updatePixVals(ssm.curr_pts.data()); double t1 = now_s(); stage_t[0] += t1 - t0;
			if(params.nt_semantics){ updatePixJacobian(curr_pix_jacobian); }         // NT/ESM.cc:197 (DISABLE_SPI undefined)
			t1 = now_s();
			am.updateSimilarity(false); stage_t[1] += now_s() - t1;
			if(params.leven_marq && !state_reset){
				double curr_similarity = am.f;
				if(iter_id > 0){
					if(curr_similarity < prev_similarity){
						leven_marq_delta *= params.lm_delta_update;
						ssm.invertState(inv_ssm_update.data(), ssm_update.data());
						ssm.compositionalUpdate(inv_ssm_update.data());
						state_reset = true;
						record(true, 0);
						continue;
					}
					if(curr_similarity > prev_similarity){ leven_marq_delta /= params.lm_delta_update; }
				}
				prev_similarity = curr_similarity;
			}
			state_reset = false;
			double t2 = now_s();
			if(params.nt_semantics){
				if(params.jac_type == ORC_ESM_JAC_ORIGINAL || params.hess_type == ORC_ESM_HESS_ORIGINAL){
					for(size_t i = 0; i < mean_pix_jacobian.size(); ++i)
						mean_pix_jacobian[i] = (init_pix_jacobian[i] + curr_pix_jacobian[i]) / 2.0;
				}
				am.updateCurrGrad(); am.updateInitGrad(); stage_t[2] += now_s() - t2;
			} else{
				am.updateCurrGrad(); am.updateInitGrad(); stage_t[2] += now_s() - t2;
				updatePixJacobian(curr_pix_jacobian);                                  // ESM.cc:170-173
				if(params.jac_type == ORC_ESM_JAC_ORIGINAL || params.hess_type == ORC_ESM_HESS_ORIGINAL){
					for(size_t i = 0; i < mean_pix_jacobian.size(); ++i)
						mean_pix_jacobian[i] = (init_pix_jacobian[i] + curr_pix_jacobian[i]) / 2.0;
				}
			}
			double t3 = now_s();
			switch(params.jac_type){                                                  // NT/ESM.cc:299-315
			case ORC_ESM_JAC_ORIGINAL: am.cmptCurrJacobian(jacobian.data(), mean_pix_jacobian.data(), S); break;
			case ORC_ESM_JAC_DIFF_OF_JACS:
				am.cmptDifferenceOfJacobians(jacobian.data(), init_pix_jacobian.data(), curr_pix_jacobian.data(), S);
				for(int i = 0; i < S; ++i) jacobian[i] *= 0.5;
				break;
			}
			double t4 = now_s(); stage_t[5] += t4 - t3;
			switch(params.hess_type){                                                 // NT/ESM.cc:317-374
			case ORC_ESM_HESS_INITIAL_SELF: if(params.leven_marq){ hessian = init_self_hessian; } break;
			case ORC_ESM_HESS_ORIGINAL: am.cmptCurrHessian(hessian.data(), mean_pix_jacobian.data(), S); break;
			case ORC_ESM_HESS_SUM_OF_STD:
				am.cmptSumOfHessians(hessian.data(), init_pix_jacobian.data(), curr_pix_jacobian.data(), S);
				for(int i = 0; i < S*S; ++i) hessian[i] *= 0.5;
				break;
			case ORC_ESM_HESS_SUM_OF_SELF:
				am.cmptSelfHessian(hessian.data(), curr_pix_jacobian.data(), S);
				for(int i = 0; i < S*S; ++i) hessian[i] = (hessian[i] + init_self_hessian[i]) * 0.5;
				break;
			case ORC_ESM_HESS_CURRENT_SELF: am.cmptSelfHessian(hessian.data(), curr_pix_jacobian.data(), S); break;
			case ORC_ESM_HESS_STD: am.cmptCurrHessian(hessian.data(), curr_pix_jacobian.data(), S); break;
			}
			stage_t[6] += now_s() - t4;
			if(params.leven_marq){ lm_damp(leven_marq_delta); }
			solve();
			std::memcpy(prev_corners, ssm.curr_corners, sizeof(prev_corners));
			double t5 = now_s();
			ssm.compositionalUpdate(ssm_update.data()); stage_t[8] += now_s() - t5;
			double update_norm = cornerDiffSqNorm();
			record(false, update_norm);
			if(update_norm < params.epsilon){ break; }
		}
		return 0;
	}
	// nt::ICLK::update NT/ICLK.cc:160-299 ; ICLK<AM,SSM>::update ICLK.cc:144-263
	int update_iclk(){
		double prev_similarity = 0, leven_marq_delta = params.lm_delta_init;
		bool state_reset = false; n_iters = 0;
		for(int iter_id = 0; iter_id < params.max_iters; iter_id++){
			++n_iters;
			double t0 = now_s();
			am.updatePixVals(ssm.curr_pts.data()); double t1 = now_s(); stage_t[0] += t1 - t0;
			am.updateSimilarity(false); double t2 = now_s(); stage_t[1] += t2 - t1;
			if(params.leven_marq && !state_reset){
				double curr_similarity = am.f;
				if(iter_id > 0){
					if(curr_similarity < prev_similarity){
						leven_marq_delta *= params.lm_delta_update;
						ssm.compositionalUpdate(ssm_update.data());                     // undo: re-apply forward update (NT/ICLK.cc:183)
						state_reset = true;
						record(true, 0);
						continue;
					}
					if(curr_similarity > prev_similarity){ leven_marq_delta /= params.lm_delta_update; }
				}
				prev_similarity = curr_similarity;
			}
			state_reset = false;
			t2 = now_s();
			am.updateInitGrad(); double t3 = now_s(); stage_t[2] += t3 - t2;
			am.cmptInitJacobian(jacobian.data(), init_pix_jacobian.data(), S); double t4 = now_s(); stage_t[5] += t4 - t3;
			switch(params.hess_type){
			case ORC_LK_HESS_INITIAL_SELF: if(params.leven_marq){ hessian = init_self_hessian; } break;
			case ORC_LK_HESS_CURRENT_SELF:
				updatePixJacobian(curr_pix_jacobian);
				am.cmptSelfHessian(hessian.data(), curr_pix_jacobian.data(), S); break;
			case ORC_LK_HESS_STD: am.cmptInitHessian(hessian.data(), init_pix_jacobian.data(), S); break;
			}
			stage_t[6] += now_s() - t4;
			if(params.leven_marq){ lm_damp(leven_marq_delta); }
			solve();
			std::memcpy(prev_corners, ssm.curr_corners, sizeof(prev_corners));
			double t5 = now_s();
			ssm.invertState(inv_ssm_update.data(), ssm_update.data());
			ssm.compositionalUpdate(inv_ssm_update.data()); stage_t[8] += now_s() - t5;
			double update_norm = cornerDiffSqNorm();
			record(false, update_norm);
			if(update_norm < params.epsilon){ break; }
		}
		return 0;
	}
};

// ---------------------------------------------------------------------------------------------
// C API
// ---------------------------------------------------------------------------------------------
extern "C" {

void orc_default_params(orc_params *p){
	// shipped cfg (Config/mtf.cfg:14,24 ; modules.cfg) where it overrides parameters.h:153-224
	p->am = ORC_AM_SSD; p->ssm = ORC_SSM_HOMOGRAPHY; p->sm = ORC_SM_FCLK;
	p->resx = 50; p->resy = 50; p->max_iters = 30; p->epsilon = 1e-4;
	p->hess_type = ORC_LK_HESS_CURRENT_SELF; p->jac_type = ORC_ESM_JAC_DIFF_OF_JACS;
	p->chained_warp = 1; p->leven_marq = 0; p->lm_delta_init = 0.01; p->lm_delta_update = 10;
	p->nt_semantics = 1; p->grad_eps = 1e-8; p->hom_normalized_init = 0;
	p->mi_n_bins = 8; p->mi_pre_seed = 10; p->mi_pou = 0; p->likelihood_alpha = 1; p->grad_mode = 0; p->fast_sums = 0;
}
orc_tracker *orc_create(const orc_params *p){ return new orc_tracker(*p); }
void orc_destroy(orc_tracker *t){ delete t; }
void orc_set_image(orc_tracker *t, const float *img, int h, int w){ t->am.setCurrImg(img, h, w); }
int orc_initialize(orc_tracker *t, const double *corners){ return t->initialize(corners); }
int orc_update(orc_tracker *t){ return t->update(); }
int orc_set_region(orc_tracker *t, const double *corners){ return t->setRegion(corners); }
int orc_n_iters(const orc_tracker *t){ return t->n_iters; }
int orc_n_log(const orc_tracker *t){ return (int)t->log.size(); }
const orc_iter_log *orc_log(const orc_tracker *t, int i){ return &t->log[i]; }
void orc_get_corners(const orc_tracker *t, double *out8){ std::memcpy(out8, t->ssm.curr_corners, 8 * sizeof(double)); }
void orc_get_state(const orc_tracker *t, double *outS){ for(int i = 0; i < t->S; ++i) outS[i] = t->ssm.curr_state[i]; }
int orc_state_size(const orc_tracker *t){ return t->S; }
// ssm.setState (ProjectiveBase.cc:41-49, Affine.cc:109-115): curr_warp = getWarpFromState, curr_pts = warp . init_pts_hm
void orc_set_state(orc_tracker *t, const double *state){ t->ssm.setState(state); }
void orc_get_pts(const orc_tracker *t, double *o){ std::memcpy(o, t->ssm.curr_pts.data(), 2 * (size_t)t->N*sizeof(double)); }
void orc_get_init_pts(const orc_tracker *t, double *o){ std::memcpy(o, t->ssm.init_pts.data(), 2 * (size_t)t->N*sizeof(double)); }
void orc_get_init_pix_vals(const orc_tracker *t, double *o){ std::memcpy(o, t->am.I0.data(), (size_t)t->N*sizeof(double)); }
void orc_get_curr_pix_vals(const orc_tracker *t, double *o){ std::memcpy(o, t->am.It.data(), (size_t)t->N*sizeof(double)); }
void orc_get_curr_pix_grad(const orc_tracker *t, double *o){ std::memcpy(o, t->am.dIt_dx.data(), 2 * (size_t)t->N*sizeof(double)); }
void orc_get_curr_pix_jacobian(const orc_tracker *t, double *o){ std::memcpy(o, t->curr_pix_jacobian.data(), (size_t)t->N*t->S*sizeof(double)); }
void orc_get_init_pix_jacobian(const orc_tracker *t, double *o){ std::memcpy(o, t->init_pix_jacobian.data(), (size_t)t->N*t->S*sizeof(double)); }
double orc_get_similarity(const orc_tracker *t){ return t->am.f; }
void orc_get_init_warp(const orc_tracker *t, double *out9){ std::memcpy(out9, t->ssm.dlt_warp.m, 9 * sizeof(double)); }
void orc_get_stage_times(const orc_tracker *t, double *out9){ std::memcpy(out9, t->stage_t, sizeof(t->stage_t)); }

// nt::PF::update particle loop NT/PF.cc:303-320 (state supplied by the caller instead of the Boost sampler)
void orc_pf_evaluate(orc_tracker *t, const double *states, int n_particles, double *likelihood, double *similarity){
	for(int p = 0; p < n_particles; ++p){
		t->ssm.setState(states + (size_t)p*t->S);
		t->am.updatePixVals(t->ssm.curr_pts.data());
		t->am.updateSimilarity(false);
		if(similarity) similarity[p] = t->am.f;
		likelihood[p] = t->am.getLikelihood();
	}
}

double orc_pix_val(const float *img, int h, int w, double x, double y){ return getPixVal(img, x, y, h, w); }
void orc_get_pix_vals(const float *img, int h, int w, const double *pts, int n, double norm_mult, double norm_add, double *out){
	getPixVals(out, img, pts, n, h, w, norm_mult, norm_add);
}
void orc_get_img_grad(const float *img, int h, int w, const double *pts, int n, double grad_eps, double pix_mult, double *o){
	getImgGrad(o, img, pts, grad_eps, n, h, w, pix_mult);
}
void orc_get_img_grad_analytic(const float *img, int h, int w, const double *pts, int n, double grad_eps, double pix_mult, double *o){
	getImgGradAnalytic(o, img, pts, grad_eps, n, h, w, pix_mult);
}
void orc_homography_dlt(const double *in_c, const double *out_c, double *H9){
	Mat3 H = computeHomographyDLT(in_c, out_c); std::memcpy(H9, H.m, sizeof(H.m));
}
void orc_affine_ndlt(const double *in_c, const double *out_c, double *H9){
	Mat3 H = computeAffineNDLT(in_c, out_c); std::memcpy(H9, H.m, sizeof(H.m));
}
void orc_colpiv_qr_solve(const double *A, const double *b, int n, double *x){
	ColPivQR qr; qr.compute(A, n, n); qr.solve(b, x);
}
// the factorisation itself, for cross-checks against LAPACK dgeqp3: qr_out = the rows x cols matrix Eigen's matrixQR() holds
// (R in the upper triangle, Householder vectors below), perm_out[i] = original column at position i, returns nonzeroPivots()
int orc_colpiv_qr(const double *A, int rows, int cols, double *qr_out, int *perm_out, double *hcoeffs_out){
	ColPivQR qr; qr.compute(A, rows, cols);
	for(size_t i = 0; i < (size_t)rows*cols; ++i) qr_out[i] = qr.qr[i];
	for(int i = 0; i < cols; ++i) perm_out[i] = qr.perm[i];
	for(int i = 0; i < qr.size; ++i) hcoeffs_out[i] = qr.hCoeffs[i];
	return qr.nonzero_pivots;
}
void orc_norm_unit_square_pts(int resx, int resy, double min_x, double min_y, double max_x, double max_y,
	double *pts2N, double *corners8){
	getNormUnitSquarePts(pts2N, corners8, resx, resy, min_x, min_y, max_x, max_y);
}

// ---- pre-processing of the raw frame: what PreProcBase::processFrame does for output_type CV_32FC1
// with the default GaussianSmoothing (Utilities/src/preprocUtils.cc:108-127, Utilities/include/mtf/Utilities/preprocUtils.h:67-78,
// Config/include/mtf/Config/parameters.h:229-235: pre_proc_type "gauss", kernel 5, sigma 3):
//   frame_raw.convertTo(CV_32F) -> cv::cvtColor(BGR2GRAY) (3-channel input) -> cv::GaussianBlur(ksize 5, sigma, sigma)
// The arithmetic lives in OpenCV (not vendored by MTF; ReadMe.md:116 asks for 2.4 / 3.x), restated from its published
// algorithm: gray = B*0.114f + G*0.587f + R*0.299f in float (imgproc color conversion, RGB2Gray<float>); the kernel of
// cv::getGaussianKernel(5, sigma, CV_32F): exp(-x^2 / (2 sigma^2)) rounded to float, normalised by the double sum of the
// floats; separable filtering, row pass then column pass through a float buffer, symmetric 5-tap form
//   s = x0*k0 + (x-1 + x+1)*k1 + (x-2 + x+2)*k2       (SymmRowSmallVec_32f / SymmColumnSmallVec_32f32f)
// borders BORDER_REFLECT_101 (cv::borderInterpolate).  OpenCV's SIMD builds fuse some of these multiply-adds, so cv2
// itself reproduces this only to 1-2 ulp (tests/golden/make_preproc_golden.py, tests/test_preproc.py).
void orc_gaussian_kernel5(double sigma, float *k5){
	const double scale2X = -0.5 / (sigma*sigma);
	double sum = 0;
	for(int i = 0; i < 5; ++i){
		const double x = i - 2.0;
		k5[i] = (float)std::exp(scale2X*x*x);
		sum += k5[i];
	}
	sum = 1. / sum;
	for(int i = 0; i < 5; ++i) k5[i] = (float)(k5[i] * sum);
}
static inline int reflect101(int i, int n){
	if(n == 1) return 0;
	while(i < 0 || i >= n){ if(i < 0) i = -i; else i = 2 * n - 2 - i; }
	return i;
}
void orc_preproc_gauss5(const unsigned char *img, int h, int w, int stride, int channels, double sigma, float *out){
	float k[5];
	orc_gaussian_kernel5(sigma, k);
	const float k0 = k[2], k1 = k[1], k2 = k[0];
	std::vector<float> gray((size_t)h*w), tmp((size_t)h*w);
	for(int y = 0; y < h; ++y){
		const unsigned char *r = img + (size_t)y*stride;
		for(int x = 0; x < w; ++x){
			if(channels == 3){
				const float b = r[3 * x], g = r[3 * x + 1], rr = r[3 * x + 2];
				float v = b*0.114f; v = v + g*0.587f; v = v + rr*0.299f;
				gray[(size_t)y*w + x] = v;
			} else gray[(size_t)y*w + x] = r[x];
		}
	}
	for(int y = 0; y < h; ++y){
		const float *a = &gray[(size_t)y*w];
		for(int x = 0; x < w; ++x){
			float s = a[x] * k0;
			s = s + (a[reflect101(x - 1, w)] + a[reflect101(x + 1, w)])*k1;
			s = s + (a[reflect101(x - 2, w)] + a[reflect101(x + 2, w)])*k2;
			tmp[(size_t)y*w + x] = s;
		}
	}
	for(int y = 0; y < h; ++y){
		const float *c0 = &tmp[(size_t)y*w], *m1 = &tmp[(size_t)reflect101(y - 1, h)*w], *p1 = &tmp[(size_t)reflect101(y + 1, h)*w],
			*m2 = &tmp[(size_t)reflect101(y - 2, h)*w], *p2 = &tmp[(size_t)reflect101(y + 2, h)*w];
		for(int x = 0; x < w; ++x){
			float s = c0[x] * k0;
			s = s + (m1[x] + p1[x])*k1;
			s = s + (m2[x] + p2[x])*k2;
			out[(size_t)y*w + x] = s;
		}
	}
}

// GridTracker.cc:247-264 fan-out: independent trackers, one per patch, OpenMP over patches
long orc_batch_track(const orc_params *p, const float *const *frames, int n_frames, int h, int w,
	const double *corners, int n_patches, int n_threads, double *final_corners, int *iters_per_patch, double *seconds){
	std::vector<orc_tracker*> tr(n_patches);
#ifdef _OPENMP
	if(n_threads > 0) omp_set_num_threads(n_threads);
#endif
	#pragma omp parallel for schedule(dynamic)
	for(int i = 0; i < n_patches; ++i){
		tr[i] = new orc_tracker(*p);
		tr[i]->am.setCurrImg(frames[0], h, w);
		tr[i]->initialize(corners + 8 * (size_t)i);
		if(iters_per_patch) iters_per_patch[i] = 0;
	}
	long total = 0;
	double t0 = now_s();
	for(int fr = 1; fr < n_frames; ++fr){
		#pragma omp parallel for schedule(dynamic) reduction(+:total)
		for(int i = 0; i < n_patches; ++i){
			tr[i]->am.setCurrImg(frames[fr], h, w);
			tr[i]->update();
			total += tr[i]->n_iters;
			if(iters_per_patch) iters_per_patch[i] += tr[i]->n_iters;
		}
	}
	if(seconds) *seconds = now_s() - t0;
	for(int i = 0; i < n_patches; ++i){
		if(final_corners) std::memcpy(final_corners + 8 * (size_t)i, tr[i]->ssm.curr_corners, 8 * sizeof(double));
		delete tr[i];
	}
	return total;
}

// The same with GridTracker's per-frame reset (grid_reset_at_each_frame = 1, SM/src/GridTracker.cc:265-285, 345-392:
// every cell is re-initialised on the current frame after its update; here at its initial corners, a static grid).
long orc_batch_track_reset(const orc_params *p, const float *const *frames, int n_frames, int h, int w,
	const double *corners, int n_patches, int n_threads, double *final_corners, int *iters_per_patch, double *seconds){
	std::vector<orc_tracker*> tr(n_patches);
#ifdef _OPENMP
	if(n_threads > 0) omp_set_num_threads(n_threads);
#endif
	#pragma omp parallel for schedule(dynamic)
	for(int i = 0; i < n_patches; ++i){
		tr[i] = new orc_tracker(*p);
		tr[i]->am.setCurrImg(frames[0], h, w);
		tr[i]->initialize(corners + 8 * (size_t)i);
		if(iters_per_patch) iters_per_patch[i] = 0;
	}
	long total = 0;
	double t0 = now_s();
	for(int fr = 1; fr < n_frames; ++fr){
		#pragma omp parallel for schedule(dynamic) reduction(+:total)
		for(int i = 0; i < n_patches; ++i){
			tr[i]->am.setCurrImg(frames[fr], h, w);
			tr[i]->update();
			total += tr[i]->n_iters;
			if(iters_per_patch) iters_per_patch[i] += tr[i]->n_iters;
			if(final_corners && fr == n_frames - 1) std::memcpy(final_corners + 8 * (size_t)i, tr[i]->ssm.curr_corners, 8 * sizeof(double));
			tr[i]->initialize(corners + 8 * (size_t)i);
		}
	}
	if(seconds) *seconds = now_s() - t0;
	for(int i = 0; i < n_patches; ++i) delete tr[i];
	return total;
}

// PF particle evaluation of n_objects templates x n_particles states (NT/PF.cc:303-320), OpenMP over (object, particle
// chunk) pairs, every thread on its own copy of the object's tracker.  states: n_objects x n_particles x S.
long orc_batch_pf_evaluate(const orc_params *p, const float *frame0, const float *frame1, int h, int w, const double *corners,
	int n_objects, const double *states, int n_particles, int n_threads, double *likelihood, double *seconds){
#ifdef _OPENMP
	if(n_threads > 0) omp_set_num_threads(n_threads);
	const int nt = n_threads > 0 ? n_threads : omp_get_max_threads();
#else
	const int nt = 1;
#endif
	int chunks = (nt + n_objects - 1) / n_objects;             // particle chunks per object: at least one task per thread
	if(chunks < 1) chunks = 1;
	const int n_tasks = n_objects*chunks;
	std::vector<orc_tracker*> tr(n_tasks);
	#pragma omp parallel for schedule(dynamic)
	for(int k = 0; k < n_tasks; ++k){
		const int o = k / chunks;
		tr[k] = new orc_tracker(*p);
		tr[k]->am.setCurrImg(frame0, h, w);
		tr[k]->initialize(corners + 8 * (size_t)o);
		tr[k]->am.setCurrImg(frame1, h, w);
	}
	const int S = tr[0]->S;
	double t0 = now_s();
	#pragma omp parallel for schedule(dynamic)
	for(int k = 0; k < n_tasks; ++k){
		const int o = k / chunks, c = k % chunks;
		const int lo = (int)((long long)n_particles*c / chunks), hi = (int)((long long)n_particles*(c + 1) / chunks);
		orc_pf_evaluate(tr[k], states + ((size_t)o*n_particles + lo)*S, hi - lo, likelihood + (size_t)o*n_particles + lo, nullptr);
	}
	if(seconds) *seconds = now_s() - t0;
	for(int k = 0; k < n_tasks; ++k) delete tr[k];
	return (long)n_objects*n_particles;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------
// nt::PF (SM/src/NT/PF.cc) with ONE sampler distribution, the random deviates supplied by the caller (the reference draws them
// from Boost generators seeded by random_device: NT/PF.cc:100-110, ProjectiveBase.cc:192-197).
// ---------------------------------------------------------------------------------------------
struct orc_pf{
	orc_tracker *t;
	orc_pf_params p;
	int S, n, n_normals;
	std::vector<double> states[2], ar[2], wts, cum;
	int cur, max_wt_id, n_iters, resampled;
	double prev_corners[8], max_similarity, measurement_factor;
	bool adaptive; double min_eff_particles;

	// ProjectiveBase::generatePerturbation ProjectiveBase.cc:301-306 ; Homography::generatePerturbation Homography.cc:899-915
	void generatePerturbation(double *z, const double *nrm){
		if(t->ssm.type == ORC_SSM_HOMOGRAPHY && p.corner_based_sampling){
			const double tx = p.ssm_mean[0] + p.ssm_sigma[0] * nrm[0], ty = p.ssm_mean[0] + p.ssm_sigma[0] * nrm[1];
			double dc[8];
			for(int c = 0; c < 4; ++c){
				const double dx = p.ssm_mean[1] + p.ssm_sigma[1] * nrm[2 + 2 * c], dy = p.ssm_mean[1] + p.ssm_sigma[1] * nrm[3 + 2 * c];
				dc[c] = (t->ssm.init_corners[c] + dx) + tx; dc[4 + c] = (t->ssm.init_corners[4 + c] + dy) + ty;
			}
			Mat3 W = computeHomographyDLT(t->ssm.init_corners, dc);        // estimateWarpFromCorners Homography.cc:877-883
			const double d = W(2, 2);
			for(int i = 0; i < 9; ++i) W.m[i] = W.m[i] / d;
			t->ssm.getStateFromWarp(z, W);
		} else{
			for(int s = 0; s < S; ++s) z[s] = p.ssm_mean[s] + p.ssm_sigma[s] * nrm[s];
		}
	}
	void normalize_hom(Mat3 &W) const{
		if(t->ssm.type != ORC_SSM_HOMOGRAPHY) return;                   // ProjectiveBase's versions do not normalise
		const double d = W(2, 2);
		for(int i = 0; i < 9; ++i) W.m[i] = W.m[i] / d;
	}
	void perturb(double *st, double *a, const double *nrm){
		double z[8], base[8], base_ar[8];
		generatePerturbation(z, nrm);
		for(int s = 0; s < S; ++s){ base[s] = st[s]; base_ar[s] = a[s]; }
		if(p.update_type == 0){
			if(p.dynamic_model == 0){ for(int s = 0; s < S; ++s) st[s] = base[s] + z[s]; }           // ProjectiveBase.cc:255-259
			else{
				for(int s = 0; s < S; ++s){ const double ns = base[s] + base_ar[s] + z[s]; st[s] = ns; a[s] = p.ar_coeff*(ns - base[s]); }
			}
			return;
		}
		Mat3 Wb, Wz, War;
		t->ssm.getWarpFromState(Wb, base); t->ssm.getWarpFromState(Wz, z);
		if(p.dynamic_model == 0){
			Mat3 W = mul3(Wb, Wz);                                       // Homography.cc:917-926
			normalize_hom(W);
			t->ssm.getStateFromWarp(st, W);
			return;
		}
		t->ssm.getWarpFromState(War, base_ar);
		Mat3 W = mul3(mul3(Wb, War), Wz);                               // Homography.cc:928-942
		normalize_hom(W);
		Mat3 A = mul3(inverse3(Wb), W);
		normalize_hom(A);
		t->ssm.getStateFromWarp(st, W);
		double na[8];
		t->ssm.getStateFromWarp(na, A);
		for(int s = 0; s < S; ++s) a[s] = na[s] * p.ar_coeff;
	}
	void initializeParticles(){                                          // NT/PF.cc:185-198
		const double init_wt = 1.0 / n;
		for(int i = 0; i < n; ++i){
			for(int s = 0; s < S; ++s){ states[cur][(size_t)i*S + s] = t->ssm.curr_state[s]; ar[cur][(size_t)i*S + s] = 0; }
			wts[i] = init_wt;
			cum[i] = i > 0 ? wts[i] + cum[i - 1] : wts[i];
		}
	}
	void multinomialResampling(const double *uni){                        // NT/PF.cc:448-540 (binary and linear pick the same particle)
		const double total = cum[n - 1];
		for(int i = 0; i < n; ++i) cum[i] = cum[i] / total;
		double max_wt = std::numeric_limits<double>::lowest();
		for(int i = 0; i < n; ++i){
			const double u = uni[i];
			int lower = 0, upper = n - 1, id = (lower + upper) / 2;
			while(upper > lower){
				if(cum[id] >= u) upper = id; else lower = id + 1;
				id = (lower + upper) / 2;
			}
			std::memcpy(&states[1 - cur][(size_t)i*S], &states[cur][(size_t)id*S], S*sizeof(double));
			std::memcpy(&ar[1 - cur][(size_t)i*S], &ar[cur][(size_t)id*S], S*sizeof(double));
			if(wts[id] >= max_wt){ max_wt = wts[id]; max_wt_id = i; }
		}
		cur = 1 - cur;
	}
	int update(const double *normals, const double *uniforms){             // NT/PF.cc:207-446
		n_iters = 0;
		for(int it = 0; it < p.max_iters; ++it){
			double max_wt = std::numeric_limits<double>::lowest();
			for(int i = 0; i < n; ++i){
				double *st = &states[cur][(size_t)i*S], *a = &ar[cur][(size_t)i*S];
				perturb(st, a, normals + ((size_t)it*n + i)*n_normals);
				t->ssm.setState(st);
				t->am.updatePixVals(t->ssm.curr_pts.data());
				t->am.updateSimilarity(false);
				const double m = max_similarity - t->am.f;
				double lik;
				if(p.likelihood_func == 0) lik = t->am.getLikelihood();
				else if(p.likelihood_func == 1) lik = measurement_factor * std::exp(-0.5*m / p.measurement_sigma);
				else lik = 1.0 / (1.0 + m);
				wts[i] = lik;
				cum[i] = i == 0 ? wts[i] : wts[i] + cum[i - 1];
				if(wts[i] >= max_wt){ max_wt = wts[i]; max_wt_id = i; }
			}
			bool perform_resampling = true;
			if(adaptive){
				double sum = 0; for(int i = 0; i < n; ++i) sum += wts[i];
				double q = 0; for(int i = 0; i < n; ++i){ const double v = wts[i] / sum; q += v*v; }
				const double n_eff = q == 0 ? 0 : 1.0 / q;
				if(n_eff > min_eff_particles) perform_resampling = false;
			}
			resampled = 0;
			if(perform_resampling && (p.resampling_type == 1 || p.resampling_type == 2)){
				multinomialResampling(uniforms + (size_t)it*n); resampled = 1;
			}
			if(p.mean_type == 0){
				t->ssm.setState(&states[cur][(size_t)max_wt_id*S]);
			} else if(p.mean_type == 1){
				double mean[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };                 // ProjectiveBase::estimateMeanOfSamples :308-314
				for(int i = 0; i < n; ++i) for(int s = 0; s < S; ++s) mean[s] += (states[cur][(size_t)i*S + s] - mean[s]) / (i + 1);
				t->ssm.setState(mean);
			} else{
				double mc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };                   // updateMeanCorners NT/PF.cc:587-594
				for(int i = 0; i < n; ++i){
					t->ssm.setState(&states[cur][(size_t)i*S]);
					for(int k = 0; k < 8; ++k) mc[k] += (t->ssm.curr_corners[k] - mc[k]) / (i + 1);
				}
				t->ssm.setCorners(mc);
			}
			double upd = 0;
			for(int k = 0; k < 8; ++k){ const double d = prev_corners[k] - t->ssm.curr_corners[k]; upd += d*d; }
			std::memcpy(prev_corners, t->ssm.curr_corners, sizeof(prev_corners));
			++n_iters;
			if(upd < p.epsilon) break;
		}
		if(p.reset_to_mean) initializeParticles();
		return 0;
	}
};

extern "C" {

orc_pf *orc_pf_create(const orc_params *tp, const orc_pf_params *pp){
	orc_pf *f = new orc_pf();
	orc_params q = *tp; q.sm = ORC_SM_FCLK;
	f->t = new orc_tracker(q);
	f->p = *pp; f->S = f->t->S; f->n = pp->n_particles;
	f->n_normals = (f->t->ssm.type == ORC_SSM_HOMOGRAPHY && pp->corner_based_sampling) ? 10 : f->S;
	for(int k = 0; k < 2; ++k){ f->states[k].assign((size_t)f->n*f->S, 0); f->ar[k].assign((size_t)f->n*f->S, 0); }
	f->wts.assign(f->n, 0); f->cum.assign(f->n, 0);
	f->cur = 0; f->max_wt_id = 0; f->n_iters = 0; f->resampled = 0;
	const double pi = 3.14159265358979323846;
	f->measurement_factor = 1.0 / std::sqrt(2 * pi*pp->measurement_sigma);
	f->adaptive = pp->adaptive_resampling_thresh > 0 && pp->adaptive_resampling_thresh <= 1;
	f->min_eff_particles = pp->adaptive_resampling_thresh*pp->n_particles;
	f->max_similarity = 0;
	return f;
}
void orc_pf_destroy(orc_pf *f){ if(f){ delete f->t; delete f; } }
void orc_pf_set_image(orc_pf *f, const float *img, int h, int w){ f->t->am.setCurrImg(img, h, w); }
int orc_pf_initialize(orc_pf *f, const double *corners){                  // NT/PF.cc:136-183
	if(!f->t->ssm.setCorners(corners)) return 1;
	f->t->am.initializePixVals(f->t->ssm.curr_pts.data());
	f->t->am.initializeSimilarity();
	f->max_similarity = f->t->am.f;
	f->initializeParticles();
	std::memcpy(f->prev_corners, f->t->ssm.curr_corners, sizeof(f->prev_corners));
	return 0;
}
int orc_pf_update(orc_pf *f, const double *normals, const double *uniforms){ return f->update(normals, uniforms); }
void orc_pf_get_corners(const orc_pf *f, double *out8){ std::memcpy(out8, f->t->ssm.curr_corners, 8 * sizeof(double)); }
void orc_pf_get_state(const orc_pf *f, double *outS){ for(int s = 0; s < f->S; ++s) outS[s] = f->t->ssm.curr_state[s]; }
int orc_pf_n_normals(const orc_pf *f){ return f->n_normals; }
int orc_pf_get_particles(const orc_pf *f, double *states, double *wts, double *cum, int *resampled){
	if(states) std::memcpy(states, f->states[f->cur].data(), (size_t)f->n*f->S*sizeof(double));
	if(wts) std::memcpy(wts, f->wts.data(), f->n*sizeof(double));
	if(cum) std::memcpy(cum, f->cum.data(), f->n*sizeof(double));
	if(resampled) *resampled = f->resampled;
	return f->max_wt_id;
}

} // extern "C"
