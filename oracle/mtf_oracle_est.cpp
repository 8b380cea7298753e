/* TEST INFRASTRUCTURE ONLY (see mtf_oracle.h): CPU restatement of the reference's robust warp estimation from point pairs,
 * the step GridTracker::update runs after its cells (SM/src/GridTracker.cc:269, ssm.estimateWarpFromPts):
 *   SSM/src/SSMEstimator.cc       runRANSAC :73-139, runLMeDS :143-217, getSubset :220-259, checkSubset :262-296,
 *                                 findInliers :35-46, cvRANSACUpdateNumIters :49-71, LevMarq :298-518
 *   SSM/src/HomographyEstimator.cc runKernel :16-78, computeReprojError :81-95, refine :97-145, estimateHomography :168-228
 *   SSM/src/AffineEstimator.cc     runKernel :17-46, computeReprojError :49-62, refine :64-106, estimateAffine :126-185
 *   SSM/src/Homography.cc:885-897, Affine.cc:359-369   warp matrix -> state update
 *
 * Third-party pieces the reference calls and this file restates from their published definitions (OpenCV 2.4 C API, absent from
 * /root/reference): cvRNG / cvRandInt (multiply-with-carry, state = (u32)state * 4164903690 + (state >> 32)), cvRound (round half
 * to even), cvEigenVV on a symmetric matrix (any convergent symmetric eigen-solver; here cyclic Jacobi, eigenvalues descending),
 * cvSVD + cvSVBkSb on the symmetric damped normal matrix (pseudo-inverse with threshold 2 eps sum(w)), Eigen's JacobiSVD least
 * squares (utils::computeAffineDLT, warpUtils.cc:344-377; here Householder QR of the same matrix).
 *
 * Parity status: the reference seeds its generator from boost::random_device (SSMEstimator.cc:22-24), so it has no reproducible
 * stream of its own and holds no golden vectors for this path; the restatement is cross-checked against OpenCV 4.13's
 * findHomography / estimateAffine2D (tests/test_estimator.py: least squares + LM on noisy data, inlier masks on data with a clear
 * margin), against NumPy (eigenvector, least squares) and by its own invariants.  "parity unpinned" for the RANSAC stream. */
#include "mtf_oracle.h"
#include <cmath>
#include <cfloat>
#include <cstring>
#include <cstdint>
#include <vector>
#include <algorithm>

namespace {

struct CvRng {
	uint64_t s;
	explicit CvRng(uint64_t seed) : s(seed ? seed : (uint64_t)(int64_t)-1) {}
	unsigned next() { s = (uint64_t)(unsigned)s * 4164903690U + (s >> 32); return (unsigned)s; }
};

inline int cv_round(double v) { return (int)std::nearbyint(v); }

/* cyclic Jacobi for a symmetric n x n matrix: w descending, row i of V = eigenvector i */
void jacobi_eigen(int n, double *A, double *V, double *w) {
	for(int i = 0; i < n; ++i) for(int j = 0; j < n; ++j) V[i * n + j] = i == j;
	for(int sweep = 0; sweep < 60; ++sweep) {
		double off = 0, diag = 0;
		for(int i = 0; i < n; ++i) { diag += A[i * n + i] * A[i * n + i]; for(int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j]; }
		if(off <= 1e-300 || off <= 1e-34 * diag) break;
		for(int p = 0; p < n - 1; ++p) for(int q = p + 1; q < n; ++q) {
			double apq = A[p * n + q];
			if(apq == 0) continue;
			double theta = (A[q * n + q] - A[p * n + p]) / (2 * apq);
			double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
			double c = 1 / std::sqrt(t * t + 1), s = t * c;
			for(int k = 0; k < n; ++k) {
				double akp = A[k * n + p], akq = A[k * n + q];
				A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq;
			}
			for(int k = 0; k < n; ++k) {
				double apk = A[p * n + k], aqk = A[q * n + k];
				A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk;
			}
			for(int k = 0; k < n; ++k) {
				double vpk = V[p * n + k], vqk = V[q * n + k];
				V[p * n + k] = c * vpk - s * vqk; V[q * n + k] = s * vpk + c * vqk;
			}
		}
	}
	std::vector<int> idx(n);
	for(int i = 0; i < n; ++i) idx[i] = i;
	std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return A[a * n + a] > A[b * n + b]; });
	std::vector<double> Vs(n * n);
	for(int i = 0; i < n; ++i) { w[i] = A[idx[i] * n + idx[i]]; for(int k = 0; k < n; ++k) Vs[i * n + k] = V[idx[i] * n + k]; }
	std::memcpy(V, Vs.data(), sizeof(double) * n * n);
}

struct Model {
	int n_params;       /* 8: 3x3 with H[8] = 1; 6: 2x3 */
	int model_points;
	bool homography;
};

/* HomographyEstimator::runKernel (HomographyEstimator.cc:16-78): normalised DLT, eigenvector of the smallest eigenvalue of LtL */
int hom_run_kernel(const double *M, const double *m, int count, double *H) {
	double cMx = 0, cMy = 0, cmx = 0, cmy = 0, sMx = 0, sMy = 0, smx = 0, smy = 0;
	for(int i = 0; i < count; ++i) { cmx += m[2 * i]; cmy += m[2 * i + 1]; cMx += M[2 * i]; cMy += M[2 * i + 1]; }
	cmx /= count; cmy /= count; cMx /= count; cMy /= count;
	for(int i = 0; i < count; ++i) {
		smx += std::fabs(m[2 * i] - cmx); smy += std::fabs(m[2 * i + 1] - cmy);
		sMx += std::fabs(M[2 * i] - cMx); sMy += std::fabs(M[2 * i + 1] - cMy);
	}
	if(std::fabs(smx) < DBL_EPSILON || std::fabs(smy) < DBL_EPSILON || std::fabs(sMx) < DBL_EPSILON || std::fabs(sMy) < DBL_EPSILON) return 0;
	smx = count / smx; smy = count / smy; sMx = count / sMx; sMy = count / sMy;
	const double invHnorm[9] = { 1. / smx, 0, cmx, 0, 1. / smy, cmy, 0, 0, 1 };
	const double Hnorm2[9] = { sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1 };
	double LtL[81] = { 0 };
	for(int i = 0; i < count; ++i) {
		double x = (m[2 * i] - cmx) * smx, y = (m[2 * i + 1] - cmy) * smy;
		double X = (M[2 * i] - cMx) * sMx, Y = (M[2 * i + 1] - cMy) * sMy;
		double Lx[9] = { X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x };
		double Ly[9] = { 0, 0, 0, X, Y, 1, -y * X, -y * Y, -y };
		for(int j = 0; j < 9; ++j) for(int k = j; k < 9; ++k) LtL[j * 9 + k] += Lx[j] * Lx[k] + Ly[j] * Ly[k];
	}
	for(int j = 0; j < 9; ++j) for(int k = 0; k < j; ++k) LtL[j * 9 + k] = LtL[k * 9 + j];
	double V[81], W[9];
	jacobi_eigen(9, LtL, V, W);
	const double *H0 = V + 72;
	double T[9], R[9];
	for(int r = 0; r < 3; ++r) for(int c = 0; c < 3; ++c) { double s = 0; for(int k = 0; k < 3; ++k) s += invHnorm[r * 3 + k] * H0[k * 3 + c]; T[r * 3 + c] = s; }
	for(int r = 0; r < 3; ++r) for(int c = 0; c < 3; ++c) { double s = 0; for(int k = 0; k < 3; ++k) s += T[r * 3 + k] * Hnorm2[k * 3 + c]; R[r * 3 + c] = s; }
	const double sc = 1. / R[8];
	for(int i = 0; i < 9; ++i) H[i] = R[i] * sc;
	return 1;
}

/* AffineEstimator::runKernel (AffineEstimator.cc:17-46) = utils::computeAffineDLT (warpUtils.cc:344-377): least squares of
 * [X Y 1] a = x, [X Y 1] b = y.  The reference takes the SVD of the 2n x 6 block matrix; the two blocks are the same n x 3
 * matrix, solved here by Householder QR. */
int aff_run_kernel(const double *M, const double *m, int count, double *H) {
	std::vector<double> B(count * 3), rx(count), ry(count);
	for(int i = 0; i < count; ++i) { B[i * 3] = M[2 * i]; B[i * 3 + 1] = M[2 * i + 1]; B[i * 3 + 2] = 1; rx[i] = m[2 * i]; ry[i] = m[2 * i + 1]; }
	for(int k = 0; k < 3; ++k) {
		double nrm = 0;
		for(int i = k; i < count; ++i) nrm += B[i * 3 + k] * B[i * 3 + k];
		nrm = std::sqrt(nrm);
		if(nrm == 0) return 0;
		const double alpha = B[k * 3 + k] > 0 ? -nrm : nrm;
		std::vector<double> v(count, 0.0);
		for(int i = k; i < count; ++i) v[i] = B[i * 3 + k];
		v[k] -= alpha;
		double vv = 0;
		for(int i = k; i < count; ++i) vv += v[i] * v[i];
		if(vv == 0) continue;
		auto reflect = [&](double *col, int stride) {
			double d = 0;
			for(int i = k; i < count; ++i) d += v[i] * col[i * stride];
			d = 2 * d / vv;
			for(int i = k; i < count; ++i) col[i * stride] -= d * v[i];
		};
		for(int c = k; c < 3; ++c) reflect(B.data() + c, 3);
		reflect(rx.data(), 1); reflect(ry.data(), 1);
	}
	for(int pass = 0; pass < 2; ++pass) {
		double *r = pass ? ry.data() : rx.data(), x[3];
		for(int k = 2; k >= 0; --k) {
			double s = r[k];
			for(int c = k + 1; c < 3; ++c) s -= B[k * 3 + c] * x[c];
			x[k] = s / B[k * 3 + k];
		}
		H[pass * 3] = x[0]; H[pass * 3 + 1] = x[1]; H[pass * 3 + 2] = x[2];
	}
	return 1;
}

void reproj_err(const Model &md, const double *M, const double *m, int n, const double *H, float *err) {
	for(int i = 0; i < n; ++i) {
		double dx, dy;
		if(md.homography) {
			double ww = 1. / (H[6] * M[2 * i] + H[7] * M[2 * i + 1] + 1.);
			dx = (H[0] * M[2 * i] + H[1] * M[2 * i + 1] + H[2]) * ww - m[2 * i];
			dy = (H[3] * M[2 * i] + H[4] * M[2 * i + 1] + H[5]) * ww - m[2 * i + 1];
		} else {
			dx = (H[0] * M[2 * i] + H[1] * M[2 * i + 1] + H[2]) - m[2 * i];
			dy = (H[3] * M[2 * i] + H[4] * M[2 * i + 1] + H[5]) - m[2 * i + 1];
		}
		err[i] = (float)(dx * dx + dy * dy);
	}
}

int run_kernel(const Model &md, const double *M, const double *m, int n, double *H) {
	return md.homography ? hom_run_kernel(M, m, n, H) : aff_run_kernel(M, m, n, H);
}

/* SSMEstimator::findInliers (:35-46) */
int find_inliers(const Model &md, const double *M, const double *m, int n, const double *H, float *err, unsigned char *mask, double threshold) {
	reproj_err(md, M, m, n, H, err);
	threshold *= threshold;
	int good = 0;
	for(int i = 0; i < n; ++i) good += mask[i] = err[i] <= threshold;
	return good;
}

/* cvRANSACUpdateNumIters as restated in SSMEstimator.cc:49-71 */
int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
	p = std::max(p, 0.); p = std::min(p, 1.);
	ep = std::max(ep, 0.); ep = std::min(ep, 1.);
	double num = std::max(1. - p, DBL_MIN);
	double denom = 1. - std::pow(1. - ep, model_points);
	if(denom < DBL_MIN) return 0;
	num = std::log(num); denom = std::log(denom);
	return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : cv_round(num / denom);
}

/* SSMEstimator::checkSubset (:262-296) for checkPartialSubsets = false (both estimators clear it) */
bool check_subset(const double *p, int count) {
	if(count <= 2) return true;
	int i, j, k;
	for(i = 0; i <= count - 1; ++i) {
		for(j = 0; j < i; ++j) {
			double dx1 = p[2 * j] - p[2 * i], dy1 = p[2 * j + 1] - p[2 * i + 1];
			for(k = 0; k < j; ++k) {
				double dx2 = p[2 * k] - p[2 * i], dy2 = p[2 * k + 1] - p[2 * i + 1];
				if(std::fabs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (std::fabs(dx1) + std::fabs(dy1) + std::fabs(dx2) + std::fabs(dy2))) break;
			}
			if(k < j) break;
		}
		if(j < i) break;
	}
	return i > count - 1;
}

/* SSMEstimator::getSubset (:220-259), cvRandInt stream, checkPartialSubsets = false */
bool get_subset(CvRng &rng, const double *m1, const double *m2, int count, int mp, double *ms1, double *ms2, int max_attempts, int *idx_out) {
	std::vector<int> idx(mp);
	int i = 0, j, iters = 0;
	for(; iters < max_attempts; ++iters) {
		for(i = 0; i < mp && iters < max_attempts;) {
			int idx_i = idx[i] = (int)(rng.next() % (unsigned)count);
			for(j = 0; j < i; ++j) if(idx_i == idx[j]) break;
			if(j < i) continue;
			ms1[2 * i] = m1[2 * idx_i]; ms1[2 * i + 1] = m1[2 * idx_i + 1];
			ms2[2 * i] = m2[2 * idx_i]; ms2[2 * i + 1] = m2[2 * idx_i + 1];
			++i;
		}
		if(i == mp && (!check_subset(ms1, i) || !check_subset(ms2, i))) continue;
		break;
	}
	if(idx_out) for(int k = 0; k < mp; ++k) idx_out[k] = idx[k];
	return i == mp && iters < max_attempts;
}

/* LevMarq::updateAlt / step (SSMEstimator.cc:436-518) with err.empty(), completeSymmFlag = false, all parameters free */
struct LevMarq {
	enum { DONE = 0, STARTED = 1, CALC_J = 2, CHECK_ERR = 3 };
	int n, state, iters, max_iter, lambdaLg10;
	double epsilon, prevErrNorm, errNorm;
	std::vector<double> param, prevParam, JtJ, JtErr;
	LevMarq(int nparams, int maxIters) : n(nparams), state(STARTED), iters(0), lambdaLg10(-3), epsilon(DBL_EPSILON), prevErrNorm(DBL_MAX), errNorm(0),
		param(nparams), prevParam(nparams), JtJ(nparams * nparams), JtErr(nparams) {
		max_iter = std::min(std::max(maxIters, 1), 1000);
	}
	void step() {
		const double lambda = std::exp(lambdaLg10 * std::log(10.));
		/* cvCompleteSymm(JtJ, false): the upper half is copied to the lower */
		for(int i = 0; i < n; ++i) for(int j = 0; j < i; ++j) JtJ[i * n + j] = JtJ[j * n + i];
		std::vector<double> A(JtJ), V(n * n), w(n), x(n, 0.0);
		for(int i = 0; i < n; ++i) A[i * n + i] *= 1. + lambda;
		/* cvSVD + cvSVBkSb of a symmetric matrix: A = V' diag(w) V, singular values |w|, threshold 2 eps sum|w| */
		jacobi_eigen(n, A.data(), V.data(), w.data());
		double thr = 0;
		for(int i = 0; i < n; ++i) thr += std::fabs(w[i]);
		thr *= 2 * DBL_EPSILON;
		for(int i = 0; i < n; ++i) {
			if(std::fabs(w[i]) <= thr) continue;
			double d = 0;
			for(int k = 0; k < n; ++k) d += V[i * n + k] * JtErr[k];
			d /= w[i];
			for(int k = 0; k < n; ++k) x[k] += d * V[i * n + k];
		}
		for(int i = 0; i < n; ++i) param[i] = prevParam[i] - x[i];
	}
	/* returns false when finished; on true the caller accumulates JtJ / JtErr if want_J and errNorm if want_err (the
	 * reference hands out the pointers it wants filled: after an accepted step only JtJ / JtErr, errNorm keeps its value) */
	bool update_alt(bool &want_J, bool &want_err) {
		want_J = want_err = false;
		if(state == DONE) return false;
		if(state == STARTED) {
			std::fill(JtJ.begin(), JtJ.end(), 0.0); std::fill(JtErr.begin(), JtErr.end(), 0.0);
			errNorm = 0; want_J = want_err = true; state = CALC_J; return true;
		}
		if(state == CALC_J) {
			prevParam = param; step();
			prevErrNorm = errNorm; errNorm = 0; want_err = true; state = CHECK_ERR; return true;
		}
		if(errNorm > prevErrNorm) {
			if(++lambdaLg10 <= 16) { step(); errNorm = 0; want_err = true; state = CHECK_ERR; return true; }
		}
		lambdaLg10 = std::max(lambdaLg10 - 1, -16);
		double dn = 0, pn = 0;
		for(int i = 0; i < n; ++i) { dn += (param[i] - prevParam[i]) * (param[i] - prevParam[i]); pn += prevParam[i] * prevParam[i]; }
		/* cvNorm(param, prevParam, CV_RELATIVE_L2) = |param - prevParam| / |prevParam| */
		const double change = std::sqrt(dn) / (std::sqrt(pn) + DBL_EPSILON);
		if(++iters >= max_iter || change < epsilon) { state = DONE; return false; }
		prevErrNorm = errNorm;
		std::fill(JtJ.begin(), JtJ.end(), 0.0); std::fill(JtErr.begin(), JtErr.end(), 0.0);
		want_J = true; state = CALC_J; return true;
	}
};

/* HomographyEstimator::refine (:97-145) / AffineEstimator::refine (:64-106) */
int refine(const Model &md, const double *M, const double *m, int count, double *H, int maxIters) {
	const int n = md.n_params;
	LevMarq solver(n, maxIters);
	for(int i = 0; i < n; ++i) solver.param[i] = H[i];
	int evals = 0;
	for(;;) {
		bool want_J, want_err;
		if(!solver.update_alt(want_J, want_err)) break;
		++evals;
		const double *h = solver.param.data();
		for(int i = 0; i < count; ++i) {
			const double Mx = M[2 * i], My = M[2 * i + 1];
			double J[2][8] = { { 0 } }, err[2];
			if(md.homography) {
				double ww = h[6] * Mx + h[7] * My + 1.;
				ww = std::fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
				const double xi = (h[0] * Mx + h[1] * My + h[2]) * ww, yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
				err[0] = xi - m[2 * i]; err[1] = yi - m[2 * i + 1];
				J[0][0] = Mx * ww; J[0][1] = My * ww; J[0][2] = ww; J[0][6] = -Mx * ww * xi; J[0][7] = -My * ww * xi;
				J[1][3] = Mx * ww; J[1][4] = My * ww; J[1][5] = ww; J[1][6] = -Mx * ww * yi; J[1][7] = -My * ww * yi;
			} else {
				const double xi = h[0] * Mx + h[1] * My + h[2], yi = h[3] * Mx + h[4] * My + h[5];
				err[0] = xi - m[2 * i]; err[1] = yi - m[2 * i + 1];
				J[0][0] = Mx; J[0][1] = My; J[0][2] = 1; J[1][3] = Mx; J[1][4] = My; J[1][5] = 1;
			}
			if(want_J) {
				for(int j = 0; j < n; ++j) {
					for(int k = j; k < n; ++k) solver.JtJ[j * n + k] += J[0][j] * J[0][k] + J[1][j] * J[1][k];
					solver.JtErr[j] += J[0][j] * err[0] + J[1][j] * err[1];
				}
			}
			if(want_err) solver.errNorm += err[0] * err[0] + err[1] * err[1];
		}
	}
	for(int i = 0; i < n; ++i) H[i] = solver.param[i];
	return evals;
}

/* SSMEstimator::runRANSAC (:73-139) */
bool run_ransac(const Model &md, CvRng &rng, const double *m1, const double *m2, int count, double *model, unsigned char *mask0,
	double thresh, double confidence, int maxIters, int maxAttempts, int *iters_out) {
	const int mp = md.model_points;
	if(count < mp) return false;
	std::vector<unsigned char> mask(mask0, mask0 + count), tmask(count);
	std::vector<float> err(count);
	std::vector<double> ms1(2 * mp), ms2(2 * mp), models(9);
	int iter, niters = maxIters, maxGood = 0;
	if(count <= mp) {
		niters = 1;
		std::memcpy(ms1.data(), m1, sizeof(double) * 2 * count); std::memcpy(ms2.data(), m2, sizeof(double) * 2 * count);
	}
	for(iter = 0; iter < niters; ++iter) {
		if(count > mp) {
			bool found = get_subset(rng, m1, m2, count, mp, ms1.data(), ms2.data(), maxAttempts, nullptr);
			if(!found) { if(iter == 0) { *iters_out = 0; return false; } break; }
		}
		if(run_kernel(md, ms1.data(), ms2.data(), mp, models.data()) <= 0) continue;
		int good = find_inliers(md, m1, m2, count, models.data(), err.data(), tmask.data(), thresh);
		if(good > std::max(maxGood, mp - 1)) {
			std::swap(tmask, mask);
			std::memcpy(model, models.data(), sizeof(double) * (md.homography ? 9 : 6));
			maxGood = good;
			niters = ransac_update_num_iters(confidence, (double)(count - good) / count, mp, niters);
		}
	}
	*iters_out = iter;
	if(maxGood > 0) { std::memcpy(mask0, mask.data(), count); return true; }
	return false;
}

/* SSMEstimator::runLMeDS (:143-217) */
bool run_lmeds(const Model &md, CvRng &rng, const double *m1, const double *m2, int count, double *model, unsigned char *mask,
	double confidence, int maxIters, int maxAttempts, int *iters_out) {
	const double outlierRatio = 0.45;
	const int mp = md.model_points;
	if(count < mp) return false;
	std::vector<float> err(count);
	std::vector<double> ms1(2 * mp), ms2(2 * mp), models(9);
	int iter, niters;
	double minMedian = DBL_MAX;
	if(count <= mp) { std::memcpy(ms1.data(), m1, sizeof(double) * 2 * count); std::memcpy(ms2.data(), m2, sizeof(double) * 2 * count); }
	niters = cv_round(std::log(1 - confidence) / std::log(1 - std::pow(1 - outlierRatio, (double)mp)));
	niters = std::min(std::max(niters, 3), maxIters);
	for(iter = 0; iter < niters; ++iter) {
		if(count > mp) {
			bool found = get_subset(rng, m1, m2, count, mp, ms1.data(), ms2.data(), maxAttempts, nullptr);
			if(!found) { if(iter == 0) { *iters_out = 0; return false; } break; }
		}
		if(run_kernel(md, ms1.data(), ms2.data(), mp, models.data()) <= 0) continue;
		reproj_err(md, m1, m2, count, models.data(), err.data());
		/* icvSortDistances sorts the floats' bit patterns as ints: the same order for non-negative values */
		std::sort(reinterpret_cast<int *>(err.data()), reinterpret_cast<int *>(err.data()) + count);
		double median = count % 2 != 0 ? err[count / 2] : (err[count / 2 - 1] + err[count / 2]) * 0.5;
		if(median < minMedian) { minMedian = median; std::memcpy(model, models.data(), sizeof(double) * (md.homography ? 9 : 6)); }
	}
	*iters_out = iter;
	if(minMedian < DBL_MAX) {
		double sigma = 2.5 * 1.4826 * (1 + 5. / (count - mp)) * std::sqrt(minMedian);
		sigma = std::max(sigma, 0.001);
		int good = find_inliers(md, m1, m2, count, model, err.data(), mask, sigma);
		return good >= mp;
	}
	return false;
}

} // namespace

extern "C" {

void orc_est_default_params(orc_est_params *p) {
	/* SSMEstimatorParams.cc:5-13 */
	p->method = ORC_EST_RANSAC; p->ransac_reproj_thresh = 10.0; p->n_model_pts = 4; p->refine = 1; p->max_iters = 2000;
	p->max_subset_attempts = 300; p->confidence = 0.995; p->lm_max_iters = 10; p->seed = 0;
}

unsigned orc_cv_rand_int(unsigned long long *state) {
	CvRng r(*state); unsigned v = r.next(); *state = r.s; return v;
}

void orc_sym_eigen(const double *A, int n, double *V, double *w) {
	std::vector<double> a(A, A + n * n);
	jacobi_eigen(n, a.data(), V, w);
}

/* estimateHomography (HomographyEstimator.cc:168-228) / estimateAffine (AffineEstimator.cc:126-185) followed by
 * Homography::estimateWarpFromPts (Homography.cc:885-897) / Affine::estimateWarpFromPts (Affine.cc:359-369).
 * warp: 9 doubles (3x3; affine: third row 0 0 1); zeros when the estimation fails.  info: [0] result, [1] RANSAC / LMedS
 * hypotheses drawn, [2] inliers, [3] LM evaluations. */
int orc_estimate_warp(int ssm, const float *in_pts, const float *out_pts, int n, const orc_est_params *ep,
	double *warp, unsigned char *mask, double *state_update, int *info) {
	Model md;
	md.homography = ssm == ORC_SSM_HOMOGRAPHY;
	md.n_params = md.homography ? 8 : 6;
	md.model_points = ep->n_model_pts;
	if(ssm != ORC_SSM_HOMOGRAPHY && ssm != ORC_SSM_AFFINE) return -1;
	if(n < ep->n_model_pts) return -1;
	double thresh = ep->ransac_reproj_thresh;
	if(thresh <= 0) thresh = 3;                       /* SSMEstimatorParams.cc:55-57 */
	std::vector<double> M(2 * n), m(2 * n);
	for(int i = 0; i < 2 * n; ++i) { M[i] = in_pts[i]; m[i] = out_pts[i]; }
	std::vector<unsigned char> tmask(n, 1);
	double H[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 1 };
	CvRng rng(ep->seed);
	int drawn = 0, evals = 0;
	bool result;
	const int method = n == ep->n_model_pts ? ORC_EST_LEAST_SQUARES : ep->method;
	if(method == ORC_EST_LMEDS) result = run_lmeds(md, rng, M.data(), m.data(), n, H, tmask.data(), ep->confidence, ep->max_iters, ep->max_subset_attempts, &drawn);
	else if(method == ORC_EST_RANSAC) result = run_ransac(md, rng, M.data(), m.data(), n, H, tmask.data(), thresh, ep->confidence, ep->max_iters, ep->max_subset_attempts, &drawn);
	else result = run_kernel(md, M.data(), m.data(), n, H) > 0;
	int n_in = n;
	if(result && n > ep->n_model_pts) {
		int q = 0;
		for(int i = 0; i < n; ++i) if(tmask[i]) { M[2 * q] = M[2 * i]; M[2 * q + 1] = M[2 * i + 1]; m[2 * q] = m[2 * i]; m[2 * q + 1] = m[2 * i + 1]; ++q; }
		n_in = q;
		if(method == ORC_EST_RANSAC) run_kernel(md, M.data(), m.data(), n_in, H);
		if(ep->refine) evals = refine(md, M.data(), m.data(), n_in, H, ep->lm_max_iters);
	}
	if(!md.homography) { H[6] = 0; H[7] = 0; H[8] = 1; }
	if(!result) for(int i = 0; i < 9; ++i) H[i] = 0;
	if(warp) std::memcpy(warp, H, sizeof(H));
	if(mask) std::memcpy(mask, tmask.data(), n);
	if(state_update) {
		if(md.homography) {
			state_update[0] = H[0] - 1; state_update[1] = H[1]; state_update[2] = H[2]; state_update[3] = H[3];
			state_update[4] = H[4] - 1; state_update[5] = H[5]; state_update[6] = H[6]; state_update[7] = H[7];
		} else {
			state_update[0] = H[2]; state_update[1] = H[5]; state_update[2] = H[0] - 1; state_update[3] = H[1];
			state_update[4] = H[3]; state_update[5] = H[4] - 1;
		}
	}
	if(info) { info[0] = result; info[1] = drawn; info[2] = n_in; info[3] = evals; }
	return result ? 1 : 0;
}

/* the subsets the stream of `seed` produces for these points (n_subsets x model_points indices; -1 rows after a failure) */
int orc_est_subsets(const float *in_pts, const float *out_pts, int n, int model_points, int max_attempts, unsigned long long seed,
	int n_subsets, int *idx) {
	std::vector<double> M(2 * n), m(2 * n), a(2 * model_points), b(2 * model_points);
	for(int i = 0; i < 2 * n; ++i) { M[i] = in_pts[i]; m[i] = out_pts[i]; }
	CvRng rng(seed);
	int k = 0;
	for(; k < n_subsets; ++k) if(!get_subset(rng, M.data(), m.data(), n, model_points, a.data(), b.data(), max_attempts, idx + k * model_points)) break;
	for(int r = k; r < n_subsets; ++r) for(int c = 0; c < model_points; ++c) idx[r * model_points + c] = -1;
	return k;
}

}
