// grid_estimator.cu -- SSM::estimateWarpFromPts for Homography / Affine on the device: the RANSAC / LMedS / least-squares
// estimation of one warp from the P cell centroids that GridTracker::update runs once per frame
// (SM/src/GridTracker.cc:253-269 -> SSM/src/Homography.cc:885-897, Affine.cc:359-369 -> estimateHomography
// HomographyEstimator.cc:168-228 / estimateAffine AffineEstimator.cc:126-185 -> SSMEstimator.cc runRANSAC :73-139, runLMeDS
// :143-217, getSubset :220-259, LevMarq :298-518).
//
// One CTA of eight warps.  The reference's loop "draw a subset, fit, count" is sequential only through its generator and its
// best-so-far bookkeeping: thread 0 draws a round of 32 subsets from the cvRNG stream (checkSubset included), the warps fit
// and score them in parallel (one hypothesis per warp at a time: lane 0 fits, all lanes score the P points), thread 0 replays
// the acceptance rule over the round in the reference's order, so the accepted model, the adaptive iteration count and the
// mask are those of the sequential loop.  The final fit on the inliers and the Levenberg-Marquardt refinement are block-wide
// reductions with thread 0 running the solver's state machine.
//
// Numerics: double precision throughout, the reprojection errors rounded to float as in the reference.  Where the reference
// calls cvEigenVV for the eigenvector of the smallest eigenvalue of LtL this file runs inverse iteration on an LDLt
// factorisation, and where it calls cvSVD + cvSVBkSb on the damped normal matrix a diagonally scaled elimination with partial
// pivoting: the same vector / solution to rounding, not the same rounding.
#include "grid_estimator.cuh"
#include <cfloat>
#include "lk_warp.cuh"

namespace mtfb {
namespace {

constexpr unsigned FULL = 0xffffffffu;
// -DEST_PROF: thread 0 accumulates clock64 deltas of the kernel's phases into out[21..31] (profiles/experiments)
#ifdef EST_PROF
#define PROF_T0() const long long prof_t0 = clock64()
#define PROF_ADD(slot) do{ if(threadIdx.x == 0) g_prof[slot] += (double)(clock64() - prof_t0); }while(0)
__device__ double g_prof[11];
#else
#define PROF_T0() do{}while(0)
#define PROF_ADD(slot) do{}while(0)
#endif
constexpr int BATCH = 32;                 // hypotheses per round
constexpr int EST_STAGED = 1024;          // points kept in shared memory
enum { M_RANSAC = 0, M_LMEDS = 1, M_LS = 2 };   // SSMEstimatorParams::EstType (SSMEstimatorParams.h:11)

// cvRNG / cvRandInt (OpenCV core: multiply-with-carry)
struct CvRng {
	unsigned long long s;
	__device__ unsigned next(){ s = (unsigned long long)(unsigned)s * 4164903690ULL + (s >> 32); return (unsigned)s; }
};

// cvRANSACUpdateNumIters as SSMEstimator.cc:49-71 has it
__device__ int ransac_update_num_iters(double p, double ep, int model_points, int max_iters){
	p = fmax(p, 0.); p = fmin(p, 1.);
	ep = fmax(ep, 0.); ep = fmin(ep, 1.);
	double num = fmax(1. - p, DBL_MIN);
	double denom = 1. - pow(1. - ep, (double)model_points);
	if(denom < DBL_MIN) return 0;
	num = log(num); denom = log(denom);
	return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : __double2int_rn(num / denom);
}

// SSMEstimator::checkSubset (:262-296) with checkPartialSubsets = false
__device__ bool check_subset(const double *p, int count){
	if(count <= 2) return true;
	int i, j, k;
	for(i = 0; i <= count - 1; ++i){
		for(j = 0; j < i; ++j){
			const double dx1 = p[2 * j] - p[2 * i], dy1 = p[2 * j + 1] - p[2 * i + 1];
			for(k = 0; k < j; ++k){
				const double dx2 = p[2 * k] - p[2 * i], dy2 = p[2 * k + 1] - p[2 * i + 1];
				if(fabs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) break;
			}
			if(k < j) break;
		}
		if(j < i) break;
	}
	return i > count - 1;
}

// SSMEstimator::getSubset (:220-259)
__device__ bool get_subset(CvRng &rng, const float *m1, const float *m2, int count, int mp, int max_attempts, int *idx){
	double ms1[2 * EST_MAX_MODEL_PTS], ms2[2 * EST_MAX_MODEL_PTS];
	int i = 0, j, iters = 0;
	for(; iters < max_attempts; ++iters){
		for(i = 0; i < mp && iters < max_attempts;){
			const int idx_i = idx[i] = (int)(rng.next() % (unsigned)count);
			for(j = 0; j < i; ++j) if(idx_i == idx[j]) break;
			if(j < i) continue;
			ms1[2 * i] = m1[2 * idx_i]; ms1[2 * i + 1] = m1[2 * idx_i + 1];
			ms2[2 * i] = m2[2 * idx_i]; ms2[2 * i + 1] = m2[2 * idx_i + 1];
			++i;
		}
		if(i == mp && (!check_subset(ms1, i) || !check_subset(ms2, i))) continue;
		break;
	}
	return i == mp && iters < max_attempts;
}

// eigenvector of the smallest eigenvalue of a symmetric positive semi-definite 9 x 9 matrix A (shared memory, overwritten):
// inverse iteration with a zero shift on an LDLt factorisation, by one warp -- lane i holds row i of L, then also column i, so
// that both triangular solves are "broadcast the newest unknown, every lane updates its own partial sum".  x (shared, 9) is
// scaled to a largest component of +1.
__device__ void smallest_eigvec9_warp(double *A, double *x, int lane){
	const int row = lane < 9 ? lane : 8;
	double a[9], d[9];
#pragma unroll
	for(int k = 0; k < 9; ++k) a[k] = A[row * 9 + k];
	double tr = 0;
#pragma unroll
	for(int i = 0; i < 9; ++i) tr += A[i * 9 + i];
	const double tiny = tr * 1e-30 + 1e-300;
	__syncwarp();
#pragma unroll
	for(int j = 0; j < 9; ++j){
		double Lj[9];
#pragma unroll
		for(int k = 0; k < j; ++k) Lj[k] = __shfl_sync(FULL, a[k], j);
		double dj = __shfl_sync(FULL, a[j], j);
#pragma unroll
		for(int k = 0; k < j; ++k) dj -= Lj[k] * Lj[k] * d[k];
		if(fabs(dj) < tiny) dj = tiny;
		d[j] = dj;
		double sacc = a[j];
#pragma unroll
		for(int k = 0; k < j; ++k) sacc -= a[k] * Lj[k] * d[k];
		if(row > j) a[j] = sacc / dj;
	}
	// column `row` of L: lt[k] = L[k][row], k > row
#pragma unroll
	for(int k = 0; k < 9; ++k) if(lane < 9 && k < row) A[row * 9 + k] = a[k];
	__syncwarp();
	double lt[9];
#pragma unroll
	for(int k = 0; k < 9; ++k) lt[k] = k > row ? A[k * 9 + row] : 0.0;
	double xv[9], yv[9];
#pragma unroll
	for(int i = 0; i < 9; ++i) xv[i] = 1.0;
	double prev_diff = 1e300;
	for(int it = 0; it < 100; ++it){
		double acc = 1.0;
#pragma unroll
		for(int i = 0; i < 9; ++i) if(row == i) acc = xv[i];
#pragma unroll
		for(int i = 0; i < 9; ++i){
			yv[i] = __shfl_sync(FULL, acc, i);
			if(row > i) acc -= a[i] * yv[i];
		}
#pragma unroll
		for(int i = 0; i < 9; ++i){ yv[i] /= d[i]; if(row == i) acc = yv[i]; }
#pragma unroll
		for(int i = 8; i >= 0; --i){
			yv[i] = __shfl_sync(FULL, acc, i);
			if(row < i) acc -= lt[i] * yv[i];
		}
		double big = yv[0];
#pragma unroll
		for(int i = 1; i < 9; ++i) if(fabs(yv[i]) > fabs(big)) big = yv[i];
		const double sc = 1.0 / big;
		double diff = 0;
#pragma unroll
		for(int i = 0; i < 9; ++i){ const double v = yv[i] * sc; diff = fmax(diff, fabs(v - xv[i])); xv[i] = v; }
		// converged, or at the rounding floor of this matrix (the iterates stop contracting)
		if(diff < 1e-15 || (it >= 2 && diff < 1e-10 && diff > 0.25 * prev_diff)) break;
		prev_diff = diff;
	}
#pragma unroll
	for(int i = 0; i < 9; ++i) if(lane == i) x[i] = xv[i];
	__syncwarp();
}

// the tail of HomographyEstimator::runKernel (:70-76): H = invHnorm * H0 * Hnorm2, scaled to H[8] = 1
__device__ void denormalise(const double *H0, double cmx, double cmy, double smx, double smy, double cMx, double cMy, double sMx, double sMy, double *H){
	const double invHnorm[9] = { 1. / smx, 0, cmx, 0, 1. / smy, cmy, 0, 0, 1 };
	const double Hnorm2[9] = { sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1 };
	double T[9], R[9];
	for(int r = 0; r < 3; ++r) for(int c = 0; c < 3; ++c){ double s = 0; for(int k = 0; k < 3; ++k) s += invHnorm[r * 3 + k] * H0[k * 3 + c]; T[r * 3 + c] = s; }
	for(int r = 0; r < 3; ++r) for(int c = 0; c < 3; ++c){ double s = 0; for(int k = 0; k < 3; ++k) s += T[r * 3 + k] * Hnorm2[k * 3 + c]; R[r * 3 + c] = s; }
	const double sc = 1. / R[8];
	for(int i = 0; i < 9; ++i) H[i] = R[i] * sc;
}

// entry c of the row Lx (r = 0) / Ly (r = 1) of the DLT system for one normalised point pair (HomographyEstimator.cc:61-62)
__device__ __forceinline__ double l_entry(int r, int c, double x, double y, double X, double Y){
	const double t = r ? y : x;
	if(c >= 6) return c == 6 ? -t * X : (c == 7 ? -t * Y : -t);
	const int base = r ? 3 : 0;
	if(c < base || c >= base + 3) return 0.0;
	return c == base ? X : (c == base + 1 ? Y : 1.0);
}

struct WarpWs { double A[81], x[9], y[9], M[2 * EST_MAX_MODEL_PTS], m[2 * EST_MAX_MODEL_PTS]; };

// HomographyEstimator::runKernel (:16-78) for the few points of one hypothesis: lane 0 accumulates the normalisation and LtL with
// the sums in the reference's order, the warp finds the eigenvector, lane 0 undoes the normalisation
__device__ int hom_fit_subset(WarpWs &w, int count, double *H, int lane){
	if(lane == 0){
		const double *M = w.M, *m = w.m;
		double cMx = 0, cMy = 0, cmx = 0, cmy = 0, sMx = 0, sMy = 0, smx = 0, smy = 0;
		for(int i = 0; i < count; ++i){ cmx += m[2 * i]; cmy += m[2 * i + 1]; cMx += M[2 * i]; cMy += M[2 * i + 1]; }
		cmx /= count; cmy /= count; cMx /= count; cMy /= count;
		for(int i = 0; i < count; ++i){
			smx += fabs(m[2 * i] - cmx); smy += fabs(m[2 * i + 1] - cmy);
			sMx += fabs(M[2 * i] - cMx); sMy += fabs(M[2 * i + 1] - cMy);
		}
		int ok = 1;
		if(fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON) ok = 0;
		if(ok){ smx = count / smx; smy = count / smy; sMx = count / sMx; sMy = count / sMy; }
		w.y[0] = cmx; w.y[1] = cmy; w.y[2] = smx; w.y[3] = smy; w.y[4] = cMx; w.y[5] = cMy; w.y[6] = sMx; w.y[7] = sMy; w.y[8] = ok;
	}
	__syncwarp();
	if(w.y[8] == 0) return 0;
	{
		// LtL (HomographyEstimator.cc:59-68): entry (j, k) by one lane, the points in the reference's order
		const double cmx = w.y[0], cmy = w.y[1], smx = w.y[2], smy = w.y[3], cMx = w.y[4], cMy = w.y[5], sMx = w.y[6], sMy = w.y[7];
		for(int q = lane; q < 81; q += 32){
			const int j = q / 9, k = q % 9;
			double acc = 0;
			for(int i = 0; i < count; ++i){
				const double x = (w.m[2 * i] - cmx) * smx, y = (w.m[2 * i + 1] - cmy) * smy;
				const double X = (w.M[2 * i] - cMx) * sMx, Y = (w.M[2 * i + 1] - cMy) * sMy;
				acc += l_entry(0, j, x, y, X, Y) * l_entry(0, k, x, y, X, Y) + l_entry(1, j, x, y, X, Y) * l_entry(1, k, x, y, X, Y);
			}
			w.A[q] = acc;
		}
		__syncwarp();
	}
	smallest_eigvec9_warp(w.A, w.x, lane);
	if(lane == 0) denormalise(w.x, w.y[0], w.y[1], w.y[2], w.y[3], w.y[4], w.y[5], w.y[6], w.y[7], H);
	__syncwarp();
	return 1;
}

// least squares of [X Y 1] a = x, [X Y 1] b = y from the centred second moments (utils::computeAffineDLT, warpUtils.cc:344-377,
// takes the SVD of the block matrix; the solution is the same)
__device__ int aff_from_moments(double n, double cX, double cY, double cx, double cy, double Sxx, double Sxy, double Syy,
	double SXx, double SYx, double SXy, double SYy, double *H){
	const double det = Sxx * Syy - Sxy * Sxy;
	if(!(fabs(det) > 0)) return 0;
	const double a0 = (SXx * Syy - SYx * Sxy) / det, a1 = (SYx * Sxx - SXx * Sxy) / det;
	const double b0 = (SXy * Syy - SYy * Sxy) / det, b1 = (SYy * Sxx - SXy * Sxy) / det;
	H[0] = a0; H[1] = a1; H[2] = cx - a0 * cX - a1 * cY;
	H[3] = b0; H[4] = b1; H[5] = cy - b0 * cX - b1 * cY;
	H[6] = 0; H[7] = 0; H[8] = 1;
	return 1;
}
__device__ int aff_fit_subset(const WarpWs &w, int count, double *H){
	const double *M = w.M, *m = w.m;
	double cX = 0, cY = 0, cx = 0, cy = 0;
	for(int i = 0; i < count; ++i){ cX += M[2 * i]; cY += M[2 * i + 1]; cx += m[2 * i]; cy += m[2 * i + 1]; }
	cX /= count; cY /= count; cx /= count; cy /= count;
	double Sxx = 0, Sxy = 0, Syy = 0, SXx = 0, SYx = 0, SXy = 0, SYy = 0;
	for(int i = 0; i < count; ++i){
		const double X = M[2 * i] - cX, Y = M[2 * i + 1] - cY, x = m[2 * i] - cx, y = m[2 * i + 1] - cy;
		Sxx += X * X; Sxy += X * Y; Syy += Y * Y; SXx += X * x; SYx += Y * x; SXy += X * y; SYy += Y * y;
	}
	return aff_from_moments(count, cX, cY, cx, cy, Sxx, Sxy, Syy, SXx, SYx, SXy, SYy, H);
}

// HomographyEstimator / AffineEstimator::computeReprojError (:81-95 / :49-62) for one point
template<bool HOM> __device__ __forceinline__ float reproj_err(const double *H, double Mx, double My, double mx, double my){
	double dx, dy;
	if(HOM){
		const double ww = 1. / (H[6] * Mx + H[7] * My + 1.);
		dx = (H[0] * Mx + H[1] * My + H[2]) * ww - mx;
		dy = (H[3] * Mx + H[4] * My + H[5]) * ww - my;
	} else {
		dx = (H[0] * Mx + H[1] * My + H[2]) - mx;
		dy = (H[3] * Mx + H[4] * My + H[5]) - my;
	}
	return (float)(dx * dx + dy * dy);
}

// k-th smallest (0-based) of n non-negative floats given as their bit patterns: radix select by one warp
__device__ unsigned warp_select(const unsigned *v, int n, int k, unsigned *hist, int lane){
	unsigned prefix = 0, mask_bits = 0;
	int kk = k;
	for(int shift = 24; shift >= 0; shift -= 8){
		for(int i = lane; i < 256; i += 32) hist[i] = 0;
		__syncwarp();
		for(int i = lane; i < n; i += 32){ const unsigned x = v[i]; if((x & mask_bits) == prefix) atomicAdd(&hist[(x >> shift) & 255u], 1u); }
		__syncwarp();
		unsigned c[8], tot = 0;
#pragma unroll
		for(int j = 0; j < 8; ++j){ c[j] = hist[8 * lane + j]; tot += c[j]; }
		unsigned incl = tot;
#pragma unroll
		for(int off = 1; off < 32; off <<= 1){ const unsigned t = __shfl_up_sync(FULL, incl, off); if(lane >= off) incl += t; }
		const unsigned excl = incl - tot;
		const bool mine = (unsigned)kk >= excl && (unsigned)kk < incl;
		const int src = __ffs(__ballot_sync(FULL, mine)) - 1;
		int bin = 0, newk = 0;
		if(mine){
			unsigned run = excl;
			bool found = false;
#pragma unroll
			for(int j = 0; j < 8; ++j){
				if(!found && (unsigned)kk < run + c[j]){ bin = 8 * lane + j; newk = kk - (int)run; found = true; }
				run += c[j];
			}
		}
		bin = __shfl_sync(FULL, bin, src); newk = __shfl_sync(FULL, newk, src);
		prefix |= (unsigned)bin << shift; mask_bits |= 255u << shift; kk = newk;
		__syncwarp();
	}
	return prefix;
}

// sum of each of P2 (a power of two <= 32) values over the lanes of a warp in P2 - 1 + (5 - log2 P2) exchanges instead of 5 P2:
// at every stage a lane keeps one half of its values and trades the other half with its partner.  v[0] ends up holding the
// total of element (lane >> (5 - log2 P2)).
template<int P2> __device__ __forceinline__ void warp_transpose_sum(double *v, int lane){
	int off = 16;
#pragma unroll
	for(int h = P2 / 2; h >= 1; h >>= 1){
		const bool up = (lane & off) != 0;
#pragma unroll
		for(int j = 0; j < h; ++j){
			const double send = up ? v[j] : v[j + h];
			const double keep = up ? v[j + h] : v[j];
			v[j] = keep + __shfl_xor_sync(FULL, send, off);
		}
		off >>= 1;
	}
#pragma unroll
	for(; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(FULL, v[0], off);
}

template<int K> __device__ void block_sum(double (&v)[K], double (*s_red)[48], double *s_sum){
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	if constexpr(K >= 24){
		static_assert(K <= 48, "block_sum: at most 48 values");
		double a[32], b[16];
#pragma unroll
		for(int k = 0; k < 32; ++k) a[k] = v[k];
#pragma unroll
		for(int k = 0; k < 16; ++k) b[k] = 32 + k < K ? v[32 + k < K ? 32 + k : 0] : 0.0;
		warp_transpose_sum<32>(a, lane);
		warp_transpose_sum<16>(b, lane);
		s_red[warp][lane] = a[0];
		if(!(lane & 1) && 32 + (lane >> 1) < K) s_red[warp][32 + (lane >> 1)] = b[0];
	} else {
#pragma unroll
		for(int k = 0; k < K; ++k){
			double x = v[k];
#pragma unroll
			for(int off = 16; off >= 1; off >>= 1) x += __shfl_xor_sync(FULL, x, off);
			if(lane == 0) s_red[warp][k] = x;
		}
	}
	__syncthreads();
	if(tid < K){ double s = 0; for(int w = 0; w < EST_WARPS; ++w) s += s_red[w][tid]; s_sum[tid] = s; }
	__syncthreads();
}

// runKernel on the points mask selects, by the whole block; the model goes to s_model
template<bool HOM> __device__ int fit_masked(const EstDev &e, WarpWs &ws, double (*s_red)[48], double *s_sum, double *s_model){
	const int tid = threadIdx.x, n = e.n;
	double a[5] = { 0, 0, 0, 0, 0 };
	for(int i = tid; i < n; i += EST_THREADS) if(e.mask[i]){
		a[0] += e.out_pts[2 * i]; a[1] += e.out_pts[2 * i + 1]; a[2] += e.in_pts[2 * i]; a[3] += e.in_pts[2 * i + 1]; a[4] += 1;
	}
	block_sum<5>(a, s_red, s_sum);
	const double count = s_sum[4];
	const double cmx = s_sum[0] / count, cmy = s_sum[1] / count, cMx = s_sum[2] / count, cMy = s_sum[3] / count;
	if(HOM){
		double d[4] = { 0, 0, 0, 0 };
		for(int i = tid; i < n; i += EST_THREADS) if(e.mask[i]){
			d[0] += fabs(e.out_pts[2 * i] - cmx); d[1] += fabs(e.out_pts[2 * i + 1] - cmy);
			d[2] += fabs(e.in_pts[2 * i] - cMx); d[3] += fabs(e.in_pts[2 * i + 1] - cMy);
		}
		block_sum<4>(d, s_red, s_sum);
		double smx = s_sum[0], smy = s_sum[1], sMx = s_sum[2], sMy = s_sum[3];
		if(fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON) return 0;
		smx = count / smx; smy = count / smy; sMx = count / sMx; sMy = count / sMy;
		double L[45];
#pragma unroll
		for(int k = 0; k < 45; ++k) L[k] = 0;
		for(int i = tid; i < n; i += EST_THREADS) if(e.mask[i]){
			const double x = (e.out_pts[2 * i] - cmx) * smx, y = (e.out_pts[2 * i + 1] - cmy) * smy;
			const double X = (e.in_pts[2 * i] - cMx) * sMx, Y = (e.in_pts[2 * i + 1] - cMy) * sMy;
			const double Lx[9] = { X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x };
			const double Ly[9] = { 0, 0, 0, X, Y, 1, -y * X, -y * Y, -y };
			int q = 0;
#pragma unroll
			for(int j = 0; j < 9; ++j)
#pragma unroll
				for(int k = j; k < 9; ++k) L[q++] += Lx[j] * Lx[k] + Ly[j] * Ly[k];
		}
		block_sum<45>(L, s_red, s_sum);
		if(tid < 32){
			if(tid == 0){
				int q = 0;
				for(int j = 0; j < 9; ++j) for(int k = j; k < 9; ++k){ ws.A[j * 9 + k] = s_sum[q]; ws.A[k * 9 + j] = s_sum[q]; ++q; }
			}
			__syncwarp();
			smallest_eigvec9_warp(ws.A, ws.x, tid);
			if(tid == 0) denormalise(ws.x, cmx, cmy, smx, smy, cMx, cMy, sMx, sMy, s_model);
		}
		__syncthreads();
		return 1;
	} else {
		double d[7] = { 0, 0, 0, 0, 0, 0, 0 };
		for(int i = tid; i < n; i += EST_THREADS) if(e.mask[i]){
			const double X = e.in_pts[2 * i] - cMx, Y = e.in_pts[2 * i + 1] - cMy, x = e.out_pts[2 * i] - cmx, y = e.out_pts[2 * i + 1] - cmy;
			d[0] += X * X; d[1] += X * Y; d[2] += Y * Y; d[3] += X * x; d[4] += Y * x; d[5] += X * y; d[6] += Y * y;
		}
		block_sum<7>(d, s_red, s_sum);
		__shared__ int s_ok;
		if(tid == 0){
			double H[9];
			s_ok = aff_from_moments(count, cMx, cMy, cmx, cmy, s_sum[0], s_sum[1], s_sum[2], s_sum[3], s_sum[4], s_sum[5], s_sum[6], H);
			if(s_ok) for(int i = 0; i < 9; ++i) s_model[i] = H[i];
		}
		__syncthreads();
		return s_ok;
	}
}

// mask[i] = err_i <= threshold^2 for every point, returns the count (SSMEstimator::findInliers :35-46) -- whole block
template<bool HOM> __device__ int find_inliers_block(const EstDev &e, const double *H, double threshold, double (*s_red)[48], double *s_sum){
	const double t2 = threshold * threshold;
	double c[1] = { 0 };
	for(int i = threadIdx.x; i < e.n; i += EST_THREADS){
		const float er = reproj_err<HOM>(H, e.in_pts[2 * i], e.in_pts[2 * i + 1], e.out_pts[2 * i], e.out_pts[2 * i + 1]);
		const bool in = (double)er <= t2;
		e.mask[i] = in; c[0] += in;
	}
	block_sum<1>(c, s_red, s_sum);
	return (int)s_sum[0];
}

// LevMarq::step (SSMEstimator.cc:489-516): param = prevParam - (JtJ with its diagonal scaled by 1 + lambda)^-1 JtErr, by one
// warp: lane i holds row i of the diagonally scaled system, elimination with partial pivoting without moving rows (the pivot
// row is broadcast, a row that has been a pivot drops out), back-substitution by broadcasting each unknown
template<int NP> __device__ void lm_step_warp(const double *JtJ, const double *JtErr, const double *prev, double *param, int lambdaLg10, int lane){
	const double lambda = exp(lambdaLg10 * log(10.));
	const int row = lane < NP ? lane : NP - 1;
	double sc[NP], a[NP];
#pragma unroll
	for(int j = 0; j < NP; ++j){ const double dd = JtJ[j * NP + j] * (1. + lambda); sc[j] = dd > 0 ? rsqrt(dd) : 1.0; }
	double scr = 1.0;
#pragma unroll
	for(int j = 0; j < NP; ++j) if(row == j) scr = sc[j];
#pragma unroll
	for(int j = 0; j < NP; ++j){
		double v = row <= j ? JtJ[row * NP + j] : JtJ[j * NP + row];
		if(row == j) v *= 1. + lambda;
		a[j] = v * scr * sc[j];
	}
	double b = JtErr[row] * scr;
	bool used = lane >= NP;
	int piv[NP];
	double pinv[NP];
#pragma unroll
	for(int k = 0; k < NP; ++k){
		double mag = used ? -1.0 : fabs(a[k]);
		int who = lane;
#pragma unroll
		for(int off = 16; off >= 1; off >>= 1){
			const double m2 = __shfl_xor_sync(FULL, mag, off);
			const int w2 = __shfl_xor_sync(FULL, who, off);
			if(m2 > mag || (m2 == mag && w2 < who)){ mag = m2; who = w2; }
		}
		piv[k] = who;
		double pk[NP];
#pragma unroll
		for(int j = k; j < NP; ++j) pk[j] = __shfl_sync(FULL, a[j], who);
		const double pb = __shfl_sync(FULL, b, who);
		double dpv = pk[k];
		if(fabs(dpv) < 1e-300) dpv = 1e-300;
		pinv[k] = 1.0 / dpv;                       // one reciprocal per pivot, shared by the elimination and the back-substitution
		if(lane == who) used = true;
		if(!used){
			const double f = a[k] * pinv[k];
#pragma unroll
			for(int j = k + 1; j < NP; ++j) a[j] -= f * pk[j];
			b -= f * pb;
		}
	}
	double x[NP];
#pragma unroll
	for(int k = NP - 1; k >= 0; --k){
		double t = b;
#pragma unroll
		for(int j = k + 1; j < NP; ++j) t -= a[j] * x[j];
		x[k] = __shfl_sync(FULL, t * pinv[k], piv[k]);
	}
#pragma unroll
	for(int k = 0; k < NP; ++k) if(lane == k) param[k] = prev[k] - x[k] * sc[k];
	__syncwarp();
}

// HomographyEstimator::refine (:97-145) / AffineEstimator::refine (:64-106) around LevMarq::updateAlt (:436-487)
template<bool HOM> __device__ int lm_refine(const EstDev &e, double (*s_red)[48], double *s_sum, double *s_model){
	constexpr int NP = HOM ? 8 : 6, NJ = NP * (NP + 1) / 2, K = NJ + NP + 1;
	enum { DONE = 0, STARTED = 1, CALC_J = 2, CHECK_ERR = 3 };
	__shared__ double s_JtJ[64], s_JtErr[8], s_param[8], s_prev[8];
	__shared__ int s_flag[5];
	const int tid = threadIdx.x;
	int state = STARTED, iters = 0, lambdaLg10 = -3, evals = 0;
	const int max_iter = min(max(e.lm_max_iters, 1), 1000);
	double prevErrNorm = DBL_MAX, errNorm = 0;
	if(tid < NP) s_param[tid] = s_model[tid];
	__syncthreads();
	for(;;){
		{ PROF_T0();
		if(tid < 32){
			if(tid == 0){
				int cont = 1, want_J = 0, want_err = 0, step = 0;
				if(state == STARTED){ errNorm = 0; want_J = want_err = 1; state = CALC_J; }
				else if(state == CALC_J){
					for(int i = 0; i < NP; ++i) s_prev[i] = s_param[i];
					step = 1;
					prevErrNorm = errNorm; errNorm = 0; want_err = 1; state = CHECK_ERR;
				} else {
					bool retried = false;
					if(errNorm > prevErrNorm){
						if(++lambdaLg10 <= 16){ step = 1; errNorm = 0; want_err = 1; state = CHECK_ERR; retried = true; }
					}
					if(!retried){
						lambdaLg10 = max(lambdaLg10 - 1, -16);
						double dn = 0, pn = 0;
						for(int i = 0; i < NP; ++i){ const double d = s_param[i] - s_prev[i]; dn += d * d; pn += s_prev[i] * s_prev[i]; }
						const double change = sqrt(dn) / (sqrt(pn) + DBL_EPSILON);
						if(++iters >= max_iter || change < DBL_EPSILON){ state = DONE; cont = 0; }
						else { prevErrNorm = errNorm; want_J = 1; state = CALC_J; }
					}
				}
				s_flag[0] = cont; s_flag[1] = want_J; s_flag[2] = want_err; s_flag[3] = step; s_flag[4] = lambdaLg10;
			}
			__syncwarp();
			if(s_flag[3]) lm_step_warp<NP>(s_JtJ, s_JtErr, s_prev, s_param, s_flag[4], tid);
		}
		__syncthreads();
		PROF_ADD(4); }
		if(!s_flag[0]) break;
		PROF_T0();
		const bool want_J = s_flag[1] != 0, want_err = s_flag[2] != 0;
		++evals;
		double h[NP];
#pragma unroll
		for(int i = 0; i < NP; ++i) h[i] = s_param[i];
		double acc[K];
#pragma unroll
		for(int k = 0; k < K; ++k) acc[k] = 0;
		for(int i = tid; i < e.n; i += EST_THREADS) if(e.mask[i]){
			const double Mx = e.in_pts[2 * i], My = e.in_pts[2 * i + 1];
			double J0[NP], J1[NP], er0, er1;
			if(HOM){
				double ww = h[6] * Mx + h[7] * My + 1.;
				ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
				const double xi = (h[0] * Mx + h[1] * My + h[2]) * ww, yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
				er0 = xi - e.out_pts[2 * i]; er1 = yi - e.out_pts[2 * i + 1];
				J0[0] = Mx * ww; J0[1] = My * ww; J0[2] = ww; J0[3] = 0; J0[4] = 0; J0[5] = 0; J0[6] = -Mx * ww * xi; J0[7] = -My * ww * xi;
				J1[0] = 0; J1[1] = 0; J1[2] = 0; J1[3] = Mx * ww; J1[4] = My * ww; J1[5] = ww; J1[6] = -Mx * ww * yi; J1[7] = -My * ww * yi;
			} else {
				const double xi = h[0] * Mx + h[1] * My + h[2], yi = h[3] * Mx + h[4] * My + h[5];
				er0 = xi - e.out_pts[2 * i]; er1 = yi - e.out_pts[2 * i + 1];
				J0[0] = Mx; J0[1] = My; J0[2] = 1; J0[3] = 0; J0[4] = 0; J0[5] = 0;
				J1[0] = 0; J1[1] = 0; J1[2] = 0; J1[3] = Mx; J1[4] = My; J1[5] = 1;
			}
			if(want_J){
				// the rows of J are zero in columns 3-5 / 0-2: only the products that are not identically zero (the reference adds the zeros)
				int q = 0;
#pragma unroll
				for(int j = 0; j < NP; ++j){
#pragma unroll
					for(int k = j; k < NP; ++k){
						const bool a0 = (j < 3 || j > 5) && (k < 3 || k > 5), a1 = j >= 3 && k >= 3;
						if(a0 && a1) acc[q] += J0[j] * J0[k] + J1[j] * J1[k];
						else if(a0) acc[q] += J0[j] * J0[k];
						else if(a1) acc[q] += J1[j] * J1[k];
						++q;
					}
					if(j < 3) acc[NJ + j] += J0[j] * er0;
					else if(j < 6) acc[NJ + j] += J1[j] * er1;
					else acc[NJ + j] += J0[j] * er0 + J1[j] * er1;
				}
			}
			acc[K - 1] += er0 * er0 + er1 * er1;
		}
		PROF_ADD(5);
		block_sum<K>(acc, s_red, s_sum);
		PROF_ADD(6);
		if(tid == 0){
			if(want_J){
				int q = 0;
				for(int j = 0; j < NP; ++j){ for(int k = j; k < NP; ++k) s_JtJ[j * NP + k] = s_sum[q++]; s_JtErr[j] = s_sum[NJ + j]; }
			}
			if(want_err) errNorm += s_sum[K - 1];
		}
	}
	if(tid < NP) s_model[tid] = s_param[tid];
	__syncthreads();
	return evals;
}

template<bool HOM>
__global__ void __launch_bounds__(EST_THREADS, 1) estimate_kernel(EstDev e_in){
	const long long prof_k0 = clock64(); (void)prof_k0;
	// up to EST_STAGED points live in shared memory for the whole kernel (every phase walks them; the subset draw by one thread
	// is pure load latency otherwise); the mask goes back at the end
	__shared__ float s_in[2 * EST_STAGED], s_out[2 * EST_STAGED];
	__shared__ unsigned char s_mask[EST_STAGED];
	EstDev e = e_in;
	const bool staged = e_in.n <= EST_STAGED;
	if(staged){
		for(int i = threadIdx.x; i < 2 * e_in.n; i += EST_THREADS){ s_in[i] = e_in.in_pts[i]; s_out[i] = e_in.out_pts[i]; }
		e.in_pts = s_in; e.out_pts = s_out; e.mask = s_mask;
	}
	__shared__ WarpWs s_ws[EST_WARPS];
	__shared__ int s_idx[BATCH][EST_MAX_MODEL_PTS];
	__shared__ int s_found[BATCH], s_valid[BATCH], s_good[BATCH];
	__shared__ double s_hyp[BATCH][9], s_median[BATCH];
	__shared__ double s_model[9];
	__shared__ double s_red[EST_WARPS][48], s_sum[48];
	__shared__ unsigned s_hist[EST_WARPS][256];
	__shared__ int s_ctl[2];
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int n = e.n, mp = e.model_points;
	const int method = n == mp ? (int)M_LS : e.method;
	constexpr int NM = HOM ? 9 : 6;

	for(int i = tid; i < n; i += EST_THREADS) e.mask[i] = 1;
	if(tid < 9) s_model[tid] = tid == 8 ? 1.0 : 0.0;
	__syncthreads();

	int result = 0, drawn = 0;
	if(method == M_LS){
		result = fit_masked<HOM>(e, s_ws[0], s_red, s_sum, s_model);
	} else {
		// thread 0's sequential state
		CvRng rng; rng.s = e.seed ? e.seed : ~0ULL;
		int iter = 0, niters = e.max_iters, max_good = 0;
		double min_median = DBL_MAX;
		bool first_failed = false;
		if(method == M_LMEDS){
			niters = __double2int_rn(log(1 - e.confidence) / log(1 - pow(1 - 0.45, (double)mp)));
			niters = min(max(niters, 3), e.max_iters);
		}
		for(;;){
			PROF_T0();
			if(tid == 0){
				// the first round is one hypothesis per warp: with mostly inliers RANSAC's adaptive count ends there
				int nb = min(iter == 0 ? (int)EST_WARPS : BATCH, niters - iter);
				for(int b = 0; b < nb; ++b){
					s_found[b] = get_subset(rng, e.in_pts, e.out_pts, n, mp, e.max_attempts, s_idx[b]);
					if(!s_found[b]){ nb = b + 1; break; }
				}
				s_ctl[0] = nb;
			}
			__syncthreads();
			PROF_ADD(0);
			const int nb = s_ctl[0];
			for(int b = warp; b < nb; b += EST_WARPS){
				if(!s_found[b]) continue;
				WarpWs &w = s_ws[warp];
				if(lane < mp){
					const int id = s_idx[b][lane];
					w.M[2 * lane] = e.in_pts[2 * id]; w.M[2 * lane + 1] = e.in_pts[2 * id + 1];
					w.m[2 * lane] = e.out_pts[2 * id]; w.m[2 * lane + 1] = e.out_pts[2 * id + 1];
				}
				__syncwarp();
				if(HOM){ const int ok = hom_fit_subset(w, mp, s_hyp[b], lane); if(lane == 0) s_valid[b] = ok; }
				else if(lane == 0) s_valid[b] = aff_fit_subset(w, mp, s_hyp[b]);
				__syncwarp();
				if(!s_valid[b]) continue;
				double H[9];
#pragma unroll
				for(int i = 0; i < NM; ++i) H[i] = s_hyp[b][i];
				if(method == M_RANSAC){
					const double t2 = e.thresh * e.thresh;
					int good = 0;
					for(int i = lane; i < n; i += 32){
						const float er = reproj_err<HOM>(H, e.in_pts[2 * i], e.in_pts[2 * i + 1], e.out_pts[2 * i], e.out_pts[2 * i + 1]);
						good += (double)er <= t2;
					}
#pragma unroll
					for(int off = 16; off >= 1; off >>= 1) good += __shfl_xor_sync(FULL, good, off);
					if(lane == 0) s_good[b] = good;
				} else {
					float *er = e.err + (size_t)warp * n;
					for(int i = lane; i < n; i += 32)
						er[i] = reproj_err<HOM>(H, e.in_pts[2 * i], e.in_pts[2 * i + 1], e.out_pts[2 * i], e.out_pts[2 * i + 1]);
					__syncwarp();
					const unsigned *bits = reinterpret_cast<const unsigned *>(er);
					const float hi = __uint_as_float(warp_select(bits, n, n / 2, s_hist[warp], lane));
					double median = hi;
					if(n % 2 == 0){
						const float lo = __uint_as_float(warp_select(bits, n, n / 2 - 1, s_hist[warp], lane));
						median = (double)(lo + hi) * 0.5;
					}
					if(lane == 0) s_median[b] = median;
				}
			}
			__syncthreads();
			PROF_ADD(1);
			if(tid == 0){
				// the reference's loop body over this round's hypotheses, in order (SSMEstimator.cc:101-127 / :176-202)
				int stop = 0;
				for(int b = 0; b < nb && iter < niters; ++b, ++iter){
					if(!s_found[b]){ if(iter == 0) first_failed = true; stop = 1; break; }
					if(!s_valid[b]) continue;
					if(method == M_RANSAC){
						const int good = s_good[b];
						if(good > max(max_good, mp - 1)){
							for(int i = 0; i < NM; ++i) s_model[i] = s_hyp[b][i];
							max_good = good;
							niters = ransac_update_num_iters(e.confidence, (double)(n - good) / n, mp, niters);
						}
					} else if(s_median[b] < min_median){
						min_median = s_median[b];
						for(int i = 0; i < NM; ++i) s_model[i] = s_hyp[b][i];
					}
				}
				if(iter >= niters) stop = 1;
				s_ctl[1] = stop;
			}
			__syncthreads();
			if(s_ctl[1]) break;
		}
		__shared__ int s_res[2];
		__shared__ double s_sigma;
		if(tid == 0){
			int mode = 0;         // 0: failed, mask stays all ones; 1: RANSAC mask of the best model; 2: LMedS mask at sigma
			if(!first_failed){
				if(method == M_RANSAC) mode = max_good > 0 ? 1 : 0;
				else if(min_median < DBL_MAX){
					double sigma = 2.5 * 1.4826 * (1 + 5. / (n - mp)) * sqrt(min_median);
					s_sigma = fmax(sigma, 0.001); mode = 2;
				}
			}
			s_res[0] = mode; s_res[1] = iter;
		}
		__syncthreads();
		drawn = s_res[1];
		if(s_res[0] == 1){ find_inliers_block<HOM>(e, s_model, e.thresh, s_red, s_sum); result = 1; }
		else if(s_res[0] == 2){ const int cnt = find_inliers_block<HOM>(e, s_model, s_sigma, s_red, s_sum); result = cnt >= mp; }
	}
	__syncthreads();
	// estimateHomography :205-214: the points the mask keeps, the model re-fitted on them after RANSAC, the refinement
	int n_in = n, evals = 0;
	if(result && n > mp){
		double c[1] = { 0 };
		for(int i = tid; i < n; i += EST_THREADS) c[0] += e.mask[i] != 0;
		block_sum<1>(c, s_red, s_sum);
		n_in = (int)s_sum[0];
		__syncthreads();
		PROF_T0();
		if(method == M_RANSAC) fit_masked<HOM>(e, s_ws[0], s_red, s_sum, s_model);
		PROF_ADD(2);
		if(e.refine) evals = lm_refine<HOM>(e, s_red, s_sum, s_model);
		PROF_ADD(3);
	}
	if(staged){
		__syncthreads();
		for(int i = tid; i < n; i += EST_THREADS) e_in.mask[i] = s_mask[i];
	}
	if(tid == 0){
		double H[9];
		for(int i = 0; i < 9; ++i) H[i] = s_model[i];
		if(!HOM){ H[6] = 0; H[7] = 0; H[8] = 1; }
		if(!result) for(int i = 0; i < 9; ++i) H[i] = 0;
		for(int i = 0; i < 9; ++i) e.out[i] = H[i];
		double *su = e.out + 9;
		if(HOM){
			// Homography::estimateWarpFromPts (Homography.cc:885-897)
			su[0] = H[0] - 1; su[1] = H[1]; su[2] = H[2]; su[3] = H[3]; su[4] = H[4] - 1; su[5] = H[5]; su[6] = H[6]; su[7] = H[7];
		} else {
			// Affine::estimateWarpFromPts (Affine.cc:359-369)
			su[0] = H[2]; su[1] = H[5]; su[2] = H[0] - 1; su[3] = H[1]; su[4] = H[3]; su[5] = H[4] - 1; su[6] = 0; su[7] = 0;
		}
		// result | hypotheses drawn | inliers | LM evaluations, in the same buffer: one copy back
		e.out[17] = result; e.out[18] = drawn; e.out[19] = n_in; e.out[20] = evals;
#ifdef EST_PROF
		g_prof[10] = (double)(clock64() - prof_k0);
		for(int i = 0; i < 11; ++i){ e.out[21 + i] = g_prof[i]; g_prof[i] = 0; }
#endif
	}
}

__global__ void centroid_kernel(const double *corners, int P, float *pts){
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if(p >= P) return;
	const double *c = corners + 8 * (size_t)p;
	pts[2 * p] = (float)((c[0] + c[1] + c[2] + c[3]) / 4.0);
	pts[2 * p + 1] = (float)((c[4] + c[5] + c[6] + c[7]) / 4.0);
}

// One CTA.  Warp 0: the new region and its DLT (warp_homography_dlt: the reference's computeHomographyDLT, lk_warp.cuh); then
// one thread per cell.
__global__ void grid_layout_kernel(int homography, const double *__restrict__ su, double *region, int gx, int gy, double psx, double psy,
	double *__restrict__ cells){
	__shared__ double s_c[8], s_H[9];
	const int tid = threadIdx.x, lane = tid & 31;
	if(tid < 32){
		// ssm.applyWarpToCorners: q = W [c; 1], c' = q / q_z with W = getWarpFromState(state_update) (Homography.cc:94-107, Affine.cc:117-131)
		double W[9];
		if(homography){ W[0] = 1 + su[0]; W[1] = su[1]; W[2] = su[2]; W[3] = su[3]; W[4] = 1 + su[4]; W[5] = su[5]; W[6] = su[6]; W[7] = su[7]; W[8] = 1; }
		else{ W[0] = 1 + su[2]; W[1] = su[3]; W[2] = su[0]; W[3] = su[4]; W[4] = 1 + su[5]; W[5] = su[1]; W[6] = 0; W[7] = 0; W[8] = 1; }
		if(lane < 4){
			const double x = region[lane], y = region[4 + lane];
			const double qx = W[0]*x + W[1]*y + W[2], qy = W[3]*x + W[4]*y + W[5], qz = W[6]*x + W[7]*y + W[8];
			s_c[lane] = qx / qz; s_c[4 + lane] = qy / qz;
		}
		__syncwarp();
		const double nc[8] = { -0.5, 0.5, 0.5, -0.5, -0.5, -0.5, 0.5, 0.5 };
		double c[8];
#pragma unroll
		for(int i = 0; i < 8; ++i) c[i] = s_c[i];
		const Mat3 H = warp_homography_dlt(nc, c, lane);
		if(lane < 9) s_H[lane] = H.m[lane];
		if(lane < 8) region[lane] = c[lane];
	}
	__syncthreads();
	const int w = gx + 1, h = gy + 1;
	const double sx = 1.0 / (w - 1), sy = 1.0 / (h - 1);
	for(int t = tid; t < gx*gy; t += blockDim.x){
		const int r = t / gx, cc = t - r*gx;
		double cx = 0, cy = 0;
		// the four grid points of the cell: (r, c), (r, c + 1), (r + 1, c + 1), (r + 1, c) -- GridTracker.cc:361-371; LinSpaced values
#pragma unroll
		for(int k = 0; k < 4; ++k){
			const int j = cc + ((k == 1 || k == 2) ? 1 : 0), i = r + ((k >= 2) ? 1 : 0);
			const double x = (j == w - 1) ? 0.5 : -0.5 + j*sx, y = (i == h - 1) ? 0.5 : -0.5 + i*sy;
			const double qx = s_H[0]*x + s_H[1]*y + s_H[2], qy = s_H[3]*x + s_H[4]*y + s_H[5], qz = s_H[6]*x + s_H[7]*y + s_H[8];
			cx += qx / qz; cy += qy / qz;
		}
		cx /= 4.0; cy /= 4.0;
		const double x0 = cx - psx / 2.0, y0 = cy - psy / 2.0, x1 = x0 + psx, y1 = y0 + psy;
		double *o = cells + 8 * (size_t)t;
		o[0] = x0; o[1] = x1; o[2] = x1; o[3] = x0; o[4] = y0; o[5] = y0; o[6] = y1; o[7] = y1;
	}
}

} // namespace

cudaError_t launch_grid_layout(int homography, const double *d_state_update, double *d_region, int gx, int gy, double psx, double psy,
	double *d_cells, cudaStream_t st){
	grid_layout_kernel<<<1, 256, 0, st>>>(homography, d_state_update, d_region, gx, gy, psx, psy, d_cells);
	return cudaGetLastError();
}

cudaError_t launch_centroids(const double *corners, int P, float *pts, cudaStream_t st){
	centroid_kernel<<<(P + 255) / 256, 256, 0, st>>>(corners, P, pts);
	return cudaGetLastError();
}

cudaError_t launch_estimate(const EstDev &e, cudaStream_t st){
	if(e.homography) estimate_kernel<true><<<1, EST_THREADS, 0, st>>>(e);
	else estimate_kernel<false><<<1, EST_THREADS, 0, st>>>(e);
	return cudaGetLastError();
}

} // namespace mtfb
