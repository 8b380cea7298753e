// lk_warp.cuh -- warp-level building blocks of the LK kernels (device only):
//   * transposing butterfly reduction of a per-thread accumulator vector
//   * column-pivoted Householder QR (the solver behind Eigen's colPivHouseholderQr().solve() that
//     nt::FCLK / ESM / ICLK call every iteration: SM/src/NT/FCLK.cc:298, NT/ESM.cc:266, NT/ICLK.cc:228)
//     with one matrix column per lane
//   * the 4-point DLT of utils::computeHomographyDLT (Utilities/src/warpUtils.cc:171-223)
#pragma once
#include <cfloat>
#include "lk_math.cuh"

namespace mtfb {

constexpr unsigned FULL_MASK = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// Butterfly "reduce-scatter" across the 32 lanes of a warp.
// In: v[0..CNT) per lane.  Out: the warp-wide sum of entry e is left in exactly one lane; every lane ends
// with at most 3 live entries in v[0..2] whose entry indices are idx[0..2] (-1 = none).
// Costs about CNT 64-bit shuffles in total instead of 5*CNT for CNT independent all-reduces.
// ------------------------------------------------------------------------------------------------
template<int CNT, int MASK> struct Butterfly {
	static constexpr int HALF = (CNT + 1) / 2;
	template<class V, int FULL> __device__ __forceinline__ static void run(V (&v)[FULL], int lane, int &base, bool &ok){
		const bool up = (lane & MASK) != 0;
#pragma unroll
		for(int i = 0; i < HALF; ++i){
			V lo = v[i];
			V hi = (i + HALF < CNT) ? v[i + HALF] : V(0);
			V keep = up ? hi : lo;
			V send = up ? lo : hi;
			v[i] = keep + __shfl_xor_sync(FULL_MASK, send, MASK);
		}
		// slot i of this lane now stands for previous-level slot i + HALF*up
		// (caller composes the index map; see warp_reduce_scatter)
		(void)base; (void)ok;
	}
};

template<int CNT> struct ReduceMap {
	static constexpr int H1 = (CNT + 1) / 2, H2 = (H1 + 1) / 2, H3 = (H2 + 1) / 2, H4 = (H3 + 1) / 2, H5 = (H4 + 1) / 2;
	static_assert(H5 <= 3, "accumulator vector too long for the 32-lane butterfly (max 96 entries)");
};

// returns through idx[0..2] the accumulator entry each of v[0..2] holds (or -1)
template<int CNT, class V> __device__ __forceinline__ void warp_reduce_scatter(V (&v)[CNT], int lane, int (&idx)[3]){
	typedef ReduceMap<CNT> M;
	int dummy = 0; bool ok = true;
	Butterfly<CNT, 16>::run(v, lane, dummy, ok);
	Butterfly<M::H1, 8>::run(v, lane, dummy, ok);
	Butterfly<M::H2, 4>::run(v, lane, dummy, ok);
	Butterfly<M::H3, 2>::run(v, lane, dummy, ok);
	Butterfly<M::H4, 1>::run(v, lane, dummy, ok);
	const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1, b0 = lane & 1;
#pragma unroll
	for(int j = 0; j < 3; ++j){
		bool valid = j < M::H5;
		int s4 = j + M::H5*b0;  valid = valid && (s4 < M::H4);     // slot at level 4
		int s3 = s4 + M::H4*b1; valid = valid && (s3 < M::H3);
		int s2 = s3 + M::H3*b2; valid = valid && (s2 < M::H2);
		int s1 = s2 + M::H2*b3; valid = valid && (s1 < M::H1);
		int s0 = s1 + M::H1*b4; valid = valid && (s0 < CNT);
		idx[j] = valid ? s0 : -1;
	}
}

// 1 / sqrt(s) and 1 / d for the solve of the F32 precision (factor_lean): the fp64 MUFU seeds (rsqrt.approx.f64 / rcp.approx.f64,
// SASS MUFU.RSQ64H / MUFU.RCP64H: about 22 bits from the upper word, any exponent) and two Newton steps in fp64 (relative error
// ~2e-16, not correctly rounded).  The inputs of that solve carry fp32 rounding from the pixel sums, so the last bit is noise
// there; the F64 kernels keep sqrt() / __drcp_rn().
__device__ __forceinline__ double rsqrt_seed(double s){
	double r;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
	return r;
}
__device__ __forceinline__ double rcp_seed(double d){
	double r;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
	return r;
}
__device__ __forceinline__ double rsqrt_newton(double s){
	const int hi = __double2hiint(s);
	if(!((unsigned)(hi - 0x00100000) < 0x7fe00000u)) return 1.0 / sqrt(s);            // zero, subnormal, negative, inf, nan
	double r = rsqrt_seed(s);
	double sr = s * r;
	r = fma(0.5 * r, fma(-sr, r, 1.0), r);
	sr = s * r;
	r = fma(0.5 * r, fma(-sr, r, 1.0), r);
	return r;
}
__device__ __forceinline__ double rcp_newton_scaled(double d){
	const int ef = (__double2hiint(d) >> 20) & 0x7ff;
	if(!(ef > 64 && ef < 1983)) return __drcp_rn(d);                                   // keeps the result's exponent normal too
	double r = rcp_seed(d);
	r = fma(fma(-d, r, 1.0), r, r);
	r = fma(fma(-d, r, 1.0), r, r);
	return r;
}

// ------------------------------------------------------------------------------------------------
// Column-pivoted Householder QR, Eigen 3.3 ColPivHouseholderQR::computeInPlace + _solve_impl semantics
// (dgeqp3-style norm down-dating), ROWS x COLS, one column per lane (lanes 0..COLS-1); lane COLS may
// carry a right-hand side that receives the reflectors on the fly.  Every loop is unrolled so the
// per-lane column lives in registers.
// ------------------------------------------------------------------------------------------------
template<int ROWS, int COLS> struct WarpColPivQR {
	static constexpr int SIZE = ROWS < COLS ? ROWS : COLS;
	double a[ROWS];          // this lane's column (or the rhs on lane COLS)
	double tau[SIZE];        // Householder coefficients (uniform)
	int pos;                 // current position of this lane's column (-1: not a column)
	int lane_at_pos[SIZE];   // uniform, valid after factor()
	int nonzero_pivots;      // uniform

	__device__ __forceinline__ void factor(int lane, bool with_rhs){
		const bool is_col = lane < COLS;
		const bool is_rhs = with_rhs && (lane == COLS);
		pos = is_col ? lane : -1;
		double cnd, cnu;
		{
			double s = 0;
#pragma unroll
			for(int r = 0; r < ROWS; ++r) s += a[r] * a[r];
			cnd = sqrt(s); cnu = cnd;
		}
		double maxn;
		{
			const bool ok = is_col && (cnu == cnu);
			const unsigned hi = ok ? (unsigned)__double2hiint(cnu) : 0u;
			const unsigned mhi = __reduce_max_sync(FULL_MASK, hi);
			const unsigned lo = (ok && hi == mhi) ? (unsigned)__double2loint(cnu) : 0u;
			const unsigned mlo = __reduce_max_sync(FULL_MASK, lo);
			maxn = __hiloint2double((int)mhi, (int)mlo);
		}
		// Eigen 3.3 computeInPlace(): threshold_helper = abs2(colNormsUpdated.maxCoeff() * epsilon) / rows
		const double me = maxn * DBL_EPSILON;
		const double threshold_helper = me * me / double(ROWS);
		const double norm_downdate_threshold = sqrt(DBL_EPSILON);
		nonzero_pivots = SIZE;
#pragma unroll
		for(int k = 0; k < SIZE; ++k){
			// pivot: first position >= k holding the largest updated column norm.  Norms are >= 0, so their
			// bit patterns order like unsigned integers: two 32-bit warp max-reductions find the largest value,
			// a min-reduction over positions breaks ties the way Eigen's left-to-right scan does.
			const bool cand_ok = is_col && pos >= k && (cnu == cnu);
			const unsigned hi = cand_ok ? (unsigned)__double2hiint(cnu) : 0u;
			const unsigned mhi = __reduce_max_sync(FULL_MASK, hi);
			const bool c1 = cand_ok && hi == mhi;
			const unsigned lo = c1 ? (unsigned)__double2loint(cnu) : 0u;
			const unsigned mlo = __reduce_max_sync(FULL_MASK, lo);
			const bool c2 = c1 && lo == mlo;
			int biggest = (int)__reduce_min_sync(FULL_MASK, c2 ? (unsigned)pos : 0xffffu);
			double bmax = __hiloint2double((int)mhi, (int)mlo);
			if(biggest == 0xffff){ biggest = k; bmax = 0; }           // all-NaN norms: keep column k
			if(nonzero_pivots == SIZE && bmax * bmax < threshold_helper * double(ROWS - k)) nonzero_pivots = k;
			if(is_col){ if(pos == biggest) pos = k; else if(pos == k) pos = biggest; }
			const int piv_lane = __ffs(__ballot_sync(FULL_MASK, is_col && pos == k)) - 1;
			// makeHouseholderInPlace on this lane's column tail (only the pivot lane's result is used)
			double tau_l = 0;
			{
				double c0 = a[k], tail_sq = 0;
#pragma unroll
				for(int r = k + 1; r < ROWS; ++r) tail_sq += a[r] * a[r];
				if(lane == piv_lane){
					double beta;
					if(tail_sq <= DBL_MIN){
						tau_l = 0; beta = c0;
#pragma unroll
						for(int r = k + 1; r < ROWS; ++r) a[r] = 0;
					} else{
						beta = sqrt(c0 * c0 + tail_sq);
						if(c0 >= 0) beta = -beta;
						{
							const double den = c0 - beta, rden = ieee_rcp(den);        // tail /= (c0 - beta), exactly rounded
#pragma unroll
							for(int r = k + 1; r < ROWS; ++r) a[r] = div_by(a[r], den, rden);
						}
						tau_l = (beta - c0) / beta;
					}
					a[k] = beta;
				}
			}
			const double tk = __shfl_sync(FULL_MASK, tau_l, piv_lane);
			tau[k] = tk;
			double v[ROWS];
#pragma unroll
			for(int r = k + 1; r < ROWS; ++r) v[r] = __shfl_sync(FULL_MASK, a[r], piv_lane);
			// applyHouseholderOnTheLeft to the remaining columns (and the rhs while k < nonzero_pivots)
			const bool apply = (is_col && pos > k) || (is_rhs && k < nonzero_pivots);
			if(apply){
				if(ROWS - k == 1){
					a[k] *= (1 - tk);
				} else if(tk != 0){
					double t = 0;
#pragma unroll
					for(int r = k + 1; r < ROWS; ++r) t += v[r] * a[r];
					t += a[k];
					a[k] -= tk * t;
#pragma unroll
					for(int r = k + 1; r < ROWS; ++r) a[r] -= tk * v[r] * t;
				}
			}
			// column norm down-dating
			if(is_col && pos > k && cnu != 0){
				double t = fabs(a[k]) / cnu;
				t = (1 + t)*(1 - t);
				t = t < 0 ? 0 : t;
				double rr = cnu / cnd;
				double t2 = t * (rr * rr);
				if(t2 <= norm_downdate_threshold){
					double s = 0;
#pragma unroll
					for(int r = k + 1; r < ROWS; ++r) s += a[r] * a[r];
					cnd = sqrt(s); cnu = cnd;
				} else{
					cnu *= sqrt(t);
				}
			}
		}
#pragma unroll
		for(int i = 0; i < SIZE; ++i) lane_at_pos[i] = __ffs(__ballot_sync(FULL_MASK, is_col && pos == i)) - 1;
	}

	// ---- the same factorisation tuned for the per-pass S x S solve, where it sits on every patch's critical path:
	// pivoting and the rank test work on SQUARED column norms recomputed from the trailing rows (no square roots, no
	// divisions, no down-dating bookkeeping; for 8 rows recomputing is cheaper and more accurate than down-dating),
	// sums are pairwise trees instead of sequential chains, and every division by a pivot is a multiplication by a
	// reciprocal taken once.  Same algorithm and rank rule as factor(); results differ from it by rounding only.
	double rdiag;            // 1 / R(k, k) on the lane whose column became pivot k (factor_fast)

	template<int FROM> __device__ __forceinline__ double tail_sumsq() const{
		double p[ROWS];
#pragma unroll
		for(int r = FROM; r < ROWS; ++r) p[r] = a[r] * a[r];
#pragma unroll
		for(int w = 1; w < ROWS; w *= 2){
#pragma unroll
			for(int r = FROM; r + w < ROWS; r += 2 * w) p[r] += p[r + w];
		}
		return FROM < ROWS ? p[FROM < ROWS ? FROM : 0] : 0.0;
	}
	template<int FROM> __device__ __forceinline__ double tail_dot(const double (&v)[ROWS]) const{
		double p[ROWS];
#pragma unroll
		for(int r = FROM; r < ROWS; ++r) p[r] = v[r] * a[r];
#pragma unroll
		for(int w = 1; w < ROWS; w *= 2){
#pragma unroll
			for(int r = FROM; r + w < ROWS; r += 2 * w) p[r] += p[r + w];
		}
		return FROM < ROWS ? p[FROM < ROWS ? FROM : 0] : 0.0;
	}
	template<int K, bool APPROX> __device__ __forceinline__ void fast_step(int lane, bool is_col, bool is_rhs, double &csq, double thsq){
		const bool cand_ok = is_col && pos >= K && (csq == csq);
		const unsigned hi = cand_ok ? (unsigned)__double2hiint(csq) : 0u;
		const unsigned mhi = __reduce_max_sync(FULL_MASK, hi);
		const bool c1 = cand_ok && hi == mhi;
		const unsigned lo = c1 ? (unsigned)__double2loint(csq) : 0u;
		const unsigned mlo = __reduce_max_sync(FULL_MASK, lo);
		const bool c2 = c1 && lo == mlo;
		int biggest = (int)__reduce_min_sync(FULL_MASK, c2 ? (unsigned)pos : 0xffffu);
		double bsq = __hiloint2double((int)mhi, (int)mlo);
		if(biggest == 0xffff){ biggest = K; bsq = 0; }
		if(nonzero_pivots == SIZE && bsq < thsq * double(ROWS - K)) nonzero_pivots = K;
		if(is_col){ if(pos == biggest) pos = K; else if(pos == K) pos = biggest; }
		const int piv_lane = __ffs(__ballot_sync(FULL_MASK, is_col && pos == K)) - 1;
		double tau_l = 0;
		if(lane == piv_lane){
			const double c0 = a[K], tail_sq = tail_sumsq<K + 1>();
			double beta;
			if(tail_sq <= DBL_MIN){
				tau_l = 0; beta = c0;
#pragma unroll
				for(int r = K + 1; r < ROWS; ++r) a[r] = 0;
			} else if(APPROX){
				// the same quantities from one reciprocal square root and one reciprocal, both Newton-refined fp32 seeds
				const double nrm2 = fma(c0, c0, tail_sq);
				const double rs = rsqrt_newton(nrm2);
				double nrm = nrm2 * rs;
				nrm = fma(0.5 * rs, fma(-nrm, nrm, nrm2), nrm);
				beta = c0 >= 0 ? -nrm : nrm;
				const double rden = rcp_newton_scaled(c0 - beta);
#pragma unroll
				for(int r = K + 1; r < ROWS; ++r) a[r] = a[r] * rden;
				rdiag = c0 >= 0 ? -rs : rs;
			} else{
				beta = sqrt(c0 * c0 + tail_sq);
				if(c0 >= 0) beta = -beta;
				const double rden = ieee_rcp(c0 - beta);
#pragma unroll
				for(int r = K + 1; r < ROWS; ++r) a[r] = a[r] * rden;
			}
			if(!APPROX || tail_sq <= DBL_MIN) rdiag = ieee_rcp(beta);
			if(tail_sq > DBL_MIN) tau_l = (beta - c0) * rdiag;
			a[K] = beta;
		}
		const double tk = __shfl_sync(FULL_MASK, tau_l, piv_lane);
		tau[K] = tk;
		double v[ROWS];
#pragma unroll
		for(int r = K + 1; r < ROWS; ++r) v[r] = __shfl_sync(FULL_MASK, a[r], piv_lane);
		const bool apply = (is_col && pos > K) || (is_rhs && K < nonzero_pivots);
		if(apply){
			if(ROWS - K == 1){
				a[K] *= (1 - tk);
			} else if(tk != 0){
				const double t = tail_dot<K + 1>(v) + a[K];
				const double tt = tk * t;
				a[K] -= tt;
#pragma unroll
				for(int r = K + 1; r < ROWS; ++r) a[r] = fma(-tt, v[r], a[r]);
			}
		}
		if(is_col && pos > K) csq = tail_sumsq<K + 1>();
	}
	template<int K, bool APPROX> __device__ __forceinline__ void fast_steps(int lane, bool is_col, bool is_rhs, double &csq, double thsq){
		fast_step<K, APPROX>(lane, is_col, is_rhs, csq, thsq);
		if(K + 1 < SIZE) fast_steps<(K + 1 < SIZE ? K + 1 : K), APPROX>(lane, is_col, is_rhs, csq, thsq);
	}
	template<bool APPROX = false> __device__ __forceinline__ void factor_fast(int lane, bool with_rhs){
		const bool is_col = lane < COLS;
		const bool is_rhs = with_rhs && (lane == COLS);
		pos = is_col ? lane : -1;
		rdiag = 0;
		double csq = tail_sumsq<0>();
		double maxsq;
		{
			const bool ok = is_col && (csq == csq);
			const unsigned hi = ok ? (unsigned)__double2hiint(csq) : 0u;
			const unsigned mhi = __reduce_max_sync(FULL_MASK, hi);
			const unsigned lo = (ok && hi == mhi) ? (unsigned)__double2loint(csq) : 0u;
			const unsigned mlo = __reduce_max_sync(FULL_MASK, lo);
			maxsq = __hiloint2double((int)mhi, (int)mlo);
		}
		// Eigen's threshold_helper = (max norm * eps)^2 / rows, on the squared norms this factorisation keeps
		const double thsq = maxsq * (DBL_EPSILON * DBL_EPSILON) / double(ROWS);
		nonzero_pivots = SIZE;
		fast_steps<0, APPROX>(lane, is_col, is_rhs, csq, thsq);
#pragma unroll
		for(int i = 0; i < SIZE; ++i) lane_at_pos[i] = __ffs(__ballot_sync(FULL_MASK, is_col && pos == i)) - 1;
	}
	// ---- the factorisation for the solve of the F32 precision (lk_f32.cuh solve_reference_warp): the same column-pivoted
	// Householder QR and rank rule, with the reflector of step k applied in its un-normalised form
	//     H_k = I - g w w^T,   w = (c0 - beta, a[k+1..]),   g = 1 / (beta (beta - c0)) = 1 / (|a_k|^2 + |c0| |a_k|)
	// (identical to Eigen's I - tau v v^T with v = w / w_k, tau = g w_k^2): one reciprocal instead of a square root, a
	// reciprocal and a division; the tail of the pivot column is broadcast as it is (before the scalar chain finishes), and
	// the dot products w[k+1..] . a_j[k+1..] do not wait for beta either.  Reciprocal square root and reciprocal come from
	// Newton-refined fp32 seeds (rsqrt_newton / rcp_newton_scaled).  Results differ from factor_fast<false>() by rounding
	// only (a few ulp of each R entry); rdiag = 1 / R(k, k) is kept for solve_fast_cols().
	template<int K> __device__ __forceinline__ void lean_step(int lane, bool is_col, bool is_rhs, double &csq, double thsq){
		// pivot: the largest squared tail norm among the columns at positions >= K; ties -> the lowest position, like
		// Eigen's left-to-right scan
		const bool cand_ok = is_col && pos >= K && (csq == csq);
		const unsigned hi = cand_ok ? (unsigned)__double2hiint(csq) : 0u;
		const unsigned mhi = __reduce_max_sync(FULL_MASK, hi);
		const bool c1 = cand_ok && hi == mhi;
		const unsigned m1 = __ballot_sync(FULL_MASK, c1);
		int piv_lane;
		double bsq;
		if(__popc(m1) == 1){
			piv_lane = __ffs(m1) - 1;
			bsq = __shfl_sync(FULL_MASK, csq, piv_lane);
		} else if(m1 == 0u){
			piv_lane = __ffs(__ballot_sync(FULL_MASK, is_col && pos == K)) - 1;            // all-NaN norms: keep column K
			bsq = 0;
		} else{
			const unsigned lo = c1 ? (unsigned)__double2loint(csq) : 0u;
			const unsigned mlo = __reduce_max_sync(FULL_MASK, lo);
			const bool c2 = c1 && lo == mlo;
			const int biggest = (int)__reduce_min_sync(FULL_MASK, c2 ? (unsigned)pos : 0xffffu);
			piv_lane = __ffs(__ballot_sync(FULL_MASK, c2 && pos == biggest)) - 1;
			bsq = __hiloint2double((int)mhi, (int)mlo);
		}
		if(nonzero_pivots == SIZE && bsq < thsq * double(ROWS - K)) nonzero_pivots = K;
		{
			// the pivot column moves to position K, the column that was there takes its place
			const int piv_pos = __shfl_sync(FULL_MASK, pos, piv_lane);
			if(is_col){ if(lane == piv_lane) pos = K; else if(pos == K) pos = piv_pos; }
		}
		// the pivot column's entries from row K down, as they are
		double w[ROWS];
#pragma unroll
		for(int r = K; r < ROWS; ++r) w[r] = __shfl_sync(FULL_MASK, a[r], piv_lane);
		const double c0 = w[K];
		// (uniform: every lane evaluates the scalars of the reflector from the broadcast column -- no second round trip)
		double tail_sq;
		{
			double p[ROWS];
#pragma unroll
			for(int r = K + 1; r < ROWS; ++r) p[r] = w[r] * w[r];
#pragma unroll
			for(int wd = 1; wd < ROWS; wd *= 2){
#pragma unroll
				for(int r = K + 1; r + wd < ROWS; r += 2 * wd) p[r] += p[r + wd];
			}
			tail_sq = (K + 1 < ROWS) ? p[K + 1 < ROWS ? K + 1 : 0] : 0.0;
		}
		// partial dot product with this lane's column (or the rhs): independent of the scalar chain below
		const bool apply = (is_col && pos > K) || (is_rhs && K < nonzero_pivots);
		double dot = 0;
		if(K + 1 < ROWS) dot = tail_dot<K + 1>(w);
		double beta, g, wk;
		if(tail_sq <= DBL_MIN){
			beta = c0; g = 0; wk = 0;                                                        // tau = 0: the identity
		} else{
			const double nrm2 = fma(c0, c0, tail_sq);
			const int hi2 = __double2hiint(nrm2);
			double rs, nrm;
			if((unsigned)(hi2 - 0x08000000) < 0x70000000u){           // 2^-895 < |a_k|^2 < 2^897: the seeds and q below stay normal
				// the reciprocal's seed is taken from the reciprocal square root's seed, so its MUFU overlaps the Newton steps
				const double r0 = rsqrt_seed(nrm2);
				double g0 = rcp_seed(fma(fabs(c0), nrm2 * r0, nrm2));
				double sr = nrm2 * r0;
				rs = fma(0.5 * r0, fma(-sr, r0, 1.0), r0);
				sr = nrm2 * rs;
				rs = fma(0.5 * rs, fma(-sr, rs, 1.0), rs);
				nrm = nrm2 * rs;
				const double q = fma(fabs(c0), nrm, nrm2);
				g0 = fma(fma(-q, g0, 1.0), g0, g0);
				g = fma(fma(-q, g0, 1.0), g0, g0);
			} else{
				rs = rsqrt_newton(nrm2);
				nrm = nrm2 * rs;
				g = rcp_newton_scaled(fma(fabs(c0), nrm, nrm2));
			}
			beta = c0 >= 0 ? -nrm : nrm;
			wk = c0 - beta;
			if(lane == piv_lane) rdiag = c0 >= 0 ? -rs : rs;
		}
		if(lane == piv_lane){
			a[K] = beta;
			if(tail_sq <= DBL_MIN) rdiag = ieee_rcp(beta);
		}
		if(apply && g != 0){
			const double t = fma(wk, a[K], dot);
			const double tt = g * t;
			a[K] = fma(-tt, wk, a[K]);
#pragma unroll
			for(int r = K + 1; r < ROWS; ++r) a[r] = fma(-tt, w[r], a[r]);
		}
		if(is_col && pos > K) csq = tail_sumsq<K + 1>();
	}
	template<int K> __device__ __forceinline__ void lean_steps(int lane, bool is_col, bool is_rhs, double &csq, double thsq){
		lean_step<K>(lane, is_col, is_rhs, csq, thsq);
		if(K + 1 < SIZE) lean_steps<(K + 1 < SIZE ? K + 1 : K)>(lane, is_col, is_rhs, csq, thsq);
	}
	__device__ __forceinline__ void factor_lean(int lane, bool with_rhs){
		const bool is_col = lane < COLS;
		const bool is_rhs = with_rhs && (lane == COLS);
		pos = is_col ? lane : -1;
		rdiag = 0;
		double csq = tail_sumsq<0>();
		double maxsq;
		{
			const bool ok = is_col && (csq == csq);
			const unsigned hi = ok ? (unsigned)__double2hiint(csq) : 0u;
			const unsigned mhi = __reduce_max_sync(FULL_MASK, hi);
			const unsigned lo = (ok && hi == mhi) ? (unsigned)__double2loint(csq) : 0u;
			const unsigned mlo = __reduce_max_sync(FULL_MASK, lo);
			maxsq = __hiloint2double((int)mhi, (int)mlo);
		}
		const double thsq = maxsq * (DBL_EPSILON * DBL_EPSILON) / double(ROWS);
		nonzero_pivots = SIZE;
		lean_steps<0>(lane, is_col, is_rhs, csq, thsq);
#pragma unroll
		for(int i = 0; i < SIZE; ++i) lane_at_pos[i] = __ffs(__ballot_sync(FULL_MASK, is_col && pos == i)) - 1;
	}
	__device__ __forceinline__ double solve_fast(int lane){
		static_assert(ROWS == COLS, "solve_fast() is written for the square systems of the LK loop");
		const int np = nonzero_pivots;
#pragma unroll
		for(int i = SIZE - 1; i >= 0; --i){
			double s = a[i];
#pragma unroll
			for(int j = i + 1; j < SIZE; ++j){
				const double rij = __shfl_sync(FULL_MASK, a[i], lane_at_pos[j]);
				if(j < np) s = fma(-rij, a[j], s);
			}
			const double rinv = __shfl_sync(FULL_MASK, rdiag, lane_at_pos[i]);
			if(lane == COLS && i < np) a[i] = s * rinv;
		}
		double x = 0;
#pragma unroll
		for(int i = 0; i < SIZE; ++i){
			const double ci = __shfl_sync(FULL_MASK, a[i], COLS);
			if(pos == i) x = (i < np) ? ci : 0.0;
		}
		return x;
	}

	// The same back-substitution column by column: x_i = c_i / R_ii, then c_j -= R_ji x_i for j < i.  Column i of R sits in
	// one lane, so its entries come from shuffles that do not depend on x (they overlap the previous step), and the
	// dependent chain is one multiply and one fused multiply-add per step instead of a row's worth.
	__device__ __forceinline__ double solve_fast_cols(int lane){
		static_assert(ROWS == COLS, "solve_fast_cols() is written for the square systems of the LK loop");
		const int np = nonzero_pivots;
#pragma unroll
		for(int i = SIZE - 1; i >= 0; --i){
			const double rinv = __shfl_sync(FULL_MASK, rdiag, lane_at_pos[i]);
			double rc[SIZE];
#pragma unroll
			for(int j = 0; j < i; ++j) rc[j] = __shfl_sync(FULL_MASK, a[j], lane_at_pos[i]);
			const double xi = (i < np) ? a[i] * rinv : 0.0;          // meaningful on the rhs lane only
			if(lane == COLS){
				a[i] = xi;
#pragma unroll
				for(int j = 0; j < i; ++j) a[j] = fma(-rc[j], xi, a[j]);
			}
		}
		double x = 0;
#pragma unroll
		for(int i = 0; i < SIZE; ++i){
			const double ci = __shfl_sync(FULL_MASK, a[i], COLS);
			if(pos == i) x = ci;
		}
		return x;
	}

	// back-substitution on the rhs lane, then x[perm[i]] = c[i]: returns x[lane] on lanes < COLS
	__device__ __forceinline__ double solve(int lane){
		static_assert(ROWS == COLS, "solve() is written for the square systems of the LK loop");
		const int np = nonzero_pivots;
#pragma unroll
		for(int i = SIZE - 1; i >= 0; --i){
			double s = a[i];
#pragma unroll
			for(int j = i + 1; j < SIZE; ++j){
				double rij = __shfl_sync(FULL_MASK, a[i], lane_at_pos[j]);
				if(j < np) s -= rij * a[j];
			}
			double rii = __shfl_sync(FULL_MASK, a[i], lane_at_pos[i]);
			if(lane == COLS && i < np) a[i] = s / rii;
		}
		double x = 0;
#pragma unroll
		for(int i = 0; i < SIZE; ++i){
			double ci = __shfl_sync(FULL_MASK, a[i], COLS);
			if(pos == i) x = (i < np) ? ci : 0.0;
		}
		return x;
	}
};

// utils::computeHomographyDLT(in_corners, out_corners), warpUtils.cc:171-223: the null vector of the 8x9
// constraint matrix taken as JacobiSVD(ComputeFullV).matrixV().col(8).  For a wide matrix Eigen's QR
// preconditioner makes that column the last column of Q of the column-pivoted Householder QR of the
// adjoint (9x8), untouched by the Jacobi sweeps; that is what is computed here, by one warp.
// corners: x0..x3, y0..y3.  Result (uniform across the warp) is normalised by its last entry.
__device__ __forceinline__ Mat3 warp_homography_dlt(const double *in_c, const double *out_c, int lane){
	WarpColPivQR<9, 8> qr;
	{
		const int i = (lane & 7) >> 1;
		const double ix = in_c[i], iy = in_c[4 + i], ox = out_c[i], oy = out_c[4 + i];
		if((lane & 1) == 0){
			qr.a[0] = 0; qr.a[1] = 0; qr.a[2] = 0; qr.a[3] = -ix; qr.a[4] = -iy; qr.a[5] = -1;
			qr.a[6] = oy*ix; qr.a[7] = oy*iy; qr.a[8] = oy;
		} else{
			qr.a[0] = ix; qr.a[1] = iy; qr.a[2] = 1; qr.a[3] = 0; qr.a[4] = 0; qr.a[5] = 0;
			qr.a[6] = -ox*ix; qr.a[7] = -ox*iy; qr.a[8] = -ox;
		}
	}
	// JacobiSVD::compute scales by the largest absolute coefficient first
	double scale = 0;
#pragma unroll
	for(int r = 0; r < 9; ++r) scale = fmax(scale, fabs(qr.a[r]));
	if(lane >= 8) scale = 0;
#pragma unroll
	for(int off = 4; off >= 1; off >>= 1) scale = fmax(scale, __shfl_xor_sync(FULL_MASK, scale, off));
	scale = __shfl_sync(FULL_MASK, scale, 0);
	if(scale == 0) scale = 1;
#pragma unroll
	for(int r = 0; r < 9; ++r) qr.a[r] = qr.a[r] / scale;
	qr.factor(lane, false);
	// Q e_9 = H_0 ... H_7 e_9
	double h[9];
#pragma unroll
	for(int r = 0; r < 9; ++r) h[r] = 0;
	h[8] = 1;
#pragma unroll
	for(int k = 7; k >= 0; --k){
		double v[9];
#pragma unroll
		for(int r = k + 1; r < 9; ++r) v[r] = __shfl_sync(FULL_MASK, qr.a[r], qr.lane_at_pos[k]);
		const double tk = qr.tau[k];
		if(tk != 0){
			double t = 0;
#pragma unroll
			for(int r = k + 1; r < 9; ++r) t += v[r] * h[r];
			t += h[k];
			h[k] -= tk * t;
#pragma unroll
			for(int r = k + 1; r < 9; ++r) h[r] -= tk * v[r] * t;
		}
	}
	Mat3 H;
#pragma unroll
	for(int r = 0; r < 9; ++r) H.m[r] = h[r] / h[8];
	return H;
}

} // namespace mtfb
