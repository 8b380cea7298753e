// lk_solve.cuh -- the appearance-model-independent tail of one Gauss-Newton pass, run by warp 0 of the CTA:
// Levenberg-Marquardt accept / reject, Hessian selection and damping, the S x S column-pivoted QR solve,
// the (inverse) compositional update, the corner-change stopping test and the iteration log.
// Control flow follows nt::FCLK::update (SM/src/NT/FCLK.cc:187-352), nt::ESM::update (NT/ESM.cc:186-296) and
// nt::ICLK::update (NT/ICLK.cc:167-297); the templated twins differ only in how a rejected step is counted.
#pragma once
#include "lk_common.cuh"

namespace mtfb {

struct LMState {
	double prev_similarity, lm_delta;
	double ssm_update;          // lane l holds entry l of the last state update
	bool state_reset;
};

// which Hessian the solve uses: 0 = the one the pixel pass just produced, 1 = the one stored at
// initialize() (init_self_hessian), 2 = the mean of both (ESM SumOfSelf / SumOfStd, NT/ESM.cc:339-352)
template<int SM> __device__ __forceinline__ int hessian_select(int hess_type){
	if(SM == SM_ESM){
		return (hess_type == MTFB_ESM_HESS_INITIAL_SELF) ? 1 :
			(hess_type == MTFB_ESM_HESS_SUM_OF_SELF || hess_type == MTFB_ESM_HESS_SUM_OF_STD) ? 2 : 0;
	}
	if(SM == SM_FCLK || SM == SM_FALK || SM == SM_IALK) return (hess_type == MTFB_LK_HESS_INITIAL_SELF) ? 1 : 0;
	return (hess_type == MTFB_LK_HESS_CURRENT_SELF) ? 0 : 1;
}

// f: similarity of this pass.  s_J[S]: df_dp as the search method uses it (ESM's 0.5 already applied).
// s_Hc[S*S]: column-major Hessian of this pass (ignored when hsel == 1).
// Writes s_W, s_corners, the log; returns CTRL_*.  All 32 lanes of warp 0 must call it.
// PRESOLVED: the caller has solved for the state update already (s_dp[S]; lk_ssd_f32.cu solves in its own basis);
// only valid without Levenberg-Marquardt.  s_J / s_Hc are then used for the log alone (may be null without a log).
// HSEL >= 0 overrides hessian_select (kernels that assemble the complete Hessian of the pass themselves).
template<int SSM, int SM, bool PRESOLVED = false, int HSEL = -1>
__device__ __forceinline__ int serial_step(const DevBatch &b, int p, int lane, int iter_id, int n_passes, double f,
	const double *s_J, const double *s_Hc, double *s_W, double *s_corners, const double *s_init_corners,
	LMState &lm, int &patch_status, const double *s_dp = nullptr, double *s_state = nullptr){
	constexpr int S = StateSize<SSM>::value;
	// the additive searches (nt::FALK / nt::IALK) keep the STATE and add to it (ssm->additiveUpdate, ProjectiveBase.cc:51-55:
	// curr_state += update; setState(curr_state)); the compositional ones keep the warp
	constexpr bool ADDITIVE = (SM == SM_FALK || SM == SM_IALK);
#if MTFB_PROF
	const long long pt0 = clock64();
#endif
	int ctrl = CTRL_NEXT;
	bool rejected = false;
	Mat3 W;
#pragma unroll
	for(int i = 0; i < 9; ++i) W.m[i] = s_W[i];
	Mat3 Wn = W;
	double upd_norm = 0, x = 0, Jv = 0;
	double dp[S];
	if(b.leven_marq && !lm.state_reset){
		if(iter_id > 0){
			if(f < lm.prev_similarity){
				lm.lm_delta *= b.lm_delta_update;
#pragma unroll
				for(int s = 0; s < S; ++s) dp[s] = __shfl_sync(FULL_MASK, lm.ssm_update, s);
				if(ADDITIVE){
					// ssm->additiveUpdate(-ssm_update) (NT/FALK.cc:150-152, NT/IALK.cc:110-112)
					double ns[S];
#pragma unroll
					for(int s = 0; s < S; ++s) ns[s] = s_state[s] + (-dp[s]);
					Wn = warp_from_state<SSM>(ns);
					__syncwarp();
#pragma unroll
					for(int s = 0; s < S; ++s) if(lane == s) s_state[s] = ns[s];
				} else if(SM == SM_ICLK){
					// undo the inverse step by re-applying the forward update (NT/ICLK.cc:183)
					Wn = compose_update<SSM>(W, dp);
				} else{
					double inv[S];
					invert_state<SSM>(inv, dp);
					Wn = compose_update<SSM>(W, inv);
				}
				lm.state_reset = true; rejected = true; ctrl = CTRL_REJECT;
			} else if(f > lm.prev_similarity){
				lm.lm_delta /= b.lm_delta_update;
			}
		}
		if(!rejected) lm.prev_similarity = f;
	}
	if(PRESOLVED){
		x = lane < S ? s_dp[lane] : 0.0;
#pragma unroll
		for(int s = 0; s < S; ++s) dp[s] = s_dp[s];
		if(SM == SM_ICLK){
			double inv[S];
			invert_state<SSM>(inv, dp);                          // NT/ICLK.cc:270-271
			Wn = compose_update<SSM>(W, inv);
		} else{
			Wn = compose_update<SSM>(W, dp);
		}
	} else if(!rejected){
		lm.state_reset = false;
		WarpColPivQR<S, S> qr;
		const int hsel = HSEL >= 0 ? HSEL : hessian_select<SM>(b.hess_type);
		const int jc = lane < S ? lane : 0;                      // column `lane` of the Hessian
#pragma unroll
		for(int i = 0; i < S; ++i){
			const double hc = hsel != 1 ? s_Hc[jc*S + i] : 0.0;
			const double hi0 = hsel != 0 ? b.Hinit[(size_t)p * 64 + jc*S + i] : 0.0;
			qr.a[i] = hsel == 0 ? hc : (hsel == 1 ? hi0 : (hc + hi0) * 0.5);
		}
		if(lane == S){
#pragma unroll
			for(int i = 0; i < S; ++i) qr.a[i] = s_J[i];
		}
		if(lane < S) Jv = s_J[lane];
		if(b.leven_marq){
#pragma unroll
			for(int i = 0; i < S; ++i) if(i == lane) qr.a[i] += lm.lm_delta * qr.a[i];      // NT/FCLK.cc:289-296
		}
		if(b.log && n_passes <= b.log_slots && lane < S){
			mtfb_iter_log *e = b.log + (size_t)p*b.log_slots + (n_passes - 1);
#pragma unroll
			for(int i = 0; i < S; ++i) e->hessian[lane*S + i] = qr.a[i];
		}
#if MTFB_PROF
		const long long pt1 = clock64();
#endif
		qr.factor_fast(lane, true);
#if MTFB_PROF
		const long long pt2 = clock64();
#endif
		x = -qr.solve_fast(lane);                                     // state_update = -H^-1 J^T (NT/FCLK.cc:298)
		if(qr.nonzero_pivots < S) patch_status |= MTFB_PATCH_SINGULAR;
#if MTFB_PROF
		const long long pt3 = clock64();
		if(lane == 0){ atomicAdd((unsigned long long*)b.n_iters_prof + 4, (unsigned long long)(pt1 - pt0)); atomicAdd((unsigned long long*)b.n_iters_prof + 5, (unsigned long long)(pt2 - pt1));
			atomicAdd((unsigned long long*)b.n_iters_prof + 6, (unsigned long long)(pt3 - pt2)); }
#endif
		lm.ssm_update = x;
#pragma unroll
		for(int s = 0; s < S; ++s) dp[s] = __shfl_sync(FULL_MASK, x, s);
		if(ADDITIVE){
			double ns[S];
#pragma unroll
			for(int s = 0; s < S; ++s) ns[s] = s_state[s] + dp[s];
			Wn = warp_from_state<SSM>(ns);
			__syncwarp();
#pragma unroll
			for(int s = 0; s < S; ++s) if(lane == s) s_state[s] = ns[s];
		} else if(SM == SM_ICLK){
			double inv[S];
			invert_state<SSM>(inv, dp);                          // NT/ICLK.cc:270-271
			Wn = compose_update<SSM>(W, inv);
		} else{
			Wn = compose_update<SSM>(W, dp);
		}
	}
	double nc[8];
	warp_corners<SSM>(Wn, s_init_corners, nc);
	if(!rejected){
#pragma unroll
		for(int i = 0; i < 8; ++i){ double d = s_corners[i] - nc[i]; upd_norm += d*d; }
		if(upd_norm < b.epsilon) ctrl = CTRL_BREAK;
		if(!(upd_norm == upd_norm) || !(f == f)) patch_status |= MTFB_PATCH_NAN;
	}
	__syncwarp();
	// (unrolled: a lane-indexed read of a register array would go through local memory)
#pragma unroll
	for(int i = 0; i < 9; ++i) if(lane == i) s_W[i] = Wn.m[i];
#pragma unroll
	for(int i = 0; i < 8; ++i) if(lane == i) s_corners[i] = nc[i];
	if(b.log && n_passes <= b.log_slots){
		mtfb_iter_log *e = b.log + (size_t)p*b.log_slots + (n_passes - 1);
		if(lane < S){ e->jacobian[lane] = rejected ? 0.0 : Jv; e->state_update[lane] = rejected ? 0.0 : x; }
		if(lane < 8) e->corners[lane] = nc[lane];
		if(lane == 0){ e->f = f; e->update_norm = upd_norm; e->rejected = rejected; e->valid = 1; }
	}
#if MTFB_PROF
	if(lane == 0) atomicAdd((unsigned long long*)b.n_iters_prof + 7, (unsigned long long)(clock64() - pt0));
#endif
	return ctrl;
}

// nt::FCLK re-enters its while loop without counting a rejected step (NT/FCLK.cc:187,210); every other loop
// is a for(...; ++iter_id) (FCLK.cc:117,135, ESM.cc:128, NT/ESM.cc:186, ICLK.cc:150, NT/ICLK.cc:167)
template<int SM> __device__ __forceinline__ bool counts_as_iteration(int ctrl, int nt_semantics){
	return !(ctrl == CTRL_REJECT && SM == SM_FCLK && nt_semantics);
}

// final state of the patch -> global memory (warp 0)
template<int SSM> __device__ __forceinline__ void store_patch_state(const DevBatch &b, int p, int lane, const double *s_W,
	const double *s_corners, double f, int n_passes, int patch_status, const double *s_state = nullptr){
	constexpr int S = StateSize<SSM>::value;
	Mat3 W;
#pragma unroll
	for(int i = 0; i < 9; ++i) W.m[i] = s_W[i];
	double st[S];
	state_from_warp<SSM>(st, W);
	if(s_state){                                  // additive searches: the state itself is what was tracked
#pragma unroll
		for(int s = 0; s < S; ++s) st[s] = s_state[s];
	}
	if(lane < 9) b.warp[(size_t)p * 9 + lane] = W.m[lane];
	if(lane < 8){
		const double c = s_corners[lane];
		b.corners[(size_t)p * 8 + lane] = c;
		// the all-gather, fused: the same 64 bytes into the gathered array of every rank of the job (peer memory, NVLink)
		for(int r = 0; r < b.peers.n; ++r) b.peers.dst[r][(size_t)(b.peers.row0 + p) * 8 + lane] = c;
	}
#pragma unroll
	for(int s = 0; s < S; ++s) if(lane == s) b.state[(size_t)p*S + s] = st[s];
	if(lane == 0){ b.f[p] = f; b.n_iters[p] = n_passes; b.status[p] = patch_status; }
}

// utils::computeAffineNDLT(in_corners, out_corners), warpUtils.cc:378-386: normalizePts(out) (warpUtils.cc:802-832), the
// least-squares affine map of computeAffineDLT (warpUtils.cc:276-343: JacobiSVD pseudo-inverse of the 8 x 6 system, which
// is block diagonal with the 4 x 3 block P = [x y 1] for both rows, full column rank), inv_norm_mat * warp.  Householder QR
// least squares instead of Eigen's Jacobi SVD: the two agree to rounding (1e-15 relative; tolerance stated in the tests).
// Evaluated by every lane alike (uniform result).
__device__ __forceinline__ Mat3 affine_ndlt(const double *in_c, const double *out_c){
	double cx = 0, cy = 0;
#pragma unroll
	for(int i = 0; i < 4; ++i){ cx += out_c[i]; cy += out_c[4 + i]; }
	cx /= 4; cy /= 4;
	double A[4][3], c0[4], c1[4], mean_dist = 0;
#pragma unroll
	for(int i = 0; i < 4; ++i){
		c0[i] = out_c[i] - cx; c1[i] = out_c[4 + i] - cy;
		mean_dist += sqrt(c0[i] * c0[i] + c1[i] * c1[i]);
		A[i][0] = in_c[i]; A[i][1] = in_c[4 + i]; A[i][2] = 1;
	}
	mean_dist /= 4;
	const double norm_scale = sqrt(2.0) / mean_dist;
#pragma unroll
	for(int i = 0; i < 4; ++i){ c0[i] *= norm_scale; c1[i] *= norm_scale; }
#pragma unroll
	for(int k = 0; k < 3; ++k){
		double nrm = 0;
#pragma unroll
		for(int i = k; i < 4; ++i) nrm += A[i][k] * A[i][k];
		nrm = sqrt(nrm);
		const double alpha = A[k][k] >= 0 ? -nrm : nrm;
		double v[4] = { 0, 0, 0, 0 };
#pragma unroll
		for(int i = k; i < 4; ++i) v[i] = A[i][k];
		v[k] -= alpha;
		double vv = 0;
#pragma unroll
		for(int i = k; i < 4; ++i) vv += v[i] * v[i];
		if(nrm == 0 || vv == 0) continue;
#pragma unroll
		for(int j = k; j < 3; ++j){
			double d = 0;
#pragma unroll
			for(int i = k; i < 4; ++i) d += v[i] * A[i][j];
			d = 2 * d / vv;
#pragma unroll
			for(int i = k; i < 4; ++i) A[i][j] -= d * v[i];
		}
		double d0 = 0, d1 = 0;
#pragma unroll
		for(int i = k; i < 4; ++i){ d0 += v[i] * c0[i]; d1 += v[i] * c1[i]; }
		d0 = 2 * d0 / vv; d1 = 2 * d1 / vv;
#pragma unroll
		for(int i = k; i < 4; ++i){ c0[i] -= d0 * v[i]; c1[i] -= d1 * v[i]; }
	}
	double a0[3], a1[3];
#pragma unroll
	for(int k = 2; k >= 0; --k){
		double s0 = c0[k], s1 = c1[k];
#pragma unroll
		for(int j = k + 1; j < 3; ++j){ s0 -= A[k][j] * a0[j]; s1 -= A[k][j] * a1[j]; }
		a0[k] = s0 / A[k][k]; a1[k] = s1 / A[k][k];
	}
	// inv_norm_mat * affine_mat, entry by entry in Eigen's order (k ascending)
	const double is = 1.0 / norm_scale;
	Mat3 W;
	W.m[0] = is*a0[0] + 0.0*a1[0] + cx*0.0; W.m[1] = is*a0[1] + 0.0*a1[1] + cx*0.0; W.m[2] = is*a0[2] + 0.0*a1[2] + cx*1.0;
	W.m[3] = 0.0*a0[0] + is*a1[0] + cy*0.0; W.m[4] = 0.0*a0[1] + is*a1[1] + cy*0.0; W.m[5] = 0.0*a0[2] + is*a1[2] + cy*1.0;
	W.m[6] = 0; W.m[7] = 0; W.m[8] = 1;
	return W;
}

// ssm.setCorners for one patch (warp 0): 4-point DLT, identity warp, zero state
template<int SSM> __device__ __forceinline__ Mat3 set_corners(const DevBatch &b, int p, int lane, const double *c_in){
	constexpr int S = StateSize<SSM>::value;
	if(SSM == SSM_AFF && b.norm_init){
		// Affine::setCorners with normalized_init (Affine.cc:65-74): the template stays the pixel-scaled square, curr_warp =
		// computeAffineNDLT(init_corners, corners), curr_corners = curr_warp . init_corners_hm (NOT the supplied corners: a
		// general quadrilateral is only fitted in the least-squares sense)
		const Mat3 aw = affine_ndlt(b.norm_corners, c_in);
		double st[S];
		state_from_warp<SSM>(st, aw);
		if(lane < 9){ b.warp[(size_t)p * 9 + lane] = aw.m[lane]; b.dlt[(size_t)p * 9 + lane] = aw.m[lane]; }
#pragma unroll
		for(int s = 0; s < S; ++s) if(lane == s) b.state[(size_t)p*S + s] = st[s];
		if(lane < 8){
			const int i = lane & 3, r = lane >> 2;
			const double px = b.norm_corners[i], py = b.norm_corners[4 + i];
			double v = aw.m[3 * r] * px; v = v + aw.m[3 * r + 1] * py; v = v + aw.m[3 * r + 2] * 1.0;
			b.corners[(size_t)p * 8 + lane] = v;
			b.init_corners[(size_t)p * 8 + lane] = b.norm_corners[lane];
		}
		return aw;
	}
	Mat3 dlt = warp_homography_dlt(b.norm_corners, c_in, lane);
	Mat3 I = mat3_identity();
	if(b.norm_init){
		// Homography::setCorners with normalized_init (Homography.cc:57-62): the initial region is the normalised square,
		// curr_warp = the DLT warp, curr_state = getStateFromWarp(curr_warp)
		double st[S];
		state_from_warp<SSM>(st, dlt);
		if(lane < 9) b.warp[(size_t)p * 9 + lane] = dlt.m[lane];
#pragma unroll
		for(int s = 0; s < S; ++s) if(lane == s) b.state[(size_t)p*S + s] = st[s];
		if(lane < 8){ b.corners[(size_t)p * 8 + lane] = c_in[lane]; b.init_corners[(size_t)p * 8 + lane] = b.norm_corners[lane]; }
		if(lane < 9) b.dlt[(size_t)p * 9 + lane] = dlt.m[lane];
		return dlt;
	}
	if(lane < 9){ b.dlt[(size_t)p * 9 + lane] = dlt.m[lane]; b.warp[(size_t)p * 9 + lane] = I.m[lane]; }
	if(lane < S) b.state[(size_t)p*S + lane] = 0;
	if(lane < 8){ b.corners[(size_t)p * 8 + lane] = c_in[lane]; b.init_corners[(size_t)p * 8 + lane] = c_in[lane]; }
	return dlt;
}

} // namespace mtfb
