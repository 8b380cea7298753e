// lk_ssd_terms.cuh -- what one pixel contributes to the SSD sums, shared by the SSD kernels.
#pragma once
#include "lk_solve.cuh"

namespace mtfb {

// what one pixel contributes to the sums: r = I_t - I_0, Jj = the row that multiplies df/dI in the Jacobian,
// Jt = the row whose outer product goes into the Hessian, wj = df/dI
template<int S> struct PixTerms { double r, wj; double Jt[S], Jj[S]; };

template<int SSM, int SM, class MW> __device__ __forceinline__ void pixel_terms(const DevBatch &b, const MW &W, const double *abcd,
	const PixGeom &g, const Sample &smp, double i0, const double *__restrict__ G0, int pix, bool need_grad, bool esm_mean,
	PixTerms<StateSize<SSM>::value> &t){
	constexpr int S = StateSize<SSM>::value;
	t.r = smp.val - i0;                                            // I_diff (SSDBase.cc:78)
	if(SM == SM_FALK || SM == SM_IALK){
		// the additive searches (NT/FALK.cc:170-181, NT/IALK.cc:131-138): am.cmptCurrJacobian(curr_pix_jacobian) with
		// df_dIt = -I_diff; curr_pix_jacobian = ssm.cmptPixJacobian(curr grad) / ssm.cmptApproxPixJacobian(init grad, un-chained)
		t.wj = -t.r;
		if(SM == SM_FALK) additive_pix_jacobian<SSM>(g, smp.gx, smp.gy, t.Jt);
		else approx_pix_jacobian<SSM>(W, abcd, g, __ldcg(b.G0raw + (size_t)(G0 - b.G0) + pix), __ldcg(b.G0raw + (size_t)(G0 - b.G0) + b.N + pix), t.Jt);
#pragma unroll
		for(int s = 0; s < S; ++s) t.Jj[s] = t.Jt[s];
		return;
	}
	if(SM == SM_ICLK){
		// df_dI0 = I_diff (SSDBase.cc:34: I_diff aliases df_dI0); Jacobian of the template
		t.wj = t.r;
		init_pix_jacobian<SSM>(g.ix, g.iy, __ldcg(G0 + pix), __ldcg(G0 + b.N + pix), t.Jj);
		if(need_grad) pixel_jacobian_row<SSM>(b, W, abcd, g, smp.gx, smp.gy, t.Jt);
		return;
	}
	t.wj = -t.r;                                                   // df_dIt = -I_diff (SSDBase.cc:115-121)
	pixel_jacobian_row<SSM>(b, W, abcd, g, smp.gx, smp.gy, t.Jt);
	if(SM == SM_ESM){
		double J0[S];
		init_pix_jacobian<SSM>(g.ix, g.iy, __ldcg(G0 + pix), __ldcg(G0 + b.N + pix), J0);
		if(esm_mean){
			// mean_pix_jacobian = (init + curr) / 2 (NT/ESM.cc:246-248)
#pragma unroll
			for(int s = 0; s < S; ++s) J0[s] = (J0[s] + t.Jt[s]) / 2.0;
		}
		if(b.jac_type == MTFB_ESM_JAC_ORIGINAL){
#pragma unroll
			for(int s = 0; s < S; ++s) t.Jj[s] = J0[s];
		} else{
			// SSDBase::cmptDifferenceOfJacobians: df_dIt * (dI0_dp + dIt_dp) (SSDBase.cc:186)
#pragma unroll
			for(int s = 0; s < S; ++s) t.Jj[s] = esm_mean ? (2.0*J0[s]) : (J0[s] + t.Jt[s]);
		}
		if(b.hess_type == MTFB_ESM_HESS_ORIGINAL){
#pragma unroll
			for(int s = 0; s < S; ++s) t.Jt[s] = J0[s];
		}
	} else{
#pragma unroll
		for(int s = 0; s < S; ++s) t.Jj[s] = t.Jt[s];
	}
}

template<int S> __device__ __forceinline__ void accumulate_terms(double (&acc)[AccLayout<S>::NA], const PixTerms<S> &t, bool with_hessian){
	typedef AccLayout<S> L;
	acc[0] = fma(t.r, t.r, acc[0]);
#pragma unroll
	for(int s = 0; s < S; ++s) acc[1 + s] = fma(t.wj, t.Jj[s], acc[1 + s]);
	if(with_hessian){
#pragma unroll
		for(int i = 0; i < S; ++i){
#pragma unroll
			for(int j = i; j < S; ++j) acc[1 + S + L::tri(i, j)] = fma(t.Jt[i], t.Jt[j], acc[1 + S + L::tri(i, j)]);
		}
	}
}


} // namespace mtfb
