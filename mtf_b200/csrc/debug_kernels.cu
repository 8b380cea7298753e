// debug_kernels.cu -- test entry points that run single device building blocks of the path on caller-supplied
// inputs (mtfb_debug_* in include/mtf_b200.h).  Not on any product path.
#include "lk_common.cuh"

namespace mtfb {

// x = colPivHouseholderQr(A).solve(b) with the warp-level QR of lk_warp.cuh: one warp per system
template<int S, int MODE>
__global__ void debug_qr_solve_kernel(const double *__restrict__ A /* n_sys x S x S, column-major */, const double *__restrict__ rhs,
	double *__restrict__ x, int *__restrict__ nonzero_pivots, int *__restrict__ perm){
	const int sys = blockIdx.x, lane = threadIdx.x;
	WarpColPivQR<S, S> qr;
	const int jc = lane < S ? lane : 0;
#pragma unroll
	for(int i = 0; i < S; ++i) qr.a[i] = (lane == S) ? rhs[(size_t)sys*S + i] : A[(size_t)sys*S*S + jc*S + i];
	double xs;
	if(MODE == 2){ qr.factor_lean(lane, true); xs = qr.solve_fast_cols(lane); }
	else if(MODE == 3){ qr.template factor_fast<true>(lane, true); xs = qr.solve_fast_cols(lane); }
	else if(MODE == 1){ qr.template factor_fast<false>(lane, true); xs = qr.solve_fast(lane); }
	else{ qr.factor(lane, true); xs = qr.solve(lane); }
	if(lane < S){ x[(size_t)sys*S + lane] = xs; if(perm) perm[(size_t)sys*S + qr.pos] = lane; }
	if(lane == 0 && nonzero_pivots) nonzero_pivots[sys] = qr.nonzero_pivots;
}

cudaError_t launch_debug_qr_solve(int n, int fast, int n_sys, const double *d_A, const double *d_b, double *d_x, int *d_nz, int *d_perm,
	cudaStream_t st){
	if(n == 8){
		if(fast == 2) debug_qr_solve_kernel<8, 2><<<n_sys, 32, 0, st>>>(d_A, d_b, d_x, d_nz, d_perm);
		else if(fast == 3) debug_qr_solve_kernel<8, 3><<<n_sys, 32, 0, st>>>(d_A, d_b, d_x, d_nz, d_perm);
		else if(fast) debug_qr_solve_kernel<8, 1><<<n_sys, 32, 0, st>>>(d_A, d_b, d_x, d_nz, d_perm);
		else debug_qr_solve_kernel<8, 0><<<n_sys, 32, 0, st>>>(d_A, d_b, d_x, d_nz, d_perm);
	} else if(n == 6){
		if(fast == 2) debug_qr_solve_kernel<6, 2><<<n_sys, 32, 0, st>>>(d_A, d_b, d_x, d_nz, d_perm);
		else if(fast == 3) debug_qr_solve_kernel<6, 3><<<n_sys, 32, 0, st>>>(d_A, d_b, d_x, d_nz, d_perm);
		else if(fast) debug_qr_solve_kernel<6, 1><<<n_sys, 32, 0, st>>>(d_A, d_b, d_x, d_nz, d_perm);
		else debug_qr_solve_kernel<6, 0><<<n_sys, 32, 0, st>>>(d_A, d_b, d_x, d_nz, d_perm);
	} else return cudaErrorInvalidValue;
	return cudaGetLastError();
}

} // namespace mtfb
