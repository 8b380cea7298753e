// lk_common.cuh -- device-side pieces shared by the kernel translation units: accumulator layout, CTA-wide
// reduction, pixel iteration.
#pragma once
#include "lk_kernels.cuh"
#include "lk_warp.cuh"

namespace mtfb {

enum { CTRL_NEXT = 0, CTRL_BREAK = 1, CTRL_REJECT = 2 };

template<int S> struct AccLayout {
	static constexpr int NH = S*(S + 1) / 2;
	static constexpr int NA = 1 + S + NH;             // sum r^2 | J^T d | upper triangle of J^T J
	__host__ __device__ static constexpr int tri(int i, int j){ return i*S - i*(i - 1) / 2 + (j - i); }  // i <= j
};

// a CTA of one warp (T == 32: one warp tracks one patch) needs no block barrier
template<int T> __device__ __forceinline__ void cta_sync(){
	if(T == 32) __syncwarp(); else __syncthreads();
}

// CTA-wide sum of a per-thread accumulator vector; result in s_sum[0..CNT) after the call.
template<int CNT, int T> __device__ __forceinline__ void block_reduce(double (&acc)[CNT], double *s_part /* [T/32][CNT] */,
	double *s_sum){
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int idx[3];
	warp_reduce_scatter<CNT>(acc, lane, idx);
	if(T == 32){
		if(idx[0] >= 0) s_sum[idx[0]] = acc[0];
		if(CNT > 1 && idx[1] >= 0) s_sum[idx[1]] = acc[CNT > 1 ? 1 : 0];
		if(CNT > 64 && idx[2] >= 0) s_sum[idx[2]] = acc[CNT > 2 ? 2 : 0];
		__syncwarp();
		return;
	}
	if(idx[0] >= 0) s_part[warp*CNT + idx[0]] = acc[0];
	if(CNT > 1 && idx[1] >= 0) s_part[warp*CNT + idx[1]] = acc[CNT > 1 ? 1 : 0];
	if(CNT > 64 && idx[2] >= 0) s_part[warp*CNT + idx[2]] = acc[CNT > 2 ? 2 : 0];
	__syncthreads();
	for(int e = threadIdx.x; e < CNT; e += T){
		double s = s_part[e];
#pragma unroll
		for(int w = 1; w < T / 32; ++w) s += s_part[w*CNT + e];
		s_sum[e] = s;
	}
	__syncthreads();
}

// the same for fp32 per-thread sums: the butterfly runs in fp32 (half the shuffles, fp32 adds), the <= 3 entries a lane
// is left with are widened to fp64 for the sum across warps
template<int CNT, int T> __device__ __forceinline__ void block_reduce_f32(float (&acc)[CNT], double *s_part /* [T/32][CNT] */,
	double *s_sum){
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int idx[3];
	warp_reduce_scatter<CNT>(acc, lane, idx);
	double *dst = (T == 32) ? s_sum : s_part + warp*CNT;
	if(idx[0] >= 0) dst[idx[0]] = (double)acc[0];
	if(CNT > 1 && idx[1] >= 0) dst[idx[1]] = (double)acc[CNT > 1 ? 1 : 0];
	if(CNT > 64 && idx[2] >= 0) dst[idx[2]] = (double)acc[CNT > 2 ? 2 : 0];
	if(T == 32){ __syncwarp(); return; }
	__syncthreads();
	for(int e = threadIdx.x; e < CNT; e += T){
		double s = s_part[e];
#pragma unroll
		for(int w = 1; w < T / 32; ++w) s += s_part[w*CNT + e];
		s_sum[e] = s;
	}
	__syncthreads();
}

// resident CTAs per SM requested from the compiler: OCC 0 / 1 / 2 = about 8 / 12 / 16 warps per SM
// (<= 255 / 168 / 128 registers per thread)
__host__ __device__ constexpr int min_blocks(int T, int OCC){
	return (OCC == 0 ? 8 : OCC == 1 ? 12 : 16) * 32 / T > 0 ? (OCC == 0 ? 8 : OCC == 1 ? 12 : 16) * 32 / T : 1;
}

// CTA-wide sums of K scalars, result in every thread (fixed order: deterministic)
template<int K, int T> __device__ __forceinline__ void block_allreduce(double (&v)[K], double *s_tmp /* [T/32][K] */){
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for(int k = 0; k < K; ++k){
#pragma unroll
		for(int off = 16; off >= 1; off >>= 1) v[k] += __shfl_xor_sync(FULL_MASK, v[k], off);
	}
	if(T == 32) return;
	__syncthreads();                        // s_tmp may still be read from a previous call
	if(lane == 0){
#pragma unroll
		for(int k = 0; k < K; ++k) s_tmp[warp*K + k] = v[k];
	}
	__syncthreads();
#pragma unroll
	for(int k = 0; k < K; ++k){
		double s = s_tmp[k];
#pragma unroll
		for(int w = 1; w < T / 32; ++w) s += s_tmp[w*K + k];
		v[k] = s;
	}
}

// ---- TMA / mbarrier plumbing (PTX; SASS: UBLKCP for the 1-D bulk copy, UTMALDG for the tensor copy)
__device__ __forceinline__ unsigned smem_u32(const void *p){ return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count){
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes){
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned phase){
	unsigned done;
	do{
		asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
			: "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
	} while(!done);
}
// 1-D bulk async copy global -> shared
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar){
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// am.updatePixGrad + ssm.cmpt*PixJacobian in their two flavours (NT/FCLK.cc:222-236, NT/ESM.cc:387-404, NT/ICLK.cc:205-219):
//   chained:      gradient of the image at the warped point (updatePixGrad(pts)), then cmptWarpedPixJacobian
//   not chained:  gradient of the warped image (updateGradPts + getWarpedImgGrad), then cmptInitPixJacobian
// fills smp.val / gx / gy (raw pixel units times pix_mult; the caller applies pix_add to val) and the Jacobian row
template<int SSM, bool UNIT_MULT, class MW> __device__ __forceinline__ void pixel_value_and_gradient(const DevBatch &b, const MW &W,
	const PixGeom &g, Sample &smp){
	if(b.chained){
		smp = sample_fast<UNIT_MULT>(b.img, g.wx, g.wy, b.grad_eps, b.pix_mult);
		if(smp.lit) sample_literal(b.img, g.wx, g.wy, b.grad_eps, b.grad_mult, smp);
	} else{
		smp.val = sample_pixel(b.img, g.wx, g.wy); smp.lit = 0;
		warped_image_gradient<SSM>(b.img, W, g, b.grad_eps, b.grad_mult, smp.gx, smp.gy);
	}
}
template<int SSM, class MW> __device__ __forceinline__ void pixel_jacobian_row(const DevBatch &b, const MW &W, const double *abcd,
	const PixGeom &g, double gx, double gy, double *J){
	if(b.chained) warped_pix_jacobian<SSM>(W, abcd, g, gx, gy, J);
	else init_pix_jacobian<SSM>(g.ix, g.iy, gx, gy, J);
}

struct PixIter {
	int pix, row, col, dcol, drow, resx;
	__device__ __forceinline__ PixIter(){}
	__device__ __forceinline__ PixIter(int tid, int step, int _resx) : pix(tid), row(tid / _resx), col(tid % _resx),
		dcol(step % _resx), drow(step / _resx), resx(_resx){}
	__device__ __forceinline__ void next(int step){
		pix += step; col += dcol; row += drow;
		if(col >= resx){ col -= resx; ++row; }
	}
};


} // namespace mtfb
