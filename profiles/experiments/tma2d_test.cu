// standalone check of the 2-D TMA window copy used by ssd_update_kernel
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void *p){ return (unsigned)__cvta_generic_to_shared(p); }
// one lane of a CONVERGED warp issues the copy: UTMALDG is a uniform-datapath instruction; under `if(threadIdx.x == 0)` (a
// divergent branch) ptxas emitted it unpredicated and every lane faulted with "illegal instruction" (round-1 finding,
// profiles/r02_tma2d_run.log keeps that run); elect.sync inside a warp-uniform branch is the form that works
__device__ __forceinline__ bool elect_one(){
	unsigned pred;
	asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
	return pred != 0;
}
__global__ void k(const __grid_constant__ CUtensorMap tmap, int x0, int y0, float *out){
	__shared__ __align__(128) float tile[64 * 64];
	__shared__ __align__(8) unsigned long long bar;
	if(threadIdx.x == 0){
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if(threadIdx.x < 32 && elect_one()){
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(64 * 64 * 4) : "memory");
		asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
			:: "r"(smem_u32(tile)), "l"(&tmap), "r"(x0), "r"(y0), "r"(smem_u32(&bar)) : "memory");
	}
	unsigned done;
	do{ asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory"); } while(!done);
	for(int i = threadIdx.x; i < 64 * 64; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char **argv){
	const int variant = argc > 1 ? atoi(argv[1]) : 0;   // 0: as the kernel would use it; 1: no L2 promotion; 2: 128B swizzle; 3: in-bounds coordinates
	const int h = 384, w = 384, pitch = 384;
	std::vector<float> img(h*pitch);
	for(int i = 0; i < h*pitch; ++i) img[i] = (float)(i % 1000);
	float *d; cudaMalloc(&d, img.size() * 4); cudaMemcpy(d, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
	float *o; cudaMalloc(&o, 64 * 64 * 4);
	typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
		const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	void *ptr = nullptr; cudaDriverEntryPointQueryResult q;
	cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
	printf("entry point: %d %d %p\n", (int)e, (int)q, ptr);
	alignas(64) CUtensorMap tm;
	const cuuint64_t dims[2] = { (cuuint64_t)w, (cuuint64_t)h }; const cuuint64_t strides[1] = { (cuuint64_t)pitch * 4 };
	const cuuint32_t box[2] = { 64, 64 }; const cuuint32_t estr[2] = { 1, 1 };
	CUresult r = ((encode_fn)ptr)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
		CU_TENSOR_MAP_SWIZZLE_NONE, variant == 1 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
		CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	printf("variant %d encode: %d\n", variant, (int)r);
	for(int t : {32, 128}){
		k<<<1, t>>>(tm, variant == 3 ? 32 : 37, variant == 3 ? 8 : -5, o);
		e = cudaDeviceSynchronize();
		std::vector<float> res(64 * 64);
		cudaMemcpy(res.data(), o, res.size() * 4, cudaMemcpyDeviceToHost);
		int bad = 0;
		for(int y = 0; y < 64; ++y) for(int x = 0; x < 64; ++x){
			int gy = y + (variant == 3 ? 8 : -5), gx = x + (variant == 3 ? 32 : 37); float exp = (gy < 0 || gy >= h || gx >= w) ? 0.f : img[gy*pitch + gx];
			if(res[y * 64 + x] != exp) ++bad;
		}
		printf("threads %d: sync %s, mismatches %d\n", t, cudaGetErrorString(e), bad);
	}
	return 0;
}
