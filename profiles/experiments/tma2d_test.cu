// standalone check of the 2-D TMA window copy used by ssd_update_kernel
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void *p){ return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, int x0, int y0, float *out){
	__shared__ __align__(128) float tile[64 * 64];
	__shared__ __align__(8) unsigned long long bar;
	if(threadIdx.x == 0){
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if(threadIdx.x == 0){
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(64 * 64 * 4) : "memory");
		asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
			:: "r"(smem_u32(tile)), "l"(&tmap), "r"(x0), "r"(y0), "r"(smem_u32(&bar)) : "memory");
	}
	unsigned done;
	do{ asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory"); } while(!done);
	for(int i = threadIdx.x; i < 64 * 64; i += blockDim.x) out[i] = tile[i];
}
int main(){
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
		CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	printf("encode: %d\n", (int)r);
	for(int t : {32, 128}){
		k<<<1, t>>>(tm, 37, -5, o);
		e = cudaDeviceSynchronize();
		std::vector<float> res(64 * 64);
		cudaMemcpy(res.data(), o, res.size() * 4, cudaMemcpyDeviceToHost);
		int bad = 0;
		for(int y = 0; y < 64; ++y) for(int x = 0; x < 64; ++x){
			int gy = y - 5, gx = x + 37; float exp = (gy < 0 || gy >= h || gx >= w) ? 0.f : img[gy*pitch + gx];
			if(res[y * 64 + x] != exp) ++bad;
		}
		printf("threads %d: sync %s, mismatches %d\n", t, cudaGetErrorString(e), bad);
	}
	return 0;
}
