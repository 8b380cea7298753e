// Microbenchmark (B200): issue rates that bound the fp32 LK kernel (lk_ssd_f32.cu): FFMA with three register operands,
// the packed fma.rn.f32x2 (SASS FFMA2), the conversion / rounding ops of the sampling front end (FRND, F2I), MUFU.RCP,
// and an FFMA + integer mix.  Build: nvcc -arch=sm_100a -O3 fp32_pipes.cu -o fp32_pipes.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2 &d, float2 a, float2 b){
	unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d), aa = *reinterpret_cast<unsigned long long*>(&a),
		bb = *reinterpret_cast<unsigned long long*>(&b);
	asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
	d = *reinterpret_cast<float2*>(&dd);
}

template<int MODE> __global__ void k(float *out, int iters, float x){
	float acc[32];
	for(int i = 0; i < 32; ++i) acc[i] = threadIdx.x*1e-3f + i;
	float a = x + threadIdx.x*1e-6f, b = x - threadIdx.x*1e-6f;
	int n = threadIdx.x;
	for(int it = 0; it < iters; ++it){
		if(MODE == 0){                        // 32 independent FFMA (3 distinct register operands)
#pragma unroll
			for(int i = 0; i < 32; ++i) acc[i] = fmaf(acc[(i + 1) & 31] * 0 + a, b, acc[i]);
		}
		if(MODE == 1){                        // 16 independent FFMA2 = 32 fused multiply-adds
#pragma unroll
			for(int i = 0; i < 16; ++i){
				float2 d = make_float2(acc[2 * i], acc[2 * i + 1]);
				ffma2(d, make_float2(a, b), make_float2(b, a));
				acc[2 * i] = d.x; acc[2 * i + 1] = d.y;
			}
		}
		if(MODE == 2){                        // FRND.FLOOR x16
#pragma unroll
			for(int i = 0; i < 16; ++i) acc[i] = floorf(acc[i] * 1.0001f) + a;
		}
		if(MODE == 3){                        // F2I + I2F x16
#pragma unroll
			for(int i = 0; i < 16; ++i) acc[i] = (float)(__float2int_rd(acc[i]) + n);
		}
		if(MODE == 4){                        // MUFU.RCP x16
#pragma unroll
			for(int i = 0; i < 16; ++i){ float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(acc[i])); acc[i] = r + a; }
		}
		if(MODE == 5){                        // 24 FFMA + 8 IMAD interleaved
#pragma unroll
			for(int i = 0; i < 24; ++i) acc[i] = fmaf(a, b, acc[i]);
#pragma unroll
			for(int i = 0; i < 8; ++i) n = n * 3 + it;
		}
	}
	float s = n;
	for(int i = 0; i < 32; ++i) s += acc[i];
	out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}

template<int MODE> void run(const char *name, int blocks, int threads, double ops_per_iter_per_warp){
	float *out; cudaMalloc(&out, sizeof(float)*blocks*threads);
	int iters = 8192;
	k<MODE><<<blocks, threads>>>(out, 16, 1.0000001f);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	k<MODE><<<blocks, threads>>>(out, iters, 1.0000001f);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	double warps = (double)blocks*threads / 32;
	double warp_inst = warps*iters*ops_per_iter_per_warp;
	printf("%-34s blocks %4d x %4d thr: %8.3f ms  %6.3f counted-inst/clk/SM @1.965GHz\n", name, blocks, threads, ms,
		warp_inst / (ms*1e-3) / 148 / 1.965e9);
	cudaFree(out);
}

int main(){
	for(int thr : {256, 512, 1024}){
		run<0>("FFMA x32 indep (counts FFMA)", 148 * 2, thr, 32);
		run<1>("FFMA2 x16 indep (counts FFMA2)", 148 * 2, thr, 16);
		run<2>("FRND.FLOOR x16 (+FMUL,FADD)", 148 * 2, thr, 16);
		run<3>("F2I + I2F x16 (counts pairs)", 148 * 2, thr, 16);
		run<4>("MUFU.RCP x16 (+FADD)", 148 * 2, thr, 16);
		run<5>("FFMA x24 + IMAD x8 (counts 32)", 148 * 2, thr, 32);
	}
	return 0;
}
