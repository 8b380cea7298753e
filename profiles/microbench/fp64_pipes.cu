// Microbenchmark (B200): throughput of DFMA, of mma.sync.m8n8k4.f64 (DMMA), and of both interleaved,
// to decide whether the J^T J accumulation belongs on the fp64 tensor path.  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b){
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template<int MODE> __global__ void k(double *out, int iters, double x){
	double acc[16]; double c0[8], c1[8];
	for(int i = 0; i < 16; ++i) acc[i] = threadIdx.x*1e-3 + i;
	for(int i = 0; i < 8; ++i){ c0[i] = i; c1[i] = -i; }
	double a = x + threadIdx.x, b = x - threadIdx.x;
	for(int it = 0; it < iters; ++it){
		if(MODE == 0 || MODE == 2){
#pragma unroll
			for(int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
		}
		if(MODE == 1 || MODE == 2){
#pragma unroll
			for(int i = 0; i < 8; ++i) dmma(c0[i], c1[i], a, b);
		}
		if(MODE == 3){       // dependent DFMA chain: latency
#pragma unroll
			for(int i = 0; i < 16; ++i) acc[0] = fma(acc[0], a, b);
		}
		if(MODE == 4){       // IEEE reciprocal chain
#pragma unroll
			for(int i = 0; i < 16; ++i) acc[0] = __drcp_rn(acc[0]) + a;
		}
		if(MODE == 5){       // IEEE division, independent
#pragma unroll
			for(int i = 0; i < 16; ++i) acc[i] = a / acc[i];
		}
	}
	double s = 0;
	for(int i = 0; i < 16; ++i) s += acc[i];
	for(int i = 0; i < 8; ++i) s += c0[i] + c1[i];
	out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}

template<int MODE> void run(const char *name, int blocks, int threads, double ops_per_iter_per_warp){
	double *out; cudaMalloc(&out, sizeof(double)*blocks*threads);
	int iters = 4096;
	k<MODE><<<blocks, threads>>>(out, 16, 1.0000001);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	k<MODE><<<blocks, threads>>>(out, iters, 1.0000001);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	double warps = (double)blocks*threads / 32;
	double warp_inst = warps*iters*ops_per_iter_per_warp;
	printf("%-28s blocks %4d x %4d thr: %8.3f ms  %8.2f G warp-inst/s  = %6.3f warp-inst/clk/SM @1.965GHz\n", name, blocks, threads, ms,
		warp_inst / ms*1e-6, warp_inst / (ms*1e-3) / 148 / 1.965e9);
	cudaFree(out);
}

int main(){
	for(int thr : {128, 256, 512, 1024}){
		run<0>("DFMA x16 indep", 148 * 2, thr, 16);
		run<1>("DMMA m8n8k4 x8 indep", 148 * 2, thr, 8);
		run<2>("DFMA x16 + DMMA x8", 148 * 2, thr, 24);
	}
	run<3>("DFMA dependent chain", 148, 32, 16);
	run<4>("drcp_rn dependent chain", 148, 32, 16);
	run<5>("ddiv x16 indep", 148 * 2, 512, 16);
	return 0;
}
