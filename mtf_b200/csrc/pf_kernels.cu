// pf_kernels.cu -- particle evaluation for the particle filter (sm_100a).
//
// Replaces the body of the particle loop of nt::PF::update (SM/src/NT/PF.cc:303-320; templated twin
// SM/src/PF.cc:198-289):
//     ssm->setState(state)            ProjectiveBase.cc:41-49 / Affine.cc:109-115: warp from state, re-warp N points
//     am->updatePixVals(ssm->getPts())   ImageBase.cc:268-290: N bilinear samples
//     am->updateSimilarity(false)     SSDBase.cc:75-96: f = -|| It - I0 ||^2 / 2
//     am->getLikelihood()             SSD.h:41-43: exp(-alpha sqrt(-f / N))
// Random perturbation, resampling and the mean state stay on the host: this kernel is the
// warp-sample-similarity part.  One CTA serves one object (= one initialised template) and a slice of its
// particles; the object's template and its DLT-warped grid (the iteration-invariant half of the geometry) are
// staged in shared memory once per CTA -- with bulk async copies for the template -- and every warp then
// evaluates whole particles: lanes stride over the pixels, one shuffle tree per particle.
#include "lk_common.cuh"

namespace mtfb {

// dynamic shared memory: I0[N] | gA[N] | gB[N] | (gC[N] for the homography)
//   homography: (gA, gB, gC) = init_pts_hm = dlt . (u, v, 1)   (Homography.cc:68 keeps the DLT's third row)
//   affine:     (gA, gB)     = init_pts    = dehomogenize(dlt . (u, v, 1))
// AM = AM_NCC: am->updateSimilarity(false) is NCC.cc:124-161 (f = <I0c, Itc> / (|Itc| c)) and the likelihood NCC.cc:50-53
// (exp(-alpha (1/f - 1)^2)); the template is staged centred (I0 - mean) and the three sums sum It, sum It^2,
// sum I0c It of one sweep give a = sum I0c It - mean_t sum I0c and b^2 = sum It^2 - N mean_t^2.
template<int SSM, int AM, int T>
__global__ void __launch_bounds__(T) pf_evaluate_kernel(DevBatch b, const double *__restrict__ states, int n_particles,
	double *__restrict__ likelihood, double *__restrict__ similarity, double alpha){
	constexpr int S = StateSize<SSM>::value;
	extern __shared__ __align__(16) double smem[];
	__shared__ __align__(8) unsigned long long s_bar;
	const int obj = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int N = b.N;
	double *s_I0 = smem, *s_gA = smem + N, *s_gB = smem + 2 * N, *s_gC = smem + 3 * N;
	// stage the template with one bulk copy: cp.async.bulk needs a 16-byte aligned source and a multiple of 16 bytes.  The
	// object's row starts at I0 + obj * N doubles, which is only 8-byte aligned when obj * N (or the offset of I0 inside the
	// per-patch allocation: odd P) is odd -- such rows, and the odd tail element of any row, are copied by the threads instead
	const double *src = b.I0 + (size_t)obj*N;
	const bool src_aligned = (reinterpret_cast<unsigned long long>(src) & 15ull) == 0;
	const unsigned bulk_bytes = src_aligned ? (unsigned)((N * 8) & ~15) : 0u;
	if(tid == 0) mbar_init(&s_bar, 1);
	__syncthreads();
	if(tid == 0 && bulk_bytes){
		mbar_expect_tx(&s_bar, bulk_bytes);
		bulk_copy_g2s(s_I0, src, bulk_bytes, &s_bar);
	}
	for(int i = (int)(bulk_bytes / 8) + tid; i < N; i += T) s_I0[i] = src[i];
	{
		Mat3 dlt;
#pragma unroll
		for(int i = 0; i < 9; ++i) dlt.m[i] = b.dlt[(size_t)obj * 9 + i];
		for(PixIter it(tid, T, b.resx); it.pix < N; it.next(T)){
			const double u = b.xv[it.col], v = b.yv[it.row];
			double hx = dlt.m[0] * u; hx = hx + dlt.m[1] * v; hx = hx + dlt.m[2] * 1.0;
			double hy = dlt.m[3] * u; hy = hy + dlt.m[4] * v; hy = hy + dlt.m[5] * 1.0;
			double hz = dlt.m[6] * u; hz = hz + dlt.m[7] * v; hz = hz + dlt.m[8] * 1.0;
			if(b.norm_init){ hx = u; hy = v; hz = 1.0; }              // hom_normalized_init: init_pts_hm = (u, v, 1)
			if(SSM == SSM_HOM){ s_gA[it.pix] = hx; s_gB[it.pix] = hy; s_gC[it.pix] = hz; }
			else{ const double r = ieee_rcp(hz); s_gA[it.pix] = div_by(hx, hz, r); s_gB[it.pix] = div_by(hy, hz, r); }
		}
	}
	if(bulk_bytes) mbar_wait(&s_bar, 0);
	__syncthreads();
	__shared__ double s_i0c_part[T / 32];
	double sum_i0c = 0, ncc_c = 1;
	if(AM == AM_NCC){
		const double I0_mean = b.am_scal[(size_t)obj * 8];
		ncc_c = b.am_scal[(size_t)obj * 8 + 1];
		double part = 0;
		for(int i = tid; i < N; i += T){ const double v = s_I0[i] - I0_mean; s_I0[i] = v; part += v; }     // I0_cntr (NCC.cc:63-64)
#pragma unroll
		for(int off = 16; off >= 1; off >>= 1) part += __shfl_xor_sync(FULL_MASK, part, off);
		if(lane == 0) s_i0c_part[warp] = part;
		__syncthreads();
		for(int w = 0; w < T / 32; ++w) sum_i0c += s_i0c_part[w];
	}
	const int warps = T / 32;
	for(int pi = blockIdx.x*warps + warp; pi < n_particles; pi += gridDim.x*warps){
		const double *st = states + ((size_t)obj*n_particles + pi)*S;
		double sv[S];
#pragma unroll
		for(int s = 0; s < S; ++s) sv[s] = st[s];
		const Mat3 W = warp_from_state<SSM>(sv);
		double acc = 0, acc_t = 0, acc_tt = 0;
		for(int pix = lane; pix < N; pix += 32){
			double wx, wy;
			if(SSM == SSM_HOM){
				const double hx = s_gA[pix], hy = s_gB[pix], hz = s_gC[pix];
				double cx = W.m[0] * hx; cx = cx + W.m[1] * hy; cx = cx + W.m[2] * hz;
				double cy = W.m[3] * hx; cy = cy + W.m[4] * hy; cy = cy + W.m[5] * hz;
				double cz = W.m[6] * hx; cz = cz + W.m[7] * hy; cz = cz + W.m[8] * hz;
				const double r = ieee_rcp(cz);
				wx = div_by(cx, cz, r); wy = div_by(cy, cz, r);
			} else{
				const double ix = s_gA[pix], iy = s_gB[pix];
				wx = W.m[0] * ix; wx = wx + W.m[1] * iy; wx = wx + W.m[2] * 1.0;
				wy = W.m[3] * ix; wy = wy + W.m[4] * iy; wy = wy + W.m[5] * 1.0;
			}
			const double It = sample_pixel(b.img, wx, wy);
			if(AM == AM_NCC){
				acc = fma(s_I0[pix], It, acc); acc_t += It; acc_tt = fma(It, It, acc_tt);
			} else{
				const double d = It - s_I0[pix];
				acc = fma(d, d, acc);
			}
		}
#pragma unroll
		for(int off = 16; off >= 1; off >>= 1){
			acc += __shfl_xor_sync(FULL_MASK, acc, off);
			if(AM == AM_NCC){ acc_t += __shfl_xor_sync(FULL_MASK, acc_t, off); acc_tt += __shfl_xor_sync(FULL_MASK, acc_tt, off); }
		}
		if(lane == 0){
			double f, lik;
			if(AM == AM_NCC){
				const double mean_t = acc_t / N;
				const double a = acc - mean_t*sum_i0c, bb = acc_tt - N*mean_t*mean_t;
				// a particle that samples a constant (e.g. thrown out of the image: 128 everywhere) has Itc = 0: the reference's
				// 0 / 0.  The one-sweep variance only knows that to rounding, so "constant" is bb <= 1e-12 sum It^2
				f = (bb > 1e-12*acc_tt) ? a / (sqrt(bb) * ncc_c) : nan("");
				const double d = (1.0 / f) - 1;
				lik = exp(-alpha * d*d);
			} else{
				f = -acc / 2;
				lik = exp(-alpha * sqrt(-f / double(N)));
			}
			if(similarity) similarity[(size_t)obj*n_particles + pi] = f;
			if(likelihood) likelihood[(size_t)obj*n_particles + pi] = lik;
		}
	}
}

cudaError_t launch_pf_evaluate(int am, int ssm, const DevBatch &b, const double *d_states, int n_particles,
	double *d_likelihood, double *d_similarity, double alpha, cudaStream_t st){
	if(am != AM_SSD && am != AM_NCC) return cudaErrorNotSupported;
	constexpr int T = 256;
	const size_t smem = (size_t)b.N * 8 * (ssm == SSM_HOM ? 4 : 3);
	if(smem > 220 * 1024) return cudaErrorInvalidValue;
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const int ctas_per_sm = (int)((220 * 1024) / (smem + 1024)) < 2 ? 1 : 2;
	// enough CTAs per object to fill the machine, never more than one particle per warp
	int per_obj = (sms*ctas_per_sm + b.P - 1) / b.P;
	const int max_useful = (n_particles + T / 32 - 1) / (T / 32);
	if(per_obj > max_useful) per_obj = max_useful;
	if(per_obj < 1) per_obj = 1;
	dim3 grid(per_obj, b.P);
	cudaError_t e = cudaSuccess;
#define MTFB_PF_LAUNCH(SSM_, AM_) do{ \
		e = cudaFuncSetAttribute(pf_evaluate_kernel<SSM_, AM_, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
		if(e != cudaSuccess) return e; \
		pf_evaluate_kernel<SSM_, AM_, T><<<grid, T, smem, st>>>(b, d_states, n_particles, d_likelihood, d_similarity, alpha); }while(0)
	if(ssm == SSM_HOM){ if(am == AM_NCC) MTFB_PF_LAUNCH(SSM_HOM, AM_NCC); else MTFB_PF_LAUNCH(SSM_HOM, AM_SSD); }
	else{ if(am == AM_NCC) MTFB_PF_LAUNCH(SSM_AFF, AM_NCC); else MTFB_PF_LAUNCH(SSM_AFF, AM_SSD); }
#undef MTFB_PF_LAUNCH
	return cudaGetLastError();
}

} // namespace mtfb
