// pf_tracker.cu -- the particle filter search method around the particle evaluation kernels (pf_kernels.cu, lk_ssd_f32.cu,
// lk_mi.cu): everything nt::PF::update does per frame besides evaluating a particle (SM/src/NT/PF.cc:207-446), for P
// objects x n_particles particles at once, on the device:
//
//   pf_perturb_kernel    the dynamic model: ssm->{additive,compositional}{RandomWalk,AutoRegression1} (ProjectiveBase.cc:255-299,
//                        Homography.cc:917-942) on a perturbation from ssm->generatePerturbation (ProjectiveBase.cc:301-306 one
//                        normal deviate per state entry; Homography.cc:899-915 corner based: a common translation and four
//                        corner offsets pushed through the 4-point homography, in closed form, one thread per particle)
//   (evaluation)         ssm->setState -> am->updatePixVals -> am->updateSimilarity(false) -> likelihood (NT/PF.cc:303-320)
//   pf_weights_kernel    measurement likelihood from the similarity (NT/PF.cc:329-339), particle_cum_wts in the reference's
//                        sequential summation order (bit-identical prefix sums), max_wt_id, the adaptive-resampling test
//                        (NT/PF.cc:389-397), normalisation of the cumulative weights (NT/PF.cc:457)
//   pf_resample_kernel   binary / linear multinomial resampling (NT/PF.cc:448-540: both pick the first particle whose normalised
//                        cumulative weight is >= the uniform deviate), into the other particle set
//   pf_commit_kernel     copy of the resampled set back for the objects that resampled (the reference flips curr_set_id, per
//                        tracker instance; here every object of the batch keeps its particles in the same array)
//   pf_mean_kernel       mean_type None / SSM / Corners (NT/PF.cc:413-427, ProjectiveBase.cc:308-314, NT/PF.cc:587-594), setState /
//                        setCorners, the corner-change test, reset_to_mean (NT/PF.cc:436-438)
//
// Random numbers: the reference seeds Boost generators from random_device (NT/PF.cc:100-110, ProjectiveBase.cc:192-197), so
// its trajectories are not reproducible by design.  Here every deviate is a pure function of (seed, frame, iteration, object,
// particle, entry) through Philox4x32-10 and a Box-Muller transform -- or comes from arrays the host supplies
// (mtfb_pf_set_random_stream), which is how the parity tests drive the oracle's PF loop and this one with one stream.
#include "lk_common.cuh"
#include "lk_solve.cuh"
#include "pf_tracker.cuh"

namespace mtfb {

namespace {

// ---- Philox4x32-10 (Salmon et al., SC'11): counter-based, so no generator state is carried between launches
__device__ __forceinline__ void philox_round(unsigned (&c)[4], unsigned k0, unsigned k1){
	const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
	const unsigned hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
	const unsigned hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
	const unsigned n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
	c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32(unsigned (&c)[4], unsigned k0, unsigned k1){
#pragma unroll
	for(int r = 0; r < 10; ++r){
		philox_round(c, k0, k1);
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
}
// two uniform deviates in (0, 1] (53 random bits each would need 8 words; 32 + 20 bits are plenty for a sampler)
__device__ __forceinline__ void uniform2(unsigned long long seed, unsigned a, unsigned b, unsigned c, unsigned d, double &u0, double &u1){
	unsigned ctr[4] = { a, b, c, d };
	philox4x32(ctr, (unsigned)seed, (unsigned)(seed >> 32));
	const double s = 1.0 / 4503599627370496.0;                       // 2^-52
	u0 = ((double)(((unsigned long long)ctr[0] << 20) | (ctr[1] >> 12)) + 1.0) * s;
	u1 = ((double)(((unsigned long long)ctr[2] << 20) | (ctr[3] >> 12)) + 1.0) * s;
}
// two standard normal deviates (Box-Muller)
__device__ __forceinline__ void normal2(unsigned long long seed, unsigned a, unsigned b, unsigned c, unsigned d, double &n0, double &n1){
	double u0, u1;
	uniform2(seed, a, b, c, d, u0, u1);
	const double r = sqrt(-2.0 * log(u0));
	double sn, cs;
	sincospi(2.0 * u1, &sn, &cs);
	n0 = r * cs; n1 = r * sn;
}

// streams of the counter's last word
enum { STREAM_NORMALS = 0, STREAM_UNIFORMS = 0x40000000 };

// perturbed_warp of the four dynamic models; writes the new state and auto-regression term of one particle
template<int SSM> __device__ __forceinline__ void dynamic_model(const PFDev &pf, const double *z, double *st, double *ar){
	constexpr int S = StateSize<SSM>::value;
	double base[S], base_ar[S];
#pragma unroll
	for(int s = 0; s < S; ++s){ base[s] = st[s]; base_ar[s] = ar[s]; }
	if(pf.update_type == MTFB_PF_UPDATE_ADDITIVE){
		if(pf.dynamic_model == MTFB_PF_RANDOM_WALK){
#pragma unroll
			for(int s = 0; s < S; ++s) st[s] = base[s] + z[s];                                    // ProjectiveBase.cc:255-259
		} else{
#pragma unroll
			for(int s = 0; s < S; ++s){
				const double ns = base[s] + base_ar[s] + z[s];                                    // ProjectiveBase.cc:270-275
				st[s] = ns; ar[s] = pf.ar_coeff*(ns - base[s]);
			}
		}
		return;
	}
	const Mat3 Wb = warp_from_state<SSM>(base), Wz = warp_from_state<SSM>(z);
	if(pf.dynamic_model == MTFB_PF_RANDOM_WALK){
		Mat3 W = mat3_mul(Wb, Wz);                                                                // Homography.cc:917-926
		if(SSM == SSM_HOM){
			const double d = W.m[8];
#pragma unroll
			for(int i = 0; i < 9; ++i) W.m[i] = W.m[i] / d;
		}
		state_from_warp<SSM>(st, W);
		return;
	}
	const Mat3 War = warp_from_state<SSM>(base_ar);
	Mat3 W = mat3_mul(mat3_mul(Wb, War), Wz);                                                     // Homography.cc:928-942
	if(SSM == SSM_HOM){
		const double d = W.m[8];
#pragma unroll
		for(int i = 0; i < 9; ++i) W.m[i] = W.m[i] / d;
	}
	Mat3 A = mat3_mul(mat3_inverse(Wb), W);
	if(SSM == SSM_HOM){
		const double d = A.m[8];
#pragma unroll
		for(int i = 0; i < 9; ++i) A.m[i] = A.m[i] / d;
	}
	state_from_warp<SSM>(st, W);
	double na[S];
	state_from_warp<SSM>(na, A);
#pragma unroll
	for(int s = 0; s < S; ++s) ar[s] = na[s] * pf.ar_coeff;
}

// one thread per particle: generatePerturbation of ProjectiveBase (one deviate per state entry), then the dynamic model
template<int SSM> __global__ void __launch_bounds__(128) pf_perturb_kernel(PFDev pf, DevBatch b, unsigned iter_tag){
	constexpr int S = StateSize<SSM>::value;
	const int obj = blockIdx.y, i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= pf.n_particles || pf.done[obj]) return;
	const size_t pi = (size_t)obj*pf.n_particles + i;
	double z[S];
	if(pf.normals_in){
#pragma unroll
		for(int s = 0; s < S; ++s) z[s] = pf.normals_in[pi*pf.n_normals + s];
	} else{
#pragma unroll
		for(int s = 0; s < S; s += 2){
			double n0, n1;
			normal2(pf.seed, (unsigned)i, (unsigned)(obj + pf.object_offset), iter_tag, STREAM_NORMALS + s / 2, n0, n1);
			z[s] = n0; if(s + 1 < S) z[s + 1] = n1;
		}
	}
	if(pf.normals_out){
#pragma unroll
		for(int s = 0; s < S; ++s) pf.normals_out[pi*pf.n_normals + s] = z[s];
	}
#pragma unroll
	for(int s = 0; s < S; ++s) z[s] = pf.mean[s] + pf.sigma[s] * z[s];            // normal_distribution(mean, sigma)
	dynamic_model<SSM>(pf, z, pf.states + pi*S, pf.ar + pi*S);
}

// The homography that maps four points onto four points, in closed form (both quadruples in general position).  The reference
// takes it from the null vector of the 8 x 9 DLT matrix (utils::computeHomographyDLT, warpUtils.cc:171-223, JacobiSVD); four
// correspondences determine the homography exactly, so the two agree up to rounding.  Coordinates are centred and scaled with
// the IN points' frame first (x' = s (x - cx)), which keeps the 3 x 3 determinants below well conditioned for pixel coordinates;
// measured against the oracle's DLT on the perturbed corners of the tests: < 1e-12 relative on every entry (tests/test_pf.py).
//   M(p) = [l1 p1 | l2 p2 | l3 p3],  (l1, l2, l3) = adj([p1 p2 p3]) p4   maps (e1, e2, e3, e1 + e2 + e3) to p1..p4 (up to scale)
//   H' = M(out) adj(M(in)),   H = S^-1 H' S,   normalised by its last entry
__device__ __forceinline__ void basis_to_points(const double *x, const double *y, double *M){
	// adj([p1 p2 p3]) p4 with p_i = (x_i, y_i, 1)
	const double c00 = y[1] - y[2], c01 = x[2] - x[1], c02 = x[1] * y[2] - x[2] * y[1];
	const double c10 = y[2] - y[0], c11 = x[0] - x[2], c12 = x[2] * y[0] - x[0] * y[2];
	const double c20 = y[0] - y[1], c21 = x[1] - x[0], c22 = x[0] * y[1] - x[1] * y[0];
	const double l0 = c00 * x[3] + c01 * y[3] + c02, l1 = c10 * x[3] + c11 * y[3] + c12, l2 = c20 * x[3] + c21 * y[3] + c22;
	M[0] = l0 * x[0]; M[1] = l1 * x[1]; M[2] = l2 * x[2];
	M[3] = l0 * y[0]; M[4] = l1 * y[1]; M[5] = l2 * y[2];
	M[6] = l0; M[7] = l1; M[8] = l2;
}
__device__ __forceinline__ Mat3 homography_4pt(const double *in_c, const double *out_c){
	// the in points' frame
	const double cx = 0.25 * (in_c[0] + in_c[1] + in_c[2] + in_c[3]), cy = 0.25 * (in_c[4] + in_c[5] + in_c[6] + in_c[7]);
	double ext = 0;
#pragma unroll
	for(int i = 0; i < 4; ++i){ ext = fmax(ext, fabs(in_c[i] - cx)); ext = fmax(ext, fabs(in_c[4 + i] - cy)); }
	const double s = ext > 0 ? 1.0 / ext : 1.0;
	double xi[4], yi[4], xo[4], yo[4];
#pragma unroll
	for(int i = 0; i < 4; ++i){
		xi[i] = s * (in_c[i] - cx); yi[i] = s * (in_c[4 + i] - cy);
		xo[i] = s * (out_c[i] - cx); yo[i] = s * (out_c[4 + i] - cy);
	}
	double Mi[9], Mo[9];
	basis_to_points(xi, yi, Mi);
	basis_to_points(xo, yo, Mo);
	Mat3 A, B, adj;
#pragma unroll
	for(int i = 0; i < 9; ++i){ A.m[i] = Mi[i]; B.m[i] = Mo[i]; }
	// adjugate of Mi (transposed cofactors)
	adj.m[0] = A.m[4] * A.m[8] - A.m[5] * A.m[7]; adj.m[1] = A.m[2] * A.m[7] - A.m[1] * A.m[8]; adj.m[2] = A.m[1] * A.m[5] - A.m[2] * A.m[4];
	adj.m[3] = A.m[5] * A.m[6] - A.m[3] * A.m[8]; adj.m[4] = A.m[0] * A.m[8] - A.m[2] * A.m[6]; adj.m[5] = A.m[2] * A.m[3] - A.m[0] * A.m[5];
	adj.m[6] = A.m[3] * A.m[7] - A.m[4] * A.m[6]; adj.m[7] = A.m[1] * A.m[6] - A.m[0] * A.m[7]; adj.m[8] = A.m[0] * A.m[4] - A.m[1] * A.m[3];
	const Mat3 Hn = mat3_mul(B, adj);
	// H = S^-1 Hn S with S = [s 0 -s cx; 0 s -s cy; 0 0 1], S^-1 = [1/s 0 cx; 0 1/s cy; 0 0 1]
	Mat3 T;                                                               // Hn S
#pragma unroll
	for(int r = 0; r < 3; ++r){
		T.m[3 * r] = Hn.m[3 * r] * s; T.m[3 * r + 1] = Hn.m[3 * r + 1] * s;
		T.m[3 * r + 2] = Hn.m[3 * r + 2] - s * (Hn.m[3 * r] * cx + Hn.m[3 * r + 1] * cy);
	}
	Mat3 H;
	const double rs = ext > 0 ? ext : 1.0;
#pragma unroll
	for(int c = 0; c < 3; ++c){
		H.m[c] = rs * T.m[c] + cx * T.m[6 + c];
		H.m[3 + c] = rs * T.m[3 + c] + cy * T.m[6 + c];
		H.m[6 + c] = T.m[6 + c];
	}
	const double d = H.m[8];
#pragma unroll
	for(int k = 0; k < 9; ++k) H.m[k] = H.m[k] / d;
	return H;
}

// Homography with corner_based_sampling (Homography.cc:899-909), one thread per particle: rand_dist[0] gives the common
// translation (two deviates), rand_dist[1] the eight corner offsets; the perturbation is the state of the homography from the
// initial corners to the disturbed ones (estimateWarpFromCorners, Homography.cc:877-883).
__global__ void __launch_bounds__(128) pf_perturb_corner_kernel(PFDev pf, DevBatch b, unsigned iter_tag){
	constexpr int S = 8;
	const int obj = blockIdx.y, i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= pf.n_particles || pf.done[obj]) return;
	const size_t pi = (size_t)obj*pf.n_particles + i;
	double n[10];
	if(pf.normals_in){
#pragma unroll
		for(int s = 0; s < 10; ++s) n[s] = pf.normals_in[pi*pf.n_normals + s];
	} else{
#pragma unroll
		for(int s = 0; s < 10; s += 2) normal2(pf.seed, (unsigned)i, (unsigned)(obj + pf.object_offset), iter_tag, STREAM_NORMALS + s / 2, n[s], n[s + 1]);
	}
	if(pf.normals_out){
#pragma unroll
		for(int s = 0; s < 10; ++s) pf.normals_out[pi*pf.n_normals + s] = n[s];
	}
	const double tx = pf.mean[0] + pf.sigma[0] * n[0], ty = pf.mean[0] + pf.sigma[0] * n[1];
	double in_c[8], out_c[8];
#pragma unroll
	for(int c = 0; c < 4; ++c){
		// rand_d(0, c), rand_d(1, c) are drawn in this order (Homography.cc:903-906)
		const double dx = pf.mean[1] + pf.sigma[1] * n[2 + 2 * c], dy = pf.mean[1] + pf.sigma[1] * n[3 + 2 * c];
		in_c[c] = b.init_corners[(size_t)obj * 8 + c]; in_c[4 + c] = b.init_corners[(size_t)obj * 8 + 4 + c];
		out_c[c] = (in_c[c] + dx) + tx; out_c[4 + c] = (in_c[4 + c] + dy) + ty;
	}
	const Mat3 H = homography_4pt(in_c, out_c);
	double z[S];
	state_from_warp<SSM_HOM>(z, H);
	dynamic_model<SSM_HOM>(pf, z, pf.states + pi*S, pf.ar + pi*S);
}

// one CTA per object.  Weights from the evaluation kernel's outputs, their running sum in the reference's order, max_wt_id,
// the resampling decision, normalised cumulative weights.
__global__ void __launch_bounds__(256) pf_weights_kernel(PFDev pf){
	const int obj = blockIdx.x, tid = threadIdx.x, n = pf.n_particles;
	if(pf.done[obj]) return;
	extern __shared__ double s_w[];                                   // the n weights, then their running sums
	double *w = pf.weights + (size_t)obj*n, *cw = pf.cum_weights + (size_t)obj*n;
	const double *sim = pf.similarity + (size_t)obj*n;
	double *ws = s_w;
	for(int i = tid; i < n; i += blockDim.x){
		double v = w[i];                                              // the AM's likelihood (evaluation kernel)
		if(pf.likelihood_func != MTFB_PF_LIKELIHOOD_AM){
			const double m = pf.max_similarity - sim[i];                  // NT/PF.cc:326
			v = (pf.likelihood_func == MTFB_PF_LIKELIHOOD_GAUSSIAN) ? pf.measurement_factor * exp(-0.5 * m / pf.measurement_sigma)
				: 1.0 / (1.0 + m);
			w[i] = v;
		}
		s_w[i] = v;
	}
	__syncthreads();
	__shared__ int s_resample;
	__shared__ double s_total;
	if(tid == 0){
		// particle_cum_wts[i] = particle_wts[i] + particle_cum_wts[i - 1]; ties of the maximum go to the LAST particle (>=)
		double run = 0, max_wt = -1.79769313486231570815e308;
		int max_id = 0;
		for(int i = 0; i < n; ++i){
			const double v = ws[i];
			run = (i == 0) ? v : v + run;
			ws[i] = run;                                              // in place: the weight itself stays in pf.weights (smem case)
			if(v >= max_wt){ max_wt = v; max_id = i; }
		}
		pf.max_wt_id[obj] = max_id;
		s_total = run;
		s_resample = (pf.resampling_type != MTFB_PF_RESAMPLE_NONE) ? 1 : 0;
	}
	__syncthreads();
	const double total = s_total;
	if(pf.adaptive && s_resample){
		// n_eff = 1 / || w / sum(w) ||^2 (NT/PF.cc:391-396): pairwise sums here, Eigen's are vectorised -- both within an ulp
		// or two of the exact value, and the test is a threshold on a quantity of order n
		double ss[2] = { 0, 0 };
		for(int i = tid; i < n; i += blockDim.x){ const double v = w[i]; ss[0] += v; ss[1] = fma(v, v, ss[1]); }
		__shared__ double s_tmp[8 * 2];
		block_allreduce<2, 256>(ss, s_tmp);
		const double q = ss[1] / (ss[0] * ss[0]);
		const double n_eff = q == 0 ? 0 : 1.0 / q;
		if(tid == 0 && n_eff > pf.min_eff_particles) s_resample = 0;
		__syncthreads();
	}
	const bool resample = s_resample != 0;
	if(tid == 0) pf.resample_flag[obj] = resample ? 1 : 0;
	// particle_cum_wts /= particle_cum_wts[n - 1] only where resampling happens (NT/PF.cc:457, 512)
	for(int i = tid; i < n; i += blockDim.x) cw[i] = resample ? ws[i] / total : ws[i];
}

// the variant for particle counts whose weights do not fit in shared memory: the running sum goes straight to cum_weights
__global__ void __launch_bounds__(256) pf_weights_global_kernel(PFDev pf){
	const int obj = blockIdx.x, tid = threadIdx.x, n = pf.n_particles;
	if(pf.done[obj]) return;
	double *w = pf.weights + (size_t)obj*n, *cw = pf.cum_weights + (size_t)obj*n;
	const double *sim = pf.similarity + (size_t)obj*n;
	if(pf.likelihood_func != MTFB_PF_LIKELIHOOD_AM){
		for(int i = tid; i < n; i += blockDim.x){
			const double m = pf.max_similarity - sim[i];
			w[i] = (pf.likelihood_func == MTFB_PF_LIKELIHOOD_GAUSSIAN) ? pf.measurement_factor * exp(-0.5 * m / pf.measurement_sigma)
				: 1.0 / (1.0 + m);
		}
	}
	__syncthreads();
	__shared__ int s_resample;
	__shared__ double s_total;
	if(tid == 0){
		double run = 0, max_wt = -1.79769313486231570815e308;
		int max_id = 0;
		for(int i = 0; i < n; ++i){
			const double v = w[i];
			run = (i == 0) ? v : v + run;
			cw[i] = run;
			if(v >= max_wt){ max_wt = v; max_id = i; }
		}
		pf.max_wt_id[obj] = max_id;
		s_total = run;
		s_resample = (pf.resampling_type != MTFB_PF_RESAMPLE_NONE) ? 1 : 0;
	}
	__syncthreads();
	const double total = s_total;
	if(pf.adaptive && s_resample){
		double ss[2] = { 0, 0 };
		for(int i = tid; i < n; i += blockDim.x){ const double v = w[i]; ss[0] += v; ss[1] = fma(v, v, ss[1]); }
		__shared__ double s_tmp[8 * 2];
		block_allreduce<2, 256>(ss, s_tmp);
		const double q = ss[1] / (ss[0] * ss[0]);
		const double n_eff = q == 0 ? 0 : 1.0 / q;
		if(tid == 0 && n_eff > pf.min_eff_particles) s_resample = 0;
		__syncthreads();
	}
	const bool resample = s_resample != 0;
	if(tid == 0) pf.resample_flag[obj] = resample ? 1 : 0;
	if(resample) for(int i = tid; i < n; i += blockDim.x) cw[i] = cw[i] / total;
}

// one thread per new particle: the first particle whose normalised cumulative weight is >= the uniform deviate
// (binaryMultinomialResampling NT/PF.cc:463-476; linearMultinomialResampling's scan NT/PF.cc:520-521 stops at the same one)
template<int S> __global__ void __launch_bounds__(128) pf_resample_kernel(PFDev pf, unsigned iter_tag){
	const int obj = blockIdx.y, i = blockIdx.x*blockDim.x + threadIdx.x, n = pf.n_particles;
	if(i >= n || pf.done[obj] || !pf.resample_flag[obj]) return;
	const size_t pi = (size_t)obj*n + i;
	double u;
	if(pf.uniforms_in) u = pf.uniforms_in[pi];
	else{
		double u1;
		uniform2(pf.seed, (unsigned)i, (unsigned)(obj + pf.object_offset), iter_tag, STREAM_UNIFORMS, u, u1);
	}
	if(pf.uniforms_out) pf.uniforms_out[pi] = u;
	const double *cw = pf.cum_weights + (size_t)obj*n;
	int lower = 0, upper = n - 1, id = (lower + upper) / 2;
	while(upper > lower){
		if(cw[id] >= u) upper = id; else lower = id + 1;
		id = (lower + upper) / 2;
	}
	pf.src_id[pi] = id;
	const double *ss = pf.states + ((size_t)obj*n + id)*S, *sa = pf.ar + ((size_t)obj*n + id)*S;
	double *ds = pf.states_other + pi*S, *da = pf.ar_other + pi*S;
#pragma unroll
	for(int s = 0; s < S; ++s){ ds[s] = ss[s]; da[s] = sa[s]; }
}

// the resampled set becomes the current one (objects that resampled)
template<int S> __global__ void __launch_bounds__(256) pf_commit_kernel(PFDev pf){
	const int obj = blockIdx.y, n = pf.n_particles;
	if(pf.done[obj] || !pf.resample_flag[obj]) return;
	const size_t base = (size_t)obj*n*S;
	for(size_t e = (size_t)blockIdx.x*blockDim.x + threadIdx.x; e < (size_t)n*S; e += (size_t)gridDim.x*blockDim.x){
		pf.states[base + e] = pf.states_other[base + e];
		pf.ar[base + e] = pf.ar_other[base + e];
	}
}

// one CTA per object: the state the tracker reports, the corner-change test, reset_to_mean
template<int SSM> __global__ void __launch_bounds__(256) pf_mean_kernel(PFDev pf, DevBatch b){
	constexpr int S = StateSize<SSM>::value;
	const int obj = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = pf.n_particles;
	if(pf.done[obj]) return;
	__shared__ double s_red[8 * 8];
	__shared__ double s_state[S], s_corners[8];
	__shared__ int s_max_id;
	const double *states = pf.states + (size_t)obj*n*S;
	const double *w = pf.weights + (size_t)obj*n;
	if(pf.resample_flag[obj]){
		// max_wt_id after resampling: the LAST new particle whose source carries the largest resampled weight (NT/PF.cc:490-493)
		const int *src = pf.src_id + (size_t)obj*n;
		double best = -1.0; int best_i = -1;
		for(int i = tid; i < n; i += blockDim.x){
			const double v = w[src[i]];
			if(v > best || (v == best && i > best_i)){ best = v; best_i = i; }
		}
#pragma unroll
		for(int off = 16; off >= 1; off >>= 1){
			const double ov = __shfl_xor_sync(FULL_MASK, best, off); const int oi = __shfl_xor_sync(FULL_MASK, best_i, off);
			if(ov > best || (ov == best && oi > best_i)){ best = ov; best_i = oi; }
		}
		__shared__ double s_bv[8]; __shared__ int s_bi[8];
		if(lane == 0){ s_bv[warp] = best; s_bi[warp] = best_i; }
		__syncthreads();
		if(tid == 0){
			for(int k = 1; k < 8; ++k) if(s_bv[k] > best || (s_bv[k] == best && s_bi[k] > best_i)){ best = s_bv[k]; best_i = s_bi[k]; }
			s_max_id = best_i; pf.max_wt_id[obj] = best_i;
		}
	} else if(tid == 0) s_max_id = pf.max_wt_id[obj];
	__syncthreads();
	if(pf.mean_type == MTFB_PF_MEAN_NONE){
		if(tid < S) s_state[tid] = states[(size_t)s_max_id*S + tid];                 // NT/PF.cc:414-420
	} else if(pf.mean_type == MTFB_PF_MEAN_SSM){
		// ProjectiveBase::estimateMeanOfSamples (ProjectiveBase.cc:308-314) is a running mean; the plain mean agrees with it to
		// rounding (tests: 1e-12 relative)
		double acc[S];
#pragma unroll
		for(int s = 0; s < S; ++s) acc[s] = 0;
		for(int i = tid; i < n; i += blockDim.x){
#pragma unroll
			for(int s = 0; s < S; ++s) acc[s] += states[(size_t)i*S + s];
		}
		block_allreduce<S, 256>(acc, s_red);
		if(tid < S){
			double v = 0;
#pragma unroll
			for(int s = 0; s < S; ++s) if(s == tid) v = acc[s];
			s_state[tid] = v / n;
		}
	} else{
		// updateMeanCorners (NT/PF.cc:587-594): mean of the corners the particles' states give
		double acc[8];
#pragma unroll
		for(int k = 0; k < 8; ++k) acc[k] = 0;
		for(int i = tid; i < n; i += blockDim.x){
			const Mat3 W = warp_from_state<SSM>(states + (size_t)i*S);
			double c[8];
			warp_corners<SSM>(W, b.init_corners + (size_t)obj * 8, c);
#pragma unroll
			for(int k = 0; k < 8; ++k) acc[k] += c[k];
		}
		block_allreduce<8, 256>(acc, s_red);
		if(tid < 8){
			double v = 0;
#pragma unroll
			for(int k = 0; k < 8; ++k) if(k == tid) v = acc[k];
			s_corners[tid] = v / n;
		}
	}
	__syncthreads();
	if(warp == 0){
		double nc[8];
		if(pf.mean_type == MTFB_PF_MEAN_CORNERS){
			// ssm->setCorners(mean_corners): new DLT, identity warp, zero state, init_corners = the mean corners
			double c_in[8];
#pragma unroll
			for(int k = 0; k < 8; ++k){ c_in[k] = s_corners[k]; nc[k] = c_in[k]; }
			set_corners<SSM>(b, obj, lane, c_in);
		} else{
			const Mat3 W = warp_from_state<SSM>(s_state);                                // ssm->setState
			warp_corners<SSM>(W, b.init_corners + (size_t)obj * 8, nc);
			if(lane < 9) b.warp[(size_t)obj * 9 + lane] = W.m[lane];
			if(lane < S) b.state[(size_t)obj*S + lane] = s_state[lane];
#pragma unroll
			for(int k = 0; k < 8; ++k) if(lane == k) b.corners[(size_t)obj * 8 + k] = nc[k];
		}
		double upd = 0;
#pragma unroll
		for(int k = 0; k < 8; ++k){ const double d = pf.prev_corners[(size_t)obj * 8 + k] - nc[k]; upd += d*d; }
		__syncwarp();
#pragma unroll
		for(int k = 0; k < 8; ++k) if(lane == k) pf.prev_corners[(size_t)obj * 8 + k] = nc[k];
		if(lane == 0){
			b.n_iters[obj] += 1;
			if(upd < pf.epsilon) pf.done[obj] = 1;                                      // NT/PF.cc:430-434
			b.f[obj] = pf.similarity[(size_t)obj*n + (pf.resample_flag[obj] ? pf.src_id[(size_t)obj*n + s_max_id] : s_max_id)];
		}
	}
}

// initializeParticles (NT/PF.cc:185-198): every particle at the tracker's state, uniform weights, zero auto-regression term
template<int S> __global__ void __launch_bounds__(256) pf_init_particles_kernel(PFDev pf, DevBatch b, int set_prev){
	const int obj = blockIdx.y, n = pf.n_particles;
	const double init_wt = 1.0 / n;
	for(int i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x){
		const size_t pi = (size_t)obj*n + i;
#pragma unroll
		for(int s = 0; s < S; ++s){ pf.states[pi*S + s] = b.state[(size_t)obj*S + s]; pf.ar[pi*S + s] = 0; }
		pf.weights[pi] = init_wt;
		pf.cum_weights[pi] = init_wt * (i + 1);
	}
	if(set_prev && blockIdx.x == 0 && threadIdx.x < 8) pf.prev_corners[(size_t)obj * 8 + threadIdx.x] = b.corners[(size_t)obj * 8 + threadIdx.x];
}

__global__ void pf_frame_begin_kernel(PFDev pf, DevBatch b){
	const int obj = blockIdx.x*blockDim.x + threadIdx.x;
	if(obj >= b.P) return;
	pf.done[obj] = 0; b.n_iters[obj] = 0;
}

} // namespace

cudaError_t launch_pf_init_particles(int ssm, const PFDev &pf, const DevBatch &b, bool set_prev_corners, cudaStream_t st){
	const dim3 grid((pf.n_particles + 255) / 256, b.P);
	if(ssm == SSM_HOM) pf_init_particles_kernel<8><<<grid, 256, 0, st>>>(pf, b, set_prev_corners ? 1 : 0);
	else pf_init_particles_kernel<6><<<grid, 256, 0, st>>>(pf, b, set_prev_corners ? 1 : 0);
	return cudaGetLastError();
}
cudaError_t launch_pf_frame_begin(const PFDev &pf, const DevBatch &b, cudaStream_t st){
	pf_frame_begin_kernel<<<(b.P + 127) / 128, 128, 0, st>>>(pf, b);
	return cudaGetLastError();
}
cudaError_t launch_pf_perturb(int ssm, const PFDev &pf, const DevBatch &b, unsigned iter_tag, cudaStream_t st){
	if(ssm == SSM_HOM && pf.corner_based){
		const dim3 grid((pf.n_particles + 127) / 128, b.P);
		pf_perturb_corner_kernel<<<grid, 128, 0, st>>>(pf, b, iter_tag);
	} else{
		const dim3 grid((pf.n_particles + 127) / 128, b.P);
		if(ssm == SSM_HOM) pf_perturb_kernel<SSM_HOM><<<grid, 128, 0, st>>>(pf, b, iter_tag);
		else pf_perturb_kernel<SSM_AFF><<<grid, 128, 0, st>>>(pf, b, iter_tag);
	}
	return cudaGetLastError();
}
cudaError_t launch_pf_weights(const PFDev &pf, const DevBatch &b, cudaStream_t st){
	if(pf.weights_in_smem){
		const size_t dyn = (size_t)pf.n_particles*sizeof(double);
		static bool configured = false;
		if(!configured){
			cudaError_t e = cudaFuncSetAttribute(pf_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
			if(e != cudaSuccess) return e;
			configured = true;
		}
		pf_weights_kernel<<<b.P, 256, dyn, st>>>(pf);
	} else pf_weights_global_kernel<<<b.P, 256, 0, st>>>(pf);
	return cudaGetLastError();
}
cudaError_t launch_pf_resample(int ssm, const PFDev &pf, const DevBatch &b, unsigned iter_tag, cudaStream_t st){
	const dim3 grid((pf.n_particles + 127) / 128, b.P);
	const dim3 cgrid(8, b.P);
	if(ssm == SSM_HOM){
		pf_resample_kernel<8><<<grid, 128, 0, st>>>(pf, iter_tag);
		pf_commit_kernel<8><<<cgrid, 256, 0, st>>>(pf);
	} else{
		pf_resample_kernel<6><<<grid, 128, 0, st>>>(pf, iter_tag);
		pf_commit_kernel<6><<<cgrid, 256, 0, st>>>(pf);
	}
	return cudaGetLastError();
}
cudaError_t launch_pf_mean(int ssm, const PFDev &pf, const DevBatch &b, cudaStream_t st){
	if(ssm == SSM_HOM) pf_mean_kernel<SSM_HOM><<<b.P, 256, 0, st>>>(pf, b);
	else pf_mean_kernel<SSM_AFF><<<b.P, 256, 0, st>>>(pf, b);
	return cudaGetLastError();
}

} // namespace mtfb
