// pf_tracker.cuh -- device-side view of a particle filter context (pf_tracker.cu) and its launchers.
#pragma once
#include "lk_kernels.cuh"

namespace mtfb {

// all pointers are device pointers; particle arrays are [object][particle][...]
struct PFDev {
	int n_particles, n_normals;            // n_normals: deviates per particle and iteration (S, or 10 with corner based sampling)
	double *states, *ar;                   // P x n x S   current particle set, auto-regression terms
	double *states_other, *ar_other;       // P x n x S   the set resampling writes
	double *weights, *cum_weights;         // P x n       particle_wts, particle_cum_wts
	double *similarity;                    // P x n       am->getSimilarity() of every particle
	int *src_id;                           // P x n       resample_id of every new particle
	int *max_wt_id, *resample_flag, *done; // P
	double *prev_corners;                  // P x 8
	const double *normals_in, *uniforms_in;    // host-supplied random stream for this update, else null (device generator)
	double *normals_out, *uniforms_out;        // the deviates used (record_randoms), else null
	double sigma[8], mean[8];
	int dynamic_model, update_type, likelihood_func, resampling_type, mean_type;
	int corner_based, adaptive, weights_in_smem;
	double min_eff_particles, measurement_factor, measurement_sigma, max_similarity, ar_coeff, epsilon;
	unsigned long long seed;
	int object_offset;                     // index of this context's first object in the whole (sharded) job: keeps the
	                                       // random streams of the objects independent of how they are split over GPUs
};

cudaError_t launch_pf_init_particles(int ssm, const PFDev &pf, const DevBatch &b, bool set_prev_corners, cudaStream_t st);
cudaError_t launch_pf_frame_begin(const PFDev &pf, const DevBatch &b, cudaStream_t st);
cudaError_t launch_pf_perturb(int ssm, const PFDev &pf, const DevBatch &b, unsigned iter_tag, cudaStream_t st);
cudaError_t launch_pf_weights(const PFDev &pf, const DevBatch &b, cudaStream_t st);
cudaError_t launch_pf_resample(int ssm, const PFDev &pf, const DevBatch &b, unsigned iter_tag, cudaStream_t st);
cudaError_t launch_pf_mean(int ssm, const PFDev &pf, const DevBatch &b, cudaStream_t st);

} // namespace mtfb
