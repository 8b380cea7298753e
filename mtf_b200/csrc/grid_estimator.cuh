// grid_estimator.cuh -- robust warp estimation from point pairs on the device (grid_estimator.cu): the step
// GridTracker::update runs after its cells (SM/src/GridTracker.cc:253-269).
#pragma once
#include <cuda_runtime.h>

namespace mtfb {

enum { EST_MAX_MODEL_PTS = 8 };

// all pointers are device pointers
struct EstDev {
	const float *in_pts, *out_pts;     // n x 2, interleaved (std::vector<cv::Point2f>)
	int n, homography;                 // homography: 3x3 model with H[8] = 1, 8 parameters; else 2x3, 6 parameters
	int method, model_points, refine, max_iters, max_attempts, lm_max_iters;
	double thresh, confidence;
	unsigned long long seed;
	double *out;                       // 9 warp matrix | 8 state update | result, hypotheses drawn, inliers, LM evaluations
	unsigned char *mask;               // n
	float *err;                        // EST_WARPS x n floats (LMedS: the reprojection errors of a warp's hypothesis)
};

enum { EST_THREADS = 256, EST_WARPS = 8 };

// centroids of P quadrilaterals (x0..x3, y0..y3) as cv::Point2f (utils::getCentroid, miscUtils.h:473-480)
cudaError_t launch_centroids(const double *corners, int P, float *pts, cudaStream_t st);
cudaError_t launch_estimate(const EstDev &e, cudaStream_t st);
// GridTracker::update after the estimation, on the device (SM/src/GridTracker.cc:271-274 + resetTrackers :345-392 with
// patch_centroid_inside = 1, dyn_patch_size = 0): region <- ssm.applyWarpToCorners(region, state_update); the (gx + 1) x (gy + 1)
// grid of the region (utils::getPtsFromCorners: the unit square under the 4-corner DLT); every cell's region = the
// patch_size box around the centroid of its four grid points.  d_region: 8 doubles (x0..x3, y0..y3), updated in place;
// d_state_update: the estimator's output; d_cells: gx * gy x 8.
cudaError_t launch_grid_layout(int homography, const double *d_state_update, double *d_region, int gx, int gy, double psx, double psy,
	double *d_cells, cudaStream_t st);

} // namespace mtfb
