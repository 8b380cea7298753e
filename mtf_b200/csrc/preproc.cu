// preproc.cu -- the frame's pre-processing on the device: what utils::PreProcBase::processFrame does for
// output_type CV_32FC1 with the default GaussianSmoothing (Utilities/src/preprocUtils.cc:108-127,
// Utilities/include/mtf/Utilities/preprocUtils.h:67-78, Config/include/mtf/Config/parameters.h:229-235):
//     frame_raw.convertTo(CV_32F) -> cv::cvtColor(BGR2GRAY) -> cv::GaussianBlur(5 x 5, sigma)
// fused into one kernel behind the H2D copy of the raw uint8 frame, so that a frame costs 1 (gray) or 3 (BGR) bytes per
// pixel on PCIe instead of 4, and no host pass at all.  The arithmetic is OpenCV's (the test suite's CPU restatement
// uses the same operation order; no contraction: this file is compiled -fmad=false): float gray weights, float
// symmetric 5-tap row pass then column pass, BORDER_REFLECT_101.
//
// One CTA = a 64 x 16 tile of the output: gray values of the tile + 2-pixel halo go to shared memory once
// (coalesced uint8 reads), the row pass writes a second shared array, the column pass writes the frame.
// HBM-bound: 1-3 B read + 4 B written per pixel.
#include "lk_kernels.cuh"

namespace mtfb {

namespace {
constexpr int TW = 64, TH = 16, R = 2;

__device__ __forceinline__ int reflect101(int i, int n){
	if(n == 1) return 0;
	while(i < 0 || i >= n){ if(i < 0) i = -i; else i = 2 * n - 2 - i; }
	return i;
}

template<int CH> __global__ void __launch_bounds__(256) preproc_gauss5_kernel(const unsigned char *__restrict__ src, int src_pitch,
	float *__restrict__ dst, int dst_pitch, int h, int w, float k0, float k1, float k2){
	__shared__ float s_gray[(TH + 2 * R)][TW + 2 * R + 1];
	__shared__ float s_row[(TH + 2 * R)][TW + 1];
	const int x0 = blockIdx.x*TW, y0 = blockIdx.y*TH;
	for(int i = threadIdx.x; i < (TH + 2 * R)*(TW + 2 * R); i += 256){
		const int ty = i / (TW + 2 * R), tx = i - ty*(TW + 2 * R);
		const int y = reflect101(y0 + ty - R, h), x = reflect101(x0 + tx - R, w);
		const unsigned char *p = src + (size_t)y*src_pitch + (size_t)x*CH;
		float v;
		if(CH == 3){
			// cv::cvtColor(CV_BGR2GRAY) on CV_32FC3: B*0.114f + G*0.587f + R*0.299f
			const float b = p[0], g = p[1], r = p[2];
			v = b*0.114f; v = v + g*0.587f; v = v + r*0.299f;
		} else v = p[0];
		s_gray[ty][tx] = v;
	}
	__syncthreads();
	for(int i = threadIdx.x; i < (TH + 2 * R)*TW; i += 256){
		const int ty = i / TW, tx = i - ty*TW;
		const float *a = &s_gray[ty][tx + R];
		float s = a[0] * k0;
		s = s + (a[-1] + a[1])*k1;
		s = s + (a[-2] + a[2])*k2;
		s_row[ty][tx] = s;
	}
	__syncthreads();
	for(int i = threadIdx.x; i < TH*TW; i += 256){
		const int ty = i / TW, tx = i - ty*TW;
		const int x = x0 + tx, y = y0 + ty;
		if(x < w && y < h){
			float s = s_row[ty + R][tx] * k0;
			s = s + (s_row[ty + R - 1][tx] + s_row[ty + R + 1][tx])*k1;
			s = s + (s_row[ty + R - 2][tx] + s_row[ty + R + 2][tx])*k2;
			dst[(size_t)y*dst_pitch + x] = s;
		}
	}
}
} // namespace

// Note on the halo rows: the row pass of a halo row y' = reflect101(y, h) is computed from that row's own reflected
// neighbourhood, which is exactly the row-pass value the reference's column pass reads at the reflected index.
cudaError_t launch_preproc_gauss5(const unsigned char *d_src, int src_pitch, int channels, float *d_dst, int dst_pitch, int h, int w,
	const float *k5, cudaStream_t st){
	const dim3 grid((w + TW - 1) / TW, (h + TH - 1) / TH);
	if(channels == 3) preproc_gauss5_kernel<3><<<grid, 256, 0, st>>>(d_src, src_pitch, d_dst, dst_pitch, h, w, k5[2], k5[1], k5[0]);
	else if(channels == 1) preproc_gauss5_kernel<1><<<grid, 256, 0, st>>>(d_src, src_pitch, d_dst, dst_pitch, h, w, k5[2], k5[1], k5[0]);
	else return cudaErrorInvalidValue;
	return cudaGetLastError();
}

} // namespace mtfb
