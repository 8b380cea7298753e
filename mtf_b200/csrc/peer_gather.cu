// peer_gather.cu -- the per-frame all-gather of the batch's corners over NVLink peer memory, without a collective library.
//
// One process per GPU; every rank tracks a contiguous range of the job's patches (patches never interact inside
// a frame, the P x 8 corners are what the host logic of every rank needs once per frame: GridTracker's estimation,
// SM/src/GridTracker.cc:265-274).  Each rank owns two (n_total x 8) fp64 arrays (frames alternate between them) and a row
// of sequence flags, exported through CUDA IPC (mtfb_peer_export) and mapped by the others (mtfb_peer_attach).  A frame:
//   1. the update kernel stores each patch's final corners into the current array of EVERY rank (store_patch_state,
//      lk_solve.cuh) -- the transfer rides on the kernel's own epilogue, 64 bytes per patch and peer;
//   2. peer_signal_wait_kernel (one warp, same stream, so after every CTA of 1.): lane r publishes this frame's sequence
//      number in rank r's flag row (release, system scope) and waits until rank r's number has arrived in its own row.
// When that kernel has finished, the local gathered array holds the corners of all the job's patches.  Two arrays suffice:
// a rank can start frame t + 2 (which overwrites array t & 1 everywhere) only after every peer has signalled frame t + 1,
// and a peer signals t + 1 only after its stream is done with frame t.
#include <cuda_runtime.h>
#include "lk_kernels.cuh"

namespace mtfb {

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v){
	asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p){
	unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

// corners written by something else than an update kernel (initialize, setRegion, the particle filter's mean)
__global__ void peer_push_kernel(const double *__restrict__ corners, int n, PeerOut peers){
	const int i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= n) return;
	const double c = corners[i];
	for(int r = 0; r < peers.n; ++r) peers.dst[r][(size_t)peers.row0 * 8 + i] = c;
}

struct PeerFlags { unsigned *row[MTFB_MAX_PEERS]; };
__global__ void peer_signal_wait_kernel(PeerFlags peer, unsigned *my_flags, int rank, int world, unsigned seq, int *d_err){
	const int r = threadIdx.x;
	if(r >= world) return;
	__threadfence_system();
	st_release_sys(peer.row[r] + rank, seq);
	// sequence numbers only grow; (int)(have - seq) >= 0 survives the wrap-around
	unsigned long long t0, t1;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	while((int)(ld_acquire_sys(my_flags + r) - seq) < 0){
		__nanosleep(200);
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
		if(t1 - t0 > 5000000000ull){ atomicExch(d_err, 1 + r); break; }          // 5 s: a peer died; report instead of hanging
	}
}

cudaError_t launch_peer_push(const double *corners, int P, const PeerOut &peers, cudaStream_t st){
	const int n = P * 8;
	peer_push_kernel<<<(n + 255) / 256, 256, 0, st>>>(corners, n, peers);
	return cudaGetLastError();
}
cudaError_t launch_peer_signal_wait(unsigned *const *peer_flags, unsigned *my_flags, int rank, int world, unsigned seq, int *d_err,
	cudaStream_t st){
	PeerFlags pf;
	for(int r = 0; r < MTFB_MAX_PEERS; ++r) pf.row[r] = r < world ? peer_flags[r] : nullptr;
	peer_signal_wait_kernel<<<1, 32, 0, st>>>(pf, my_flags, rank, world, seq, d_err);
	return cudaGetLastError();
}

} // namespace mtfb
