// peer_sync.cu -- cross-GPU synchronisation of the data-parallel training step over NVLink peer memory (one NVSwitch domain).
//
// The reference is single-GPU; the data-parallel step needs two ordering points per step and one tiny reduction:
//   exchange 1 (behind the loss kernel): the sum over ranks of {rays kept, samples, compacted samples} and, every 16th step, of the
//             three loss sums -- what the host needs to size the next step, available while the backward pass still runs;
//   exchange 2 (behind the backward pass): "every rank's gradients are complete" + the DeltaNetwork's 9 gradient sums;
//   barrier   (behind the sharded optimizer): "every rank's slice has been stored into every rank's buffers" before the next forward.
// These were stream-ordered 4..48-byte NCCL all-reduces (30-40 us each at 8 ranks).  Here each rank owns a small sync block that all
// peers map (CUDA IPC): a rank stores its payload and then an epoch flag into ITS slot of every peer's block (st.release.sys over
// NVLink) and spins on the W flags of its own block (ld.acquire.sys, local memory).  One warp (peer_sync.cuh), a few microseconds.
// Sums are taken in rank order on every rank, so all ranks hold bit-identical results.
#include "ngp_common.cuh"
#include "peer_sync.cuh"

// One warp each; the kernels exist for the sync points that no other kernel can carry (reporting steps, whose loss sums are reduced by
// kernels of their own; the in-order step; frames whose canonical optimizer is frozen).
__global__ void k_peer_exchange(PeerXfer x, const uint32_t* __restrict__ counts_local, uint32_t* __restrict__ counts_global, float* __restrict__ loss_sums,
                                float* __restrict__ delta_acc) {
	peer_exchange_warp(x, counts_local, counts_global, loss_sums, delta_acc, (int)threadIdx.x);
}

// Optionally clears small device arrays of the caller once everybody has arrived (the step's accumulators and counters: nothing of this
// step reads them any more, the next step expects zeros).
__global__ void k_peer_barrier(PeerXfer x, PeerZero z) {
	peer_barrier_warp(x, (int)threadIdx.x);
	#pragma unroll
	for (int a = 0; a < 3; ++a) for (uint32_t i = threadIdx.x; i < z.n[a]; i += blockDim.x) z.p[a][i] = 0u;
}

int launch_peer_exchange(cudaStream_t stream, const PeerXfer& x, const uint32_t* counts_local, uint32_t* counts_global, float* loss_sums, float* delta_acc) {
	if (x.world < 1 || x.world > NEUS2_MAX_PEERS || x.rank >= x.world) { neus_set_error("bad rank / world size for the peer exchange", __FILE__, __LINE__); return -1; }
	k_peer_exchange<<<1, 32, 0, stream>>>(x, counts_local, counts_global, loss_sums, delta_acc);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}

int launch_peer_barrier(cudaStream_t stream, const PeerXfer& x, const PeerZero& z) {
	if (x.world < 1 || x.world > NEUS2_MAX_PEERS || x.rank >= x.world) { neus_set_error("bad rank / world size for the peer barrier", __FILE__, __LINE__); return -1; }
	k_peer_barrier<<<1, 32, 0, stream>>>(x, z);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}
