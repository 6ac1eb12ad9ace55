// peer_sync.cu -- cross-GPU synchronisation of the data-parallel training step over NVLink peer memory (one NVSwitch domain).
//
// The reference is single-GPU; the data-parallel step needs two ordering points per step and one tiny reduction:
//   exchange 1 (behind the loss kernel): the sum over ranks of {rays kept, samples, compacted samples} and, every 16th step, of the
//             three loss sums -- what the host needs to size the next step, available while the backward pass still runs;
//   exchange 2 (behind the backward pass): "every rank's gradients are complete" + the DeltaNetwork's 9 gradient sums;
//   barrier   (behind the sharded optimizer): "every rank's slice has been stored into every rank's buffers" before the next forward.
// These were stream-ordered 4..48-byte NCCL all-reduces (30-40 us each at 8 ranks).  Here each rank owns a small sync block that all
// peers map (CUDA IPC): a rank stores its payload and then an epoch flag into ITS slot of every peer's block (st.release.sys over
// NVLink) and spins on the W flags of its own block (ld.acquire.sys, local memory).  One CTA, one warp per peer, a few microseconds.
// Sums are taken in rank order on every rank, so all ranks hold bit-identical results.
#include "ngp_common.cuh"
#include "ngp_kernels.h"
#include <cstdio>

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) { asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) { uint32_t v; asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// A peer that never arrives (a rank died) must not hang the GPU: after ~60 s of spinning the kernel traps, which fails every later
// CUDA call of the process loudly.
#define PEER_SPIN_LIMIT_CYCLES (120ll * 1000 * 1000 * 1000)
__device__ __forceinline__ void spin_until(const uint32_t* flag, uint32_t epoch, int peer) {
	const long long t0 = clock64();
	while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
		__nanosleep(64);
		if (clock64() - t0 > PEER_SPIN_LIMIT_CYCLES) {
			printf("neus2_b200: peer %d did not reach sync epoch %u within the spin limit\n", peer, epoch);
			__trap();
		}
	}
}

// blockDim = 32 * world.  Warp w talks to peer w.
// counts_local / loss_sums / delta_acc: null = that part of the payload is not exchanged by this call.
__global__ void k_peer_exchange(uint32_t epoch, uint32_t rank, uint32_t world, PeerSyncPtrs peers, const uint32_t* __restrict__ counts_local,
                                uint32_t* __restrict__ counts_global, float* __restrict__ loss_sums, float* __restrict__ delta_acc) {
	const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t buf = epoch & 1u;
	const bool with_loss = loss_sums != nullptr;
	if (w < world) {
		uint32_t v = 0u;
		if (lane < 4) v = counts_local ? counts_local[lane] : 0u;
		else if (lane < 8) v = with_loss ? __float_as_uint(loss_sums[lane - 4]) : 0u;
		else if (lane < PEER_PAYLOAD_WORDS) v = delta_acc ? __float_as_uint(delta_acc[lane - 8]) : 0u;
		PeerSlot* dst = &peers.block[w]->exchange[buf][rank];
		if (lane < PEER_PAYLOAD_WORDS) st_relaxed_sys(&dst->payload[lane], v);
		__syncwarp();
		if (lane == 0) { __threadfence_system(); st_release_sys(&dst->flag, epoch); }
		if (lane == 0) spin_until(&peers.block[rank]->exchange[buf][w].flag, epoch, (int)w);
		__syncwarp();
	}
	__syncthreads();
	if (threadIdx.x < PEER_PAYLOAD_WORDS) {
		const PeerSlot* mine = peers.block[rank]->exchange[buf];
		const uint32_t t = threadIdx.x;
		if (t < 4) {
			uint32_t s = 0u;
			for (uint32_t r = 0; r < world; ++r) s += ld_relaxed_sys(&mine[r].payload[t]);
			if (counts_local) counts_global[t] = s;
		} else {
			float s = 0.f;
			for (uint32_t r = 0; r < world; ++r) s += __uint_as_float(ld_relaxed_sys(&mine[r].payload[t]));
			if (t < 8) { if (with_loss) loss_sums[t - 4] = s; }
			else if (delta_acc) delta_acc[t - 8] = s;
		}
	}
}

// Plain barrier (its own flags and epoch counter).  Optionally clears small device arrays of the caller once everybody has arrived
// (the step's accumulators and counters: nothing of this step reads them any more, the next step expects zeros).
__global__ void k_peer_barrier(uint32_t epoch, uint32_t rank, uint32_t world, PeerSyncPtrs peers, PeerZero z) {
	const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (w < world && lane == 0) {
		__threadfence_system();
		st_release_sys(&peers.block[w]->barrier[rank].flag, epoch);
		spin_until(&peers.block[rank]->barrier[w].flag, epoch, (int)w);
	}
	__syncthreads();
	#pragma unroll
	for (int a = 0; a < 3; ++a) for (uint32_t i = threadIdx.x; i < z.n[a]; i += blockDim.x) z.p[a][i] = 0u;
}

int launch_peer_exchange(cudaStream_t stream, uint32_t epoch, uint32_t rank, uint32_t world, const PeerSyncPtrs& peers, const uint32_t* counts_local,
                         uint32_t* counts_global, float* loss_sums, float* delta_acc) {
	if (world < 1 || world > NEUS2_MAX_PEERS || rank >= world) { neus_set_error("bad rank / world size for the peer exchange", __FILE__, __LINE__); return -1; }
	k_peer_exchange<<<1, 32 * world, 0, stream>>>(epoch, rank, world, peers, counts_local, counts_global, loss_sums, delta_acc);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}

int launch_peer_barrier(cudaStream_t stream, uint32_t epoch, uint32_t rank, uint32_t world, const PeerSyncPtrs& peers, const PeerZero& z) {
	if (world < 1 || world > NEUS2_MAX_PEERS || rank >= world) { neus_set_error("bad rank / world size for the peer barrier", __FILE__, __LINE__); return -1; }
	k_peer_barrier<<<1, 32 * world, 0, stream>>>(epoch, rank, world, peers, z);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}
