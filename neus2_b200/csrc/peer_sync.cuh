// peer_sync.cuh -- device side of the cross-GPU ordering points of the data-parallel step (see peer_sync.cu): system-scope stores /
// loads on the peers' mapped sync blocks, executed by ONE warp -- either a tiny kernel of its own or the last block of the kernel that
// produces what the peers wait for (the training kernel, the sharded optimizer), which needs no launch slot of its own.
#pragma once
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

// Exchange by one warp (world <= 8 <= 32): every lane < PEER_PAYLOAD_WORDS stores its payload word into this rank's slot of every
// peer's block, lane w releases the flag to peer w and then waits for peer w's flag in this rank's own block; finally the payload words
// are summed over ranks in rank order (identical bits on every rank).  Null pointers = that part of the payload is not exchanged.
__device__ __forceinline__ void peer_exchange_warp(const PeerXfer& x, const uint32_t* __restrict__ counts_local, uint32_t* __restrict__ counts_global,
                                                   float* __restrict__ loss_sums, float* __restrict__ delta_acc, int lane) {
	const uint32_t buf = x.epoch & 1u;
	uint32_t v = 0u;
	if (lane < 4) v = counts_local ? counts_local[lane] : 0u;
	else if (lane < 8) v = loss_sums ? __float_as_uint(loss_sums[lane - 4]) : 0u;
	else if (lane < PEER_PAYLOAD_WORDS) v = delta_acc ? __float_as_uint(delta_acc[lane - 8]) : 0u;
	if (lane < PEER_PAYLOAD_WORDS) for (uint32_t w = 0; w < x.world; ++w) st_relaxed_sys(&x.peers.block[w]->exchange[buf][x.rank].payload[lane], v);
	__syncwarp();
	if ((uint32_t)lane < x.world) {
		__threadfence_system();
		st_release_sys(&x.peers.block[lane]->exchange[buf][x.rank].flag, x.epoch);
		spin_until(&x.peers.block[x.rank]->exchange[buf][lane].flag, x.epoch, lane);
	}
	__syncwarp();
	if (lane < PEER_PAYLOAD_WORDS) {
		const PeerSlot* mine = x.peers.block[x.rank]->exchange[buf];
		if (lane < 4) {
			uint32_t s = 0u;
			for (uint32_t r = 0; r < x.world; ++r) s += ld_relaxed_sys(&mine[r].payload[lane]);
			if (counts_local) counts_global[lane] = s;
		} else {
			float s = 0.f;
			for (uint32_t r = 0; r < x.world; ++r) s += __uint_as_float(ld_relaxed_sys(&mine[r].payload[lane]));
			if (lane < 8) { if (loss_sums) loss_sums[lane - 4] = s; }
			else if (delta_acc) delta_acc[lane - 8] = s;
		}
	}
	__syncwarp();
}

// Barrier by one warp (its own flags and epoch counter).
__device__ __forceinline__ void peer_barrier_warp(const PeerXfer& x, int lane) {
	if ((uint32_t)lane < x.world) {
		__threadfence_system();
		st_release_sys(&x.peers.block[lane]->barrier[x.rank].flag, x.epoch);
		spin_until(&x.peers.block[x.rank]->barrier[lane].flag, x.epoch, lane);
	}
	__syncwarp();
}
