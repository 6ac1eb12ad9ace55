// lookback.cuh -- exclusive prefix sums ACROSS the blocks of one kernel launch, without a second kernel ("decoupled look-back").
// Used where the reference orders its outputs by atomicAdd arrival (compaction of the samples the loss keeps, testbed_nerf.cu:1682) and we
// want the deterministic order instead (ray-slot order) without paying for a scan kernel between two passes over the same rays.
//
// A tile is one block's worth of consecutive elements.  Tiles are numbered by a ticket taken when the block starts, so every predecessor
// of a running tile is itself running (or done): waiting for its descriptor cannot deadlock, whatever order the hardware starts blocks in.
// Descriptor (one aligned 64-bit word, written and read with single accesses): [epoch:30 | state:2 | value:32].
//   state 1 = "value is this tile's own sum", state 2 = "value is the inclusive prefix up to and including this tile".
// A descriptor counts only if its epoch is the launch's: the array is never cleared, the host bumps the epoch per launch (>= 1, < 2^30;
// freshly zeroed memory reads as epoch 0).  The block holding the last ticket puts the ticket counter back to 0.
#pragma once
#include <stdint.h>

struct LookbackScan { unsigned long long* desc; uint32_t* ticket; uint32_t epoch; };

// one thread of the block
__device__ __forceinline__ uint32_t lookback_take_ticket(const LookbackScan& s) {
	const uint32_t t = atomicAdd(s.ticket, 1u);
	if (t == gridDim.x - 1) atomicExch(s.ticket, 0u);   // every block of this launch has drawn by now
	return t;
}

// Warp collective (all 32 lanes, same arguments): publishes `aggregate` as the sum of tile `tile`, returns the sum of all tiles before it.
__device__ __forceinline__ uint32_t lookback_exclusive(const LookbackScan& s, uint32_t tile, uint32_t aggregate, int lane) {
	volatile unsigned long long* d = s.desc;
	const unsigned long long tag = (unsigned long long)s.epoch << 34;
	if (tile == 0) {
		if (lane == 0) d[0] = tag | (2ull << 32) | aggregate;
		return 0u;
	}
	if (lane == 0) d[tile] = tag | (1ull << 32) | aggregate;
	uint32_t exclusive = 0;
	int p = (int)tile - 1;   // the window [p - 31, p], nearest predecessor in lane 0
	while (true) {
		const int idx = p - lane;
		uint32_t state = 2u, val = 0u;   // in front of tile 0: an inclusive prefix of 0
		if (idx >= 0) {
			unsigned long long v;
			do { v = d[idx]; } while ((uint32_t)(v >> 34) != s.epoch || ((uint32_t)(v >> 32) & 3u) == 0u);
			state = (uint32_t)(v >> 32) & 3u; val = (uint32_t)v;
		}
		const uint32_t have_prefix = __ballot_sync(0xffffffffu, state == 2u);
		const int last = have_prefix ? __ffs((int)have_prefix) - 1 : 31;   // sum up to (and including) the nearest inclusive prefix
		uint32_t c = lane <= last ? val : 0u;
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
		exclusive += c;
		if (have_prefix) break;
		p -= 32;
	}
	if (lane == 0) d[tile] = tag | (2ull << 32) | (unsigned long long)(exclusive + aggregate);
	return exclusive;
}
