// density_grid.cu -- occupancy-grid maintenance (compiled with -fmad=false).
// Restates generate_grid_samples_nerf_nonuniform, splat_grid_samples_nerf_max_nearest_neighbor, ema_grid_samples_nerf,
// grid_to_bitfield, bitfield_max_pool (src/testbed_nerf.cu:640-795) and update_density_grid_mean_and_bitfield (:3377-3397).
#include "ngp_common.cuh"
#include "ngp_kernels.h"
#include "lookback.cuh"

#define G3 (NERF_GRIDSIZE * NERF_GRIDSIZE * NERF_GRIDSIZE)

__global__ void k_grid_samples(uint32_t n, Pcg32 rng, uint32_t step, Aabb bb, const float* __restrict__ grid_in, uint32_t n_cascades, float thresh,
                               float* __restrict__ positions, uint32_t* __restrict__ indices) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	rng.advance((int64_t)(uint32_t)(i * 4u));
	const uint32_t level = (uint32_t)(rng.next_float() * n_cascades) % n_cascades;
	uint32_t idx = 0;
	for (uint32_t j = 0; j < 10; ++j) {
		idx = ((i + step * n) * 56924617u + j * 19349663u + 96925573u) % G3;
		idx += level * G3;
		if (grid_in[idx] > thresh) break;
	}
	const uint32_t pos_idx = idx % G3;
	const uint32_t c[3] = {morton3D_invert(pos_idx >> 0), morton3D_invert(pos_idx >> 1), morton3D_invert(pos_idx >> 2)};
	const float r[3] = {rng.next_float(), rng.next_float(), rng.next_float()};
	#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const float p = (((float)c[k] + r[k]) / (float)NERF_GRIDSIZE - 0.5f) * scalbnf(1.0f, (int)level) + 0.5f;
		positions[(size_t)i * 3 + k] = (p - bb.mn[k]) / (bb.mx[k] - bb.mn[k]);
	}
	indices[i] = idx;
}

// The UNIFORM samples of an occupancy refresh (thresh < 0: the first candidate cell always passes), one cascade: sample i visits cell
// idx(i) = ((i + step * n) * 56924617 + 96925573) mod 128^3, a pseudo-random walk through the Morton-ordered grid, so consecutive samples
// lie anywhere in the cube and the network pass over them gathers without any locality (0.63 ns per point against 0.2 ns on a lattice).
// idx is invertible (the multiplier is odd): here one thread takes one CELL t, finds the sample that visits it,
//   i = (56924617^-1 * (t - 96925573) - step * n) mod 128^3,       chosen iff i < n,
// and the chosen samples are written in cell order -- neighbours in space are neighbours in the buffer.  With n = 128^3 (the first 256
// steps of a frame) every cell is chosen and the slot is the cell; otherwise the slot comes from a prefix sum across blocks
// (lookback.cuh), which keeps the order deterministic.  Each sample is the one the reference generates (same i, same random draws, same
// position bits); only its place in the buffers differs, and the splat is an order-independent max.
#define GS_TILE 256
__global__ void __launch_bounds__(GS_TILE)
k_grid_samples_uniform_ordered(uint32_t n, Pcg32 rng, uint32_t step, Aabb bb, const float* __restrict__ grid_in, float thresh, LookbackScan scan,
                               float* __restrict__ positions, uint32_t* __restrict__ indices) {
	__shared__ uint32_t s_tile, s_warp[GS_TILE / 32];
	const bool all = n >= G3;
	uint32_t tile = blockIdx.x;
	if (!all) {
		if (threadIdx.x == 0) s_tile = lookback_take_ticket(scan);
		__syncthreads();
		tile = s_tile;
	}
	const uint32_t t = tile * GS_TILE + threadIdx.x;   // cell; 128^3 is a multiple of the tile
	const uint32_t i = (2759905401u * (t - 96925573u) - step * n) & (G3 - 1u);   // 2759905401 * 56924617 = 1 (mod 2^32)
	const bool chosen = i < n;
	uint32_t slot = t;
	if (!all) {
		const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
		const uint32_t votes = __ballot_sync(0xffffffffu, chosen);
		if (lane == 0) s_warp[warp] = __popc(votes);
		__syncthreads();
		if (warp == 0) {
			const uint32_t c = lane < GS_TILE / 32 ? s_warp[lane] : 0u;
			uint32_t inc = c;
			#pragma unroll
			for (int o = 1; o < GS_TILE / 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
			const uint32_t aggregate = __shfl_sync(0xffffffffu, inc, GS_TILE / 32 - 1);
			const uint32_t before = lookback_exclusive(scan, tile, aggregate, lane);
			if (lane < GS_TILE / 32) s_warp[lane] = before + inc - c;
		}
		__syncthreads();
		slot = s_warp[warp] + __popc(votes & ((1u << lane) - 1u));
	}
	if (!chosen) return;
	rng.advance((int64_t)(uint32_t)(i * 4u));
	(void)rng.next_float();   // the cascade draw: one cascade
	uint32_t idx = t;
	if (!(grid_in[t] > thresh)) {   // (a NaN density in the grid: the reference moves on to its next candidates)
		for (uint32_t j = 1; j < 10; ++j) {
			idx = ((i + step * n) * 56924617u + j * 19349663u + 96925573u) % G3;
			if (grid_in[idx] > thresh) break;
		}
	}
	const uint32_t c[3] = {morton3D_invert(idx >> 0), morton3D_invert(idx >> 1), morton3D_invert(idx >> 2)};
	const float r[3] = {rng.next_float(), rng.next_float(), rng.next_float()};
	#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const float p = (((float)c[k] + r[k]) / (float)NERF_GRIDSIZE - 0.5f) * scalbnf(1.0f, 0) + 0.5f;
		positions[(size_t)slot * 3 + k] = (p - bb.mn[k]) / (bb.mx[k] - bb.mn[k]);
	}
	indices[slot] = idx;
}

__global__ void k_grid_splat(uint32_t n, const uint32_t* __restrict__ indices, const __half* __restrict__ density, float* __restrict__ grid_tmp) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	// positive floats order like their bit patterns (testbed_nerf.cu:687-689)
	atomicMax(reinterpret_cast<uint32_t*>(&grid_tmp[indices[i]]), __float_as_uint(__half2float(density[i])));
}

__global__ void k_grid_decay(uint32_t n, float decay, float* __restrict__ grid, const float* __restrict__ grid_tmp) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float prev = grid[i];
	grid[i] = (prev < 0.f) ? prev : fmaxf(prev * decay, grid_tmp[i]);
}

// mean of max(val,0)/n over cascade 0: fp32 block reduction + one atomic per block (order differs from the reference's
// reduce_sum tree; both are fp32 sums of the same 2^21 terms)
__global__ void k_grid_mean(uint32_t n, const float* __restrict__ grid, float* __restrict__ mean) {
	__shared__ float ws[32];
	float s = 0.f;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += fmaxf(grid[i], 0.f) / (float)n;
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
	__syncthreads();
	if (threadIdx.x < 32) {
		s = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.f;
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
		if (threadIdx.x == 0) atomicAdd(mean, s);
	}
}

__global__ void k_grid_to_bitfield(uint32_t n_elements, uint32_t n_nonzero, const float* __restrict__ grid, uint8_t* __restrict__ bitfield, const float* __restrict__ mean_ptr) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	if (i >= n_nonzero) { bitfield[i] = 0; return; }
	const float thresh = fminf(0.1f, *mean_ptr);   // NERF_MIN_OPTICAL_THICKNESS (:75,:764)
	const float4 a = reinterpret_cast<const float4*>(grid)[i * 2], b = reinterpret_cast<const float4*>(grid)[i * 2 + 1];
	uint8_t bits = 0;
	bits |= a.x > thresh ? 1 : 0; bits |= a.y > thresh ? 2 : 0; bits |= a.z > thresh ? 4 : 0; bits |= a.w > thresh ? 8 : 0;
	bits |= b.x > thresh ? 16 : 0; bits |= b.y > thresh ? 32 : 0; bits |= b.z > thresh ? 64 : 0; bits |= b.w > thresh ? 128 : 0;
	bitfield[i] = bits;
}

__global__ void k_bitfield_max_pool(uint32_t n_elements, const uint8_t* __restrict__ prev_level, uint8_t* __restrict__ next_level) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	const uint2 v = reinterpret_cast<const uint2*>(prev_level)[i];
	uint8_t bits = 0;
	#pragma unroll
	for (int j = 0; j < 4; ++j) {
		bits |= ((v.x >> (8 * j)) & 0xff) ? (1 << j) : 0;
		bits |= ((v.y >> (8 * j)) & 0xff) ? (1 << (4 + j)) : 0;
	}
	const uint32_t x = morton3D_invert(i >> 0) + NERF_GRIDSIZE / 8;
	const uint32_t y = morton3D_invert(i >> 1) + NERF_GRIDSIZE / 8;
	const uint32_t z = morton3D_invert(i >> 2) + NERF_GRIDSIZE / 8;
	next_level[morton3D(x, y, z)] |= bits;
}

__global__ void k_reduce_sum(uint32_t n, const float* __restrict__ in, float* __restrict__ out) {
	__shared__ float ws[32];
	float s = 0.f;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += in[i];
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
	__syncthreads();
	if (threadIdx.x < 32) {
		s = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.f;
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
		if (threadIdx.x == 0) atomicAdd(out, s);
	}
}

int launch_density_grid_samples(cudaStream_t stream, uint32_t n, const uint64_t rng_state[2], uint32_t step, const Aabb& bb, const float* grid_in,
                                uint32_t n_cascades, float thresh, float* positions, uint32_t* indices,
                                unsigned long long* lookback_desc, uint32_t* lookback_ticket, uint32_t lookback_epoch) {
	if (n == 0) return 0;
	Pcg32 rng; rng.state = rng_state[0]; rng.inc = rng_state[1];
	if (lookback_desc && thresh < 0.f && n_cascades == 1 && n <= G3) {   // uniform samples (the grid holds no negative marks here), in cell order
		const LookbackScan scan = {lookback_desc, lookback_ticket, lookback_epoch};
		k_grid_samples_uniform_ordered<<<G3 / GS_TILE, GS_TILE, 0, stream>>>(n, rng, step, bb, grid_in, thresh, scan, positions, indices);
		COUNT_LAUNCH(1);
		CUDA_CHECK(cudaGetLastError());
		return 0;
	}
	k_grid_samples<<<(n + 127) / 128, 128, 0, stream>>>(n, rng, step, bb, grid_in, n_cascades, thresh, positions, indices);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}

int launch_density_grid_splat_decay(cudaStream_t stream, uint32_t n_samples, const uint32_t* indices, const __half* density, uint32_t n_elements,
                                    float decay, float* grid_tmp, float* grid) {
	CUDA_CHECK(cudaMemsetAsync(grid_tmp, 0, sizeof(float) * n_elements, stream));
	if (n_samples) k_grid_splat<<<(n_samples + 255) / 256, 256, 0, stream>>>(n_samples, indices, density, grid_tmp);
	COUNT_LAUNCH(1);
	k_grid_decay<<<(n_elements + 255) / 256, 256, 0, stream>>>(n_elements, decay, grid, grid_tmp);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}

int launch_density_grid_bitfield(cudaStream_t stream, const float* grid, uint32_t n_cascades_used, float* mean_tmp, uint8_t* bitfield) {
	CUDA_CHECK(cudaMemsetAsync(mean_tmp, 0, sizeof(float), stream));
	k_grid_mean<<<296, 256, 0, stream>>>(G3, grid, mean_tmp);
	COUNT_LAUNCH(1);
	const uint32_t n_bytes = G3 / 8 * NERF_CASCADES;
	k_grid_to_bitfield<<<(n_bytes + 255) / 256, 256, 0, stream>>>(n_bytes, G3 / 8 * n_cascades_used, grid, bitfield, mean_tmp);
	COUNT_LAUNCH(1);
	for (uint32_t level = 1; level < NERF_CASCADES; ++level) {
		k_bitfield_max_pool<<<(G3 / 64 + 255) / 256, 256, 0, stream>>>(G3 / 64, bitfield + (size_t)G3 / 8 * (level - 1), bitfield + (size_t)G3 / 8 * level);
		COUNT_LAUNCH(1);
	}
	CUDA_CHECK(cudaGetLastError());
	return 0;
}

int launch_reduce_sum(cudaStream_t stream, const float* in, uint32_t n, float* out) {
	CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(float), stream));
	if (n) k_reduce_sum<<<min(296u, (n + 255) / 256), 256, 0, stream>>>(n, in, out);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}
