// forward_rays.cu -- the training step's FIRST network pass, ray by ray, stopping where the loss stops (compiled with
// -fmad=false like loss.cu, whose transmittance recurrence it must reproduce bit for bit).
//
// Reference: train_nerf_step runs NerfNetwork::inference_mixed_precision over ALL marched samples (src/testbed_nerf.cu:3854), and
// compute_loss_kernel_train_nerf_with_global_movement then walks each ray front to back and stops as soon as the transmittance
// falls below 1e-4 (:1567-1571): in steady state ~60 % of the evaluated samples lie behind that point and are never read.
//
// Here one warp owns a RAY (dynamic queue): it evaluates the ray's samples 16 at a time (the same forward_tile pipeline as
// k_network_forward), multiplies the tile's (1 - alpha) into the running transmittance in sample order with the loss kernel's own
// arithmetic (eval_sample), and drops the rest of the ray once T is below the loss threshold.  Every sample the loss reads has
// been evaluated, with identical bits; rows behind the cut keep stale values nobody looks at.
#include "network.cuh"
#include "ngp_common.cuh"
#include "ngp_kernels.h"
#include "kernels.h"

#define FR_WARPS 8
static constexpr int FR_SMEM = WeightSmem::FWD_END * 2 + FR_WARPS * WarpSlab::FWD_END;

__global__ void __launch_bounds__(FR_WARPS * 32, 2)
k_network_forward_rays(GridMeta gm, ParamLayout pl, const __half* __restrict__ params, float sdf_bias, float cos_anneal_ratio,
                       const uint32_t* __restrict__ n_kept_ptr, const uint32_t* __restrict__ numsteps, const float* __restrict__ coords,
                       uint32_t max_samples, __half* __restrict__ out, uint32_t* __restrict__ work_counter /* [0] next ray, [1] samples evaluated */) {
	extern __shared__ __align__(16) unsigned char smem[];
	__half* sw = reinterpret_cast<__half*>(smem);
	stage_weights(sw, params, pl, false, threadIdx.x, blockDim.x);
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	unsigned char* slab = smem + WeightSmem::FWD_END * 2 + warp * WarpSlab::FWD_END;
	float* sCoord = reinterpret_cast<float*>(slab + WarpSlab::COORD);
	const __half* sOUT = reinterpret_cast<const __half*>(slab + WarpSlab::OUT);
	const uint32_t n_kept = *n_kept_ptr;
	const float EPSILON = 1e-4f;      // testbed_nerf.cu:1567
	uint32_t evaluated = 0;
	FwdState st;
	// one tile of coordinates is always in flight: the ray's next tile or, on its last tile, the first tile of the warp's NEXT ray
	// (grabbed one ray ahead), so the HBM latency of the 448-byte tile hides behind a whole tile of work
	float cr[4];
	uint32_t pref_row = 0xffffffffu;
	auto prefetch = [&](uint32_t row, uint32_t n_valid) {
		const float* src = coords + (size_t)row * 7;
		#pragma unroll
		for (int k = 0; k < 4; ++k) { const int i = lane + 32 * k; cr[k] = (i < 16 * 7 && (uint32_t)(i / 7) < n_valid) ? src[i] : 0.f; }
		pref_row = row;
	};
	auto grab = [&]() { uint32_t r = 0; if (lane == 0) r = atomicAdd(work_counter, 1u); return __shfl_sync(0xffffffffu, r, 0); };
	auto ray_extent = [&](uint32_t r, uint32_t& base, uint32_t& n) {
		base = 0; n = 0;
		if (r < n_kept) { const uint32_t ns = numsteps[r * 2 + 0]; base = numsteps[r * 2 + 1]; n = base < max_samples ? min(ns, max_samples - base) : 0u; }
	};
	uint32_t next_ray = grab(), next_base, next_n;
	ray_extent(next_ray, next_base, next_n);
	while (next_ray < n_kept) {
		const uint32_t base = next_base, n = next_n;
		next_ray = grab();
		ray_extent(next_ray, next_base, next_n);
		float T = 1.f;
		float dir[3] = {0.f, 0.f, 1.f};
		for (uint32_t s0 = 0; s0 < n; s0 += TILE) {
			const uint32_t n_valid = min((uint32_t)TILE, n - s0);
			if (pref_row != base + s0) prefetch(base + s0, n_valid);     // only after an early cut (the tile in flight belonged to the old ray)
			#pragma unroll
			for (int k = 0; k < 4; ++k) { const int i = lane + 32 * k; if (i < 16 * 7) sCoord[i] = cr[k]; }
			__syncwarp();
			if (s0 + TILE < n) prefetch(base + s0 + TILE, min((uint32_t)TILE, n - s0 - TILE));
			else if (next_n) prefetch(next_base, min((uint32_t)TILE, next_n));
			forward_tile<false>(gm, params, pl, sw, slab, sdf_bias, lane, st);
			{   // 16 rows x 32 B, coalesced 16-byte stores
				const uint32_t r = lane >> 1;
				if (r < n_valid) reinterpret_cast<uint4*>(out)[(size_t)(base + s0 + r) * 2 + (lane & 1)] = reinterpret_cast<const uint4*>(sOUT)[lane];
			}
			evaluated += n_valid;
			if (s0 == 0) {   // BENT_DIR (:1583-1588): the view direction the loss uses comes from sample 0's output
				const float dv[3] = {__half2float(sOUT[8]) * 2.0f - 1.0f, __half2float(sOUT[9]) * 2.0f - 1.0f, __half2float(sOUT[10]) * 2.0f - 1.0f};
				normalize3(dv, dir);
			}
			float alpha = 0.f;
			if (lane < (int)n_valid) alpha = eval_sample(sOUT + lane * 16, sCoord + lane * 7, dir, cos_anneal_ratio).alpha;
			for (uint32_t k = 0; k < n_valid; ++k) T *= (1.f - __shfl_sync(0xffffffffu, alpha, (int)k));   // T *= (1 - alpha), sample order (:1626)
			__syncwarp();
			// the loss breaks BEFORE a sample once T < 1e-4; T never grows, so nothing behind this tile is read.  The factor keeps
			// a tile of slack against any last-ulp difference between the two evaluations (there is none when both are compiled alike).
			if (T < EPSILON * 0.5f) break;
		}
	}
	if (lane == 0 && evaluated) atomicAdd(work_counter + 1, evaluated);
}

int launch_network_forward_rays(cudaStream_t stream, const GridMeta& gm, const ParamLayout& pl, const __half* params, float sdf_bias, float cos_anneal_ratio,
                                const uint32_t* n_kept_ptr, const uint32_t* numsteps, const float* coords, uint32_t max_samples, __half* out, uint32_t* work_counter) {
	static bool attr[NEUS_MAX_DEVICES] = {};
	if (ensure_dynamic_smem(k_network_forward_rays, FR_SMEM, attr)) return -1;
	CUDA_CHECK(cudaMemsetAsync(work_counter, 0, 2 * sizeof(uint32_t), stream));
	k_network_forward_rays<<<device_sm_count() * 2, FR_WARPS * 32, FR_SMEM, stream>>>(gm, pl, params, sdf_bias, cos_anneal_ratio, n_kept_ptr, numsteps, coords,
	                                                                                 max_samples, out, work_counter);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}
