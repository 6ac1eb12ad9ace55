// forward_rays.cu -- the training step's FIRST network pass, ray by ray, stopping where the loss stops (compiled with
// -fmad=false like loss.cu, whose transmittance recurrence it must reproduce bit for bit).
//
// Reference: train_nerf_step runs NerfNetwork::inference_mixed_precision over ALL marched samples (src/testbed_nerf.cu:3854), and
// compute_loss_kernel_train_nerf_with_global_movement then walks each ray front to back and stops as soon as the transmittance
// falls below 1e-4 (:1567-1571): in steady state ~60 % of the evaluated samples lie behind that point and are never read.
//
// Here a work item is a SEGMENT of a ray (up to 4 tiles = 64 consecutive samples); a warp pulls items from a queue ordered
// segment-major (every ray's first segment, then every ray's second ...).  It evaluates the segment 16 samples at a time (the same
// forward_tile pipeline as k_network_forward), multiplies (1 - alpha) into a transmittance that starts at 1 for the segment, with the
// loss kernel's own arithmetic (eval_sample), and stops -- and records its segment index as the ray's terminating segment, which lets
// LATER segments of that ray be skipped -- once that product
// is below the loss threshold.  The segment-local product bounds the true transmittance from above, so every sample the loss reads
// has been evaluated (with identical bits); rows behind the cut keep stale values nobody looks at.  Segmenting bounds the serial
// chain of a long ray that never saturates to 4 tiles.
#include "network.cuh"
#include "ngp_common.cuh"
#include "ngp_kernels.h"
#include "kernels.h"

#define FR_WARPS 8
static constexpr int FR_SMEM = WeightSmem::FWD_END * 2 + FR_WARPS * WarpSlab::FWD_END;

#ifndef FR_SEGMENT
#define FR_SEGMENT 32u   // samples per work item (2 tiles; measured 16 / 32 / 48 / 64 / 128: see DESIGN.md)
#endif

__global__ void __launch_bounds__(FR_WARPS * 32, 2)
k_network_forward_rays(GridMeta gm, ParamLayout pl, const __half* __restrict__ params, float sdf_bias, float cos_anneal_ratio,
                       const uint32_t* __restrict__ n_kept_ptr, const uint32_t* __restrict__ longest_ptr, const uint32_t* __restrict__ numsteps,
                       const float* __restrict__ coords, uint32_t max_samples, __half* __restrict__ out, float4* __restrict__ composite_in,
                       uint32_t* __restrict__ ray_done, uint32_t* __restrict__ work_counter /* [0] next item, [1] samples evaluated */) {
	extern __shared__ __align__(16) unsigned char smem[];
	__half* sw = reinterpret_cast<__half*>(smem);
	stage_weights(sw, params, pl, false, threadIdx.x, blockDim.x);
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	unsigned char* slab = smem + WeightSmem::FWD_END * 2 + warp * WarpSlab::FWD_END;
	float* sCoord = reinterpret_cast<float*>(slab + WarpSlab::COORD);
	const __half* sOUT = reinterpret_cast<const __half*>(slab + WarpSlab::OUT);
	const uint32_t n_kept = *n_kept_ptr;
	const uint32_t n_segments = (min(*longest_ptr, (uint32_t)NERF_STEPS) + FR_SEGMENT - 1) / FR_SEGMENT;
	const uint64_t n_items = (uint64_t)n_kept * n_segments;
	const float EPSILON = 1e-4f;      // testbed_nerf.cu:1567
	uint32_t evaluated = 0;
	FwdState st;
	while (true) {
		uint32_t item = 0;
		if (lane == 0) item = atomicAdd(work_counter, 1u);
		item = __shfl_sync(0xffffffffu, item, 0);
		if (item >= n_items) break;
		const uint32_t seg = item / n_kept, ray = item - seg * n_kept;
		const uint32_t ns = numsteps[ray * 2 + 0], base = numsteps[ray * 2 + 1];
		if (base >= max_samples) continue;
		const uint32_t n = min(ns, max_samples - base);
		const uint32_t s_begin = seg * FR_SEGMENT, s_end = min(n, s_begin + FR_SEGMENT);
		if (s_begin >= n) continue;
		if (seg > 0 && ray_done[ray] < seg) continue;    // an EARLIER segment already pushed the transmittance below the threshold
		float T = 1.f;
		float dir[3] = {0.f, 0.f, 1.f};
		float cr[4];   // the segment's next tile of coordinates, in flight while the current tile is processed
		auto fetch = [&](uint32_t s0) {
			const uint32_t nv = min((uint32_t)TILE, s_end - s0);
			const float* src = coords + (size_t)(base + s0) * 7;
			#pragma unroll
			for (int k = 0; k < 4; ++k) { const int i = lane + 32 * k; cr[k] = (i < 16 * 7 && (uint32_t)(i / 7) < nv) ? src[i] : 0.f; }
		};
		fetch(s_begin);
		for (uint32_t s0 = s_begin; s0 < s_end; s0 += TILE) {
			const uint32_t n_valid = min((uint32_t)TILE, s_end - s0);
			#pragma unroll
			for (int k = 0; k < 4; ++k) { const int i = lane + 32 * k; if (i < 16 * 7) sCoord[i] = cr[k]; }
			__syncwarp();
			if (s0 + TILE < s_end) fetch(s0 + TILE);
			forward_tile<false>(gm, params, pl, sw, slab, sdf_bias, lane, st);
			{   // 16 rows x 32 B, coalesced 16-byte stores
				const uint32_t r = lane >> 1;
				if (r < n_valid) reinterpret_cast<uint4*>(out)[(size_t)(base + s0 + r) * 2 + (lane & 1)] = reinterpret_cast<const uint4*>(sOUT)[lane];
			}
			evaluated += n_valid;
			if (s0 == s_begin) {   // BENT_DIR (:1583-1588): the loss takes the view direction from sample 0's output; every sample of a ray carries the same one
				const float dv[3] = {__half2float(sOUT[8]) * 2.0f - 1.0f, __half2float(sOUT[9]) * 2.0f - 1.0f, __half2float(sOUT[10]) * 2.0f - 1.0f};
				normalize3(dv, dir);
			}
			float keep = 1.f;   // 1 - alpha of this lane's sample
			if (lane < (int)n_valid) {
				const SampleEval e = eval_sample(sOUT + lane * 16, sCoord + lane * 7, dir, cos_anneal_ratio);
				keep = 1.f - e.alpha;
				// what the loss's compositing pass consumes per sample -- the very bits it would compute itself from the row just stored (same
				// function, same inputs, this file is compiled without FMA contraction like loss.cu): it then only replays its recurrences
				if (composite_in) composite_in[(size_t)base + s0 + lane] = make_float4(e.alpha, e.rgb[0], e.rgb[1], e.rgb[2]);
			}
			// T *= product of the tile's (1 - alpha) (:1626), as a tree: this segment-local T only has to BOUND the loss's own (serially
			// multiplied) transmittance from above, which the factor 0.5 below guarantees against any difference in rounding
			#pragma unroll
			for (int o = 8; o > 0; o >>= 1) keep *= __shfl_xor_sync(0xffffffffu, keep, o);
			T *= __shfl_sync(0xffffffffu, keep, 0);
			__syncwarp();
			// the loss breaks BEFORE a sample once its transmittance (<= this segment-local product) is below 1e-4, and T never grows:
			// nothing behind this tile is read.
			if (T < EPSILON * 0.5f) { if (lane == 0) atomicMin(ray_done + ray, seg); break; }   // lowest terminating segment: only later ones may be skipped
		}
	}
	if (lane == 0 && evaluated) atomicAdd(work_counter + 1, evaluated);
}

int launch_network_forward_rays(cudaStream_t stream, const GridMeta& gm, const ParamLayout& pl, const __half* params, float sdf_bias, float cos_anneal_ratio,
                                const uint32_t* n_kept_ptr, const uint32_t* longest_ptr, const uint32_t* numsteps, const float* coords, uint32_t max_samples,
                                uint32_t n_rays_max, __half* out, float4* composite_in, uint32_t* ray_done, uint32_t* work_counter) {
	static bool attr[NEUS_MAX_DEVICES] = {};
	if (ensure_dynamic_smem(k_network_forward_rays, FR_SMEM, attr)) return -1;
	// Preconditions (no launches of their own): work_counter[0..1] are zero -- they are two of the step's counters, cleared with them -- and
	// ray_done[slot] = 0xffffffff ("no segment has terminated the ray") for every kept ray, written by sample generation (k_rays_emit).
	(void)n_rays_max;
	k_network_forward_rays<<<device_sm_count() * 2, FR_WARPS * 32, FR_SMEM, stream>>>(gm, pl, params, sdf_bias, cos_anneal_ratio, n_kept_ptr, longest_ptr, numsteps,
	                                                                                 coords, max_samples, out, composite_in, ray_done, work_counter);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}
