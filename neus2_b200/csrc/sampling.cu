// sampling.cu -- training-ray generation and occupancy-grid marching (compiled with -fmad=false).
// Restates generate_training_samples_nerf_with_global_movement (src/testbed_nerf.cu:1263-1456), made
// deterministic: instead of ordering rays by atomicAdd arrival (:1421,:1428) the sample base offsets and ray
// slots come from exclusive scans in ray-index order.  Per-ray numsteps are unchanged.
// Data-parallel sharding: a rank traces the global rays ray_offset + i * ray_stride, i < n_rays (rank + i * world_size: every rank sees every
// image in the same proportion, so the ranks' sample counts -- and with them the time to the next cross-rank exchange -- stay balanced),
// of a batch of n_rays_global;
// image / pixel / jitter are pure functions of the GLOBAL ray index (:1241-1261,:1304).
#include "ngp_common.cuh"
#include "ngp_kernels.h"
#include "sampling_tile.cuh"

__global__ void __launch_bounds__(128)
k_rays_count(uint32_t n_rays, uint32_t ray_offset, uint32_t ray_stride, uint32_t n_rays_global, uint32_t n_rays_total, Pcg32 rng, DatasetDev ds, Aabb bb, GlobalMove gmv,
             const uint8_t* __restrict__ bitfield, int snap, float cone_angle, uint32_t* __restrict__ numsteps_all,
             WindowRec* __restrict__ windows, uint32_t* __restrict__ n_windows, float* __restrict__ ray_setup) {
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (i >= n_rays) return;
	RaySetup r = setup_ray(i * ray_stride + ray_offset, n_rays_global, n_rays_total, rng, ds, bb, gmv, snap != 0, cone_angle);
	uint32_t j = 0;
	if (!r.skip) j = march_ray_warp(bb, r.ray_o, r.dir, r.startt, cone_angle, bitfield, NERF_STEPS, windows + (size_t)i * MAX_WINDOWS, n_windows + i, lane);
	if (lane == 0) {
		numsteps_all[i] = j;
		if (j) {   // what the emit pass needs: unnormalised ray (o, d) + moved origin / direction
			float* rs = ray_setup + (size_t)i * 12;
			#pragma unroll
			for (int k = 0; k < 3; ++k) { rs[k] = r.o[k]; rs[3 + k] = r.d[k]; rs[6 + k] = r.ray_o[k]; rs[9 + k] = r.dir[k]; }
		}
	}
}

// Single-CTA two-stage exclusive scan over <= 2^20 elements (n_rays <= 2^18 in practice):
//   base[i]  = exclusive sum of counts;      counters[1] = total (the reference's numsteps_counter incl. dropped rays)
//   keep[i]  = count>0 && base+count <= max_samples;  slot[i] = exclusive sum of keep;  counters[0] = number kept
#define SCAN_THREADS 1024
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums, uint32_t* total) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc = v;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
	if (lane == 31) warp_sums[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		uint32_t w = warp_sums[lane];
		uint32_t winc = w;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += y; }
		warp_sums[lane] = winc - w;
		if (lane == 31) *total = winc;
	}
	__syncthreads();
	const uint32_t r = warp_sums[warp] + inc - v;
	__syncthreads();
	return r;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_rays(uint32_t n, const uint32_t* __restrict__ counts, uint32_t max_samples, uint32_t* __restrict__ base, uint32_t* __restrict__ slot,
            uint32_t* __restrict__ counters) {
	__shared__ uint32_t warp_sums[32];
	__shared__ uint32_t total;
	const uint32_t per = (n + SCAN_THREADS - 1) / SCAN_THREADS;
	const uint32_t b = min(n, threadIdx.x * per), e = min(n, b + per);
	uint32_t s = 0;
	for (uint32_t i = b; i < e; ++i) s += counts[i];
	uint32_t off = block_exclusive_scan(s, warp_sums, &total);
	if (threadIdx.x == 0) counters[1] = total;
	uint32_t kept = 0, longest = 0;
	for (uint32_t i = b; i < e; ++i) {
		const uint32_t c = counts[i];
		base[i] = off;
		const bool k = c > 0 && off + c <= max_samples;
		kept += k ? 1u : 0u;
		if (k) longest = max(longest, c);
		off += c;
	}
	if (longest) atomicMax(counters + 6, longest);   // samples on the longest kept ray (sizes the first network pass's work list)
	uint32_t soff = block_exclusive_scan(kept, warp_sums, &total);
	if (threadIdx.x == 0) counters[0] = total;
	for (uint32_t i = b; i < e; ++i) {
		const uint32_t c = counts[i];
		const bool k = c > 0 && base[i] + c <= max_samples;
		slot[i] = k ? soff : 0xffffffffu;
		soff += k ? 1u : 0u;
	}
}

// Emit pass: replays the recorded windows (start of the window on the t-chain + bitmask of the positions that became
// samples); no occupancy lookups, no control flow.  Positions are recomputed with the same arithmetic as the count pass.
__global__ void __launch_bounds__(128)
k_rays_emit(uint32_t n_rays, uint32_t ray_offset, uint32_t ray_stride, Aabb bb, float cone_angle, const uint32_t* __restrict__ numsteps_all,
            const uint32_t* __restrict__ base, const uint32_t* __restrict__ slot, const WindowRec* __restrict__ windows,
            const uint32_t* __restrict__ n_windows, const float* __restrict__ ray_setup,
            uint32_t* __restrict__ ray_indices, float* __restrict__ rays, uint32_t* __restrict__ numsteps, float* __restrict__ coords,
            uint32_t* __restrict__ ray_done) {
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (i >= n_rays) return;
	const uint32_t sl = slot[i];
	if (sl == 0xffffffffu) return;
	const float* rs = ray_setup + (size_t)i * 12;
	const uint32_t j_total = numsteps_all[i], b = base[i];
	if (lane < 6) rays[(size_t)sl * 6 + lane] = rs[lane];
	if (lane == 0) {
		ray_indices[sl] = i * ray_stride + ray_offset;
		if (ray_done) ray_done[sl] = 0xffffffffu;   // first network pass: no segment has terminated this ray yet (forward_rays.cu)
		numsteps[sl * 2 + 0] = j_total;
		numsteps[sl * 2 + 1] = b;
	}
	const float ray_o[3] = {rs[6], rs[7], rs[8]}, dir[3] = {rs[9], rs[10], rs[11]};
	const float wdir[3] = {(dir[0] + 1.0f) * 0.5f, (dir[1] + 1.0f) * 0.5f, (dir[2] + 1.0f) * 0.5f};
	const float diag[3] = {bb.mx[0] - bb.mn[0], bb.mx[1] - bb.mn[1], bb.mx[2] - bb.mn[2]};
	const WindowRec* w = windows + (size_t)i * MAX_WINDOWS;
	const uint32_t nw = n_windows[i];
	float* out = coords + (size_t)b * 7;
	uint32_t j = 0;
	for (uint32_t wi = 0; wi < nw; ++wi) {
		const WindowRec rec = w[wi];
		const float t_mine = chain_position(rec.t_start, cone_angle, lane);
		if ((rec.mask >> lane) & 1u) {
			const uint32_t my_j = j + __popc(rec.mask & ((1u << lane) - 1u));
			float* c = out + (size_t)my_j * 7;
			#pragma unroll
			for (int k = 0; k < 3; ++k) c[k] = (fmaf(t_mine, dir[k], ray_o[k]) - bb.mn[k]) / diag[k];
			c[3] = warp_dt(calc_dt(t_mine, cone_angle));
			c[4] = wdir[0]; c[5] = wdir[1]; c[6] = wdir[2];
		}
		j += __popc(rec.mask);
	}
}

// Count, scans and emit in ONE kernel (gen_tile, sampling_tile.cuh): one block per tile of GF_RAYS rays.
__global__ void __launch_bounds__(GF_RAYS * 32)
k_rays_fused(GenArgs a) {
	__shared__ GenTileSmem s;
	__shared__ uint32_t s_tile;
	if (threadIdx.x == 0) { s_tile = lookback_take_ticket(a.scan_samples); (void)lookback_take_ticket(a.scan_rays); }   // the second ticket only keeps its counter self-resetting
	__syncthreads();
	gen_tile(a, s_tile, gridDim.x, s);
}

size_t sampling_window_bytes_per_ray() { return sizeof(WindowRec) * MAX_WINDOWS; }

GenArgs make_gen_args(uint32_t n_rays, uint32_t ray_offset, uint32_t ray_stride, uint32_t n_rays_global, uint32_t n_rays_total, const uint64_t rng_state[2], const DatasetDev& ds,
                      const Aabb& bb, const GlobalMove& gmv, const uint8_t* bitfield, int snap, float cone_angle, uint32_t max_samples, uint32_t* numsteps_all,
                      uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords, uint32_t* counters, uint32_t* ray_done,
                      unsigned long long* lookback_desc, uint32_t* lookback_tickets, uint32_t lookback_epoch) {
	GenArgs a;
	a.n_rays = n_rays; a.ray_offset = ray_offset; a.ray_stride = ray_stride; a.n_rays_global = n_rays_global; a.n_rays_total = n_rays_total;
	a.rng.state = rng_state[0]; a.rng.inc = rng_state[1];
	a.ds = ds; a.bb = bb; a.gmv = gmv; a.bitfield = bitfield; a.snap = snap; a.cone_angle = cone_angle; a.max_samples = max_samples;
	const size_t tiles = (n_rays + GF_RAYS - 1) / GF_RAYS;
	a.scan_samples = LookbackScan{lookback_desc, lookback_tickets, lookback_epoch};
	a.scan_rays = LookbackScan{lookback_desc + tiles, lookback_tickets + 1, lookback_epoch};
	a.numsteps_all = numsteps_all; a.ray_indices = ray_indices; a.rays = rays; a.numsteps = numsteps; a.coords = coords; a.ray_done = ray_done; a.counters = counters;
	return a;
}

int launch_generate_training_samples(cudaStream_t stream, uint32_t n_rays, uint32_t ray_offset, uint32_t ray_stride, uint32_t n_rays_global, uint32_t n_rays_total, const uint64_t rng_state[2], const DatasetDev& ds,
                                     const Aabb& bb, const GlobalMove& gmv, const uint8_t* bitfield, int snap, float cone_angle, uint32_t max_samples,
                                     uint32_t* numsteps_all, uint32_t* base, uint32_t* slot, uint32_t* ray_indices, float* rays, uint32_t* numsteps,
                                     float* coords, uint32_t* counters, void* windows, uint32_t* n_windows, float* ray_setup, uint32_t* ray_done,
                                     unsigned long long* lookback_desc, uint32_t* lookback_tickets, uint32_t lookback_epoch) {
	if (n_rays == 0) return 0;
	Pcg32 rng; rng.state = rng_state[0]; rng.inc = rng_state[1];
	if (lookback_desc) {
		const GenArgs a = make_gen_args(n_rays, ray_offset, ray_stride, n_rays_global, n_rays_total, rng_state, ds, bb, gmv, bitfield, snap, cone_angle, max_samples, numsteps_all,
		                                ray_indices, rays, numsteps, coords, counters, ray_done, lookback_desc, lookback_tickets, lookback_epoch);
		k_rays_fused<<<(n_rays + GF_RAYS - 1) / GF_RAYS, GF_RAYS * 32, 0, stream>>>(a);
		COUNT_LAUNCH(1);
		CUDA_CHECK(cudaGetLastError());
		return 0;
	}
	const uint32_t blocks = (n_rays + 3) / 4;   // one warp per ray
	k_rays_count<<<blocks, 128, 0, stream>>>(n_rays, ray_offset, ray_stride, n_rays_global, n_rays_total, rng, ds, bb, gmv, bitfield, snap, cone_angle, numsteps_all,
	                                        (WindowRec*)windows, n_windows, ray_setup);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	k_scan_rays<<<1, SCAN_THREADS, 0, stream>>>(n_rays, numsteps_all, max_samples, base, slot, counters);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	k_rays_emit<<<blocks, 128, 0, stream>>>(n_rays, ray_offset, ray_stride, bb, cone_angle, numsteps_all, base, slot, (const WindowRec*)windows, n_windows, ray_setup,
	                                       ray_indices, rays, numsteps, coords, ray_done);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}
