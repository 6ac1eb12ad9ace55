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

struct RaySetup { float o[3], d[3], ray_o[3], dir[3], startt; bool skip; };

__device__ __forceinline__ RaySetup setup_ray(uint32_t i, uint32_t n_rays, uint32_t n_rays_total, Pcg32 rng, const DatasetDev& ds,
                                               const Aabb& bb, const GlobalMove& gmv, bool snap, float cone_angle) {
	RaySetup r;
	r.skip = false;
	const uint32_t img = image_idx(i, n_rays, n_rays_total, ds.n_images);
	rng.advance((int64_t)(uint32_t)(i * N_MAX_RANDOM_SAMPLES_PER_RAY));
	float xy[2];
	random_image_pos_training(rng, ds.width, ds.height, snap, xy);
	int px, py; image_pos(xy, ds.width, ds.height, &px, &py);
	float rgba[4]; read_rgba(ds, img, px, py, rgba);
	if (rgba[0] <= 0.0f && rng.next_float() >= 0.9) { r.skip = true; return r; }   // :1310
	(void)rng.next_float();                                                         // motionblur_time :1317
	make_ray(ds, img, xy, r.o, r.d);
	#pragma unroll
	for (int k = 0; k < 3; ++k) r.ray_o[k] = r.o[k];
	normalize3(r.d, r.dir);
	#pragma unroll
	for (int k = 0; k < 3; ++k) r.ray_o[k] -= gmv.first_frame_offset[k];
	global_move(gmv, r.ray_o, r.dir);
	float tmin, tmax; ray_intersect(bb, r.ray_o, r.dir, &tmin, &tmax);
	tmin = fmaxf(tmin, 0.0f);                                                       // :1395
	r.startt = fmaf(calc_dt(tmin, cone_angle), rng.next_float(), tmin);               // :1397-1398 (nvcc contracts `t += dt * u` into one FFMA)
	return r;
}

// Warp-per-ray marching.  Every position the reference's loop can visit lies on ONE arithmetic chain
// t_0 = startt, t_{k+1} = fl(t_k + calc_dt(t_k)) -- both the occupied branch (t += dt, :1411) and the empty-space
// skip (do t += dt while t < t_target, :374-376) only ever add dt.  A warp therefore evaluates a window of 32
// consecutive chain positions in parallel (position, mip, occupancy bit, exit distance of the voxel) and then
// replays the reference's control flow over the window with ballots.  The visited set, hence every per-ray sample
// count and every emitted coordinate, is bit-identical to the serial loop; the dependent global load per step
// (the reason the serial version is latency-bound) is gone.
// Chain position `lane` of the window that starts at t_start: t_start (+) dt (+) dt ... with `lane` fp32 additions (round to nearest even),
// exactly the values the reference's serial loop visits.
// Constant step (aabb_scale 1, testbed_nerf.cu:3090: calc_dt(t, 0) is the constant minimum step), fast path: while the chain stays inside
// one binade [2^e, 2^(e+1)), adding the same dt moves the bit pattern by the same number of ulps every time -- the exact sum is
// (m + k) ulp + r with r = dt mod ulp, and which way it rounds depends on r alone; in the tie case r == ulp / 2 it depends on the parity of
// m + k, but the first addition leaves an even pattern and from then on the tie resolves the same way every time too.  So two real additions
// give the stride, the rest is one integer multiply-add on the bit pattern (instead of up to 31 dependent additions per lane and window).
// Anywhere else (near t = 0, across a power of two, cone-angle steps) the additions are done one by one.
__device__ __forceinline__ float chain_position(float t_start, float cone_angle, int lane) {
	float t = t_start;
	if (cone_angle == 0.0f) {
		const float dt0 = calc_dt(t_start, 0.0f);
		const float t_far = t_start + 33.0f * dt0;   // beyond chain position 32 with all its roundings (dt0 > 16 ulp below t = 256)
		if (t_start > 0.0f && t_start < 256.0f && (__float_as_uint(t_far) >> 23) == (__float_as_uint(t_start) >> 23)) {
			const float t1 = t_start + dt0, t2 = t1 + dt0;
			const uint32_t stride = __float_as_uint(t2) - __float_as_uint(t1);
			if (lane > 0) t = __uint_as_float(__float_as_uint(t1) + (uint32_t)(lane - 1) * stride);
		} else {
			#pragma unroll
			for (int k = 0; k < 31; ++k) if (k < lane) t += dt0;
		}
	} else {
		#pragma unroll
		for (int k = 0; k < 31; ++k) if (k < lane) t += calc_dt(t, cone_angle);
	}
	return t;
}

#define MAX_WINDOWS 72   // a ray crosses the unit cube in <= 1024 chain positions; every window but the last consumes >= 32
struct WindowRec { float t_start; uint32_t mask; };

__device__ __forceinline__ uint32_t march_ray_warp(const Aabb& bb, const float ray_o[3], const float dir[3], float startt, float cone_angle,
                                                   const uint8_t* __restrict__ bitfield, uint32_t limit, WindowRec* __restrict__ windows,
                                                   uint32_t* __restrict__ n_windows_out, int lane) {
	const float idir[3] = {1.0f / dir[0], 1.0f / dir[1], 1.0f / dir[2]};
	uint32_t j = 0;
	uint32_t n_win = 0;
	float t_start = startt;
	bool done = false;
	while (!done) {
		const float t_mine = chain_position(t_start, cone_angle, lane);   // this lane's chain position of the window
		float tl = __shfl_sync(0xffffffffu, t_mine, 31);
		tl += calc_dt(tl, cone_angle);   // chain position 32 = start of the next window
		float pos[3];
		#pragma unroll
		for (int k = 0; k < 3; ++k) pos[k] = fmaf(t_mine, dir[k], ray_o[k]);
		const bool inside = aabb_contains(bb, pos);
		const float dt = calc_dt(t_mine, cone_angle);
		bool occ = false;
		float t_target = 0.f;
		if (inside) {
			const uint32_t mip = (uint32_t)mip_from_dt(dt, pos);
			occ = density_grid_occupied_at(pos, bitfield, mip);
			if (!occ) t_target = t_mine + distance_to_next_voxel(pos, dir, idir, NERF_GRIDSIZE >> mip);
		}
		const uint32_t inside_mask = __ballot_sync(0xffffffffu, inside);
		const uint32_t occ_mask = __ballot_sync(0xffffffffu, occ);
		int my_j = -1;
		int cur = 0;
		float t_next = tl;   // start of the next window unless a skip overshoots it
		while (cur < 32) {
			if (!((inside_mask >> cur) & 1u) || j >= limit) { done = true; break; }
			if ((occ_mask >> cur) & 1u) {
				// a run of consecutive occupied positions inside the box becomes samples j, j+1, ... in one go (the serial loop
				// would visit them one by one with the same outcome); the run ends at the first empty / outside position or at `limit`
				const uint32_t run_bits = (occ_mask & inside_mask) >> cur;          // bit 0 is set
				uint32_t run = (uint32_t)__ffs((int)~run_bits) - 1u;               // ~run_bits has a set bit at or below position 32 - cur
				if (run_bits == 0xffffffffu) run = 32u;
				run = min(run, limit - j);
				if (lane >= cur && lane < cur + (int)run) my_j = (int)j + (lane - cur);
				j += run;
				cur += (int)run;
			} else {
				const float tt = __shfl_sync(0xffffffffu, t_target, cur);
				const uint32_t reach = __ballot_sync(0xffffffffu, lane > cur && t_mine >= tt);
				if (reach) {
					cur = __ffs(reach) - 1;
				} else {   // lands beyond the window: keep walking the chain
					while (t_next < tt) t_next += calc_dt(t_next, cone_angle);
					cur = 32;
				}
			}
		}
		// remember which chain positions of this window became samples: the emit pass replays (t_start, mask) only
		const uint32_t emit_mask = __ballot_sync(0xffffffffu, my_j >= 0);
		if (emit_mask) {
			if (lane == 0 && n_win < MAX_WINDOWS) { windows[n_win].t_start = t_start; windows[n_win].mask = emit_mask; }
			++n_win;
		}
		t_start = t_next;
	}
	if (lane == 0) *n_windows_out = min(n_win, (uint32_t)MAX_WINDOWS);
	return n_win > MAX_WINDOWS ? 0u : j;   // (cannot happen for rays through the unit cube; such a ray would be dropped)
}


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

size_t sampling_window_bytes_per_ray() { return sizeof(WindowRec) * MAX_WINDOWS; }

int launch_generate_training_samples(cudaStream_t stream, uint32_t n_rays, uint32_t ray_offset, uint32_t ray_stride, uint32_t n_rays_global, uint32_t n_rays_total, const uint64_t rng_state[2], const DatasetDev& ds,
                                     const Aabb& bb, const GlobalMove& gmv, const uint8_t* bitfield, int snap, float cone_angle, uint32_t max_samples,
                                     uint32_t* numsteps_all, uint32_t* base, uint32_t* slot, uint32_t* ray_indices, float* rays, uint32_t* numsteps,
                                     float* coords, uint32_t* counters, void* windows, uint32_t* n_windows, float* ray_setup, uint32_t* ray_done) {
	if (n_rays == 0) return 0;
	Pcg32 rng; rng.state = rng_state[0]; rng.inc = rng_state[1];
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
