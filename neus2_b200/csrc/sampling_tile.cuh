// sampling_tile.cuh -- device side of training-ray generation shared by sampling.cu (stand-alone kernels) and optimizer.cu (the step's
// tail kernel, which generates the NEXT step's samples beside this step's Adam update).  Compiled with -fmad=false in both.
#pragma once
#include "ngp_common.cuh"
#include "ngp_kernels.h"
#include "lookback.cuh"

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


// ---- one tile = GF_RAYS consecutive rays, one warp each: count, prefix sums across tiles, emit ----
// March (the window records stay in shared memory, the ray set-up in registers), then two prefix sums across the tiles of the launch
// (lookback.cuh), taken by two warps side by side: samples in front of the ray (its base offset) and rays with samples in front of it (its
// slot -- a ray is kept iff it has samples and they end inside the budget, and because the base offsets keep counting dropped rays, every ray
// behind a dropped one is dropped too: the kept rays are exactly the first rays with samples, so the slot does not depend on the budget
// test), then emit.  Same outputs as the count / scan / emit kernels, bit for bit: the order is ray-index order either way.
#define GF_RAYS 8
struct GenTileSmem { WindowRec windows[GF_RAYS][MAX_WINDOWS]; uint32_t n_windows[GF_RAYS], count[GF_RAYS], base[GF_RAYS], slot[GF_RAYS]; };

// All GF_RAYS * 32 threads of the block; `tile` must come from a ticket (lookback.cuh).  Ends with the block synchronised on `s` being free again.
__device__ __forceinline__ void gen_tile(const GenArgs& a, uint32_t tile, uint32_t n_tiles, GenTileSmem& s) {
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t i = tile * GF_RAYS + warp;
	RaySetup r;
	r.skip = true;
	uint32_t j = 0;
	if (i < a.n_rays) {
		r = setup_ray(i * a.ray_stride + a.ray_offset, a.n_rays_global, a.n_rays_total, a.rng, a.ds, a.bb, a.gmv, a.snap != 0, a.cone_angle);
		if (!r.skip) j = march_ray_warp(a.bb, r.ray_o, r.dir, r.startt, a.cone_angle, a.bitfield, NERF_STEPS, s.windows[warp], &s.n_windows[warp], lane);
		if (lane == 0) a.numsteps_all[i] = j;
	}
	if (lane == 0) s.count[warp] = j;
	__syncthreads();
	if (warp < 2) {   // warp 0: samples, warp 1: rays with samples
		const uint32_t c = lane < GF_RAYS ? s.count[lane] : 0u;
		const uint32_t v = warp == 0 ? c : (c > 0 ? 1u : 0u);
		uint32_t inc = v;
		#pragma unroll
		for (int o = 1; o < GF_RAYS; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
		const uint32_t aggregate = __shfl_sync(0xffffffffu, inc, GF_RAYS - 1);
		const uint32_t before = lookback_exclusive(warp == 0 ? a.scan_samples : a.scan_rays, tile, aggregate, lane);
		if (lane < GF_RAYS) (warp == 0 ? s.base : s.slot)[lane] = before + inc - v;
		if (warp == 0 && tile == n_tiles - 1 && lane == 0) a.counters[1] = before + aggregate;   // the reference's numsteps_counter (incl. dropped rays)
	}
	__syncthreads();
	const uint32_t b = s.base[warp], sl = s.slot[warp];
	const bool keep = j > 0 && b + j <= a.max_samples;
	if (threadIdx.x == 0) {   // kept rays / samples on the longest kept ray: one pair of atomics per tile
		uint32_t kept_end = 0, longest = 0;
		#pragma unroll
		for (int k = 0; k < GF_RAYS; ++k) {
			const uint32_t c = s.count[k];
			if (c > 0 && s.base[k] + c <= a.max_samples) { kept_end = s.slot[k] + 1u; longest = max(longest, c); }
		}
		if (kept_end) { atomicMax(a.counters + 0, kept_end); atomicMax(a.counters + 6, longest); }
	}
	if (keep) {   // ---- emit: replays the recorded windows (start on the t-chain + bitmask of the positions that became samples) ----
		{
			const float v = lane == 0 ? r.o[0] : lane == 1 ? r.o[1] : lane == 2 ? r.o[2] : lane == 3 ? r.d[0] : lane == 4 ? r.d[1] : r.d[2];
			if (lane < 6) a.rays[(size_t)sl * 6 + lane] = v;
		}
		if (lane == 0) {
			a.ray_indices[sl] = i * a.ray_stride + a.ray_offset;
			if (a.ray_done) a.ray_done[sl] = 0xffffffffu;   // first network pass: no segment has terminated this ray yet (forward_rays.cu)
			a.numsteps[sl * 2 + 0] = j;
			a.numsteps[sl * 2 + 1] = b;
		}
		const float wdir[3] = {(r.dir[0] + 1.0f) * 0.5f, (r.dir[1] + 1.0f) * 0.5f, (r.dir[2] + 1.0f) * 0.5f};
		const float diag[3] = {a.bb.mx[0] - a.bb.mn[0], a.bb.mx[1] - a.bb.mn[1], a.bb.mx[2] - a.bb.mn[2]};
		const uint32_t nw = s.n_windows[warp];
		float* out = a.coords + (size_t)b * 7;
		uint32_t jj = 0;
		for (uint32_t wi = 0; wi < nw; ++wi) {
			const WindowRec rec = s.windows[warp][wi];
			const float t_mine = chain_position(rec.t_start, a.cone_angle, lane);
			if ((rec.mask >> lane) & 1u) {
				const uint32_t my_j = jj + __popc(rec.mask & ((1u << lane) - 1u));
				float* c = out + (size_t)my_j * 7;
				#pragma unroll
				for (int k = 0; k < 3; ++k) c[k] = (fmaf(t_mine, r.dir[k], r.ray_o[k]) - a.bb.mn[k]) / diag[k];
				c[3] = warp_dt(calc_dt(t_mine, a.cone_angle));
				c[4] = wdir[0]; c[5] = wdir[1]; c[6] = wdir[2];
			}
			jj += __popc(rec.mask);
		}
	}
	__syncthreads();
}
