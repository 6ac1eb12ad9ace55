// render.cu -- inference rendering, compiled with -fmad=false (see ngp_common.cuh).
//
// Reference: Testbed::render_nerf (src/testbed_nerf.cu:2670-2805) = NerfTracer::init_rays_from_camera (:2397-2477;
// init_rays_with_payload_kernel_nerf :2206-2329, advance_pos_nerf :797-845) -> NerfTracer::trace (:2479-2600; a host loop
// of compact_kernel_nerf + cudaStreamSynchronize, generate_next_nerf_network_inputs :881-937, the full network on
// n_alive x 1..8 samples with the AoS inputs / SoA outputs round-tripping through HBM, composite_kernel_nerf :939-1168)
// -> shade_kernel_nerf (:2148-2181) -> CudaRenderBuffer::accumulate / ::tonemap (src/render_buffer.cu:217-252, 474-501).
//
// B200 design: three launches per frame and no host synchronisation.
//   k_render_init    one thread per pixel: camera ray, box entry, Sobol start jitter, skip to the first occupied cell;
//                    rays that survive are appended to a compact queue (warp-aggregated counter).
//   k_render_march   persistent; every warp owns 32 ray slots (one per lane).  Per iteration each slot marches to its next
//                    sample, the warp runs the whole network on the samples 16 at a time with the same register/shared-memory
//                    pipeline as the training forward (forward_tile), and each lane composites its own sample.  A slot
//                    whose ray leaves the box or saturates pulls the next ray from the queue, so tiles stay full until the
//                    queue drains.  Sample coordinates and the 16-wide network outputs never touch HBM.
//   k_render_finish  per pixel: shade (sRGB -> linear), running mean over spp, background, exposure, output transfer.
// A ray's colour does not depend on how the reference chunks the wavefront (1..8 steps between compactions): every ray
// consumes its own samples in order, which is what a slot does here.
#include "network.cuh"
#include "ngp_common.cuh"
#include "ngp_kernels.h"

#define RENDER_WARPS 8
static constexpr int RENDER_SMEM = WeightSmem::FWD_END * 2 + RENDER_WARPS * WarpSlab::FWD_END;

// ---- low-discrepancy sequence (random_val.cuh:169-322): Sobol dims 0/1 from their generating recurrences -------------
__host__ __device__ __forceinline__ uint32_t brev32(uint32_t x) {
#ifdef __CUDA_ARCH__
	return __brev(x);
#else
	x = ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
	x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
	x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
	x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
	return (x >> 16) | (x << 16);
#endif
}
__host__ __device__ __forceinline__ uint32_t sobol_dim(uint32_t index, uint32_t dim) {
	if (dim == 0) return brev32(index);
	uint32_t X = 0, v = 0x80000000u;
	#pragma unroll
	for (uint32_t bit = 0; bit < 32; ++bit) {
		if ((index >> bit) & 1u) X ^= v;
		v ^= v >> 1;
	}
	return X;
}
__host__ __device__ __forceinline__ uint32_t hash_combine(uint32_t seed, uint32_t v) { return seed ^ (v + (seed << 6) + (seed >> 2)); }
__host__ __device__ __forceinline__ uint32_t laine_karras_permutation(uint32_t x, uint32_t seed) {
	x += seed;
	x ^= x * 0x6c50b47cu;
	x ^= x * 0xb82f1e52u;
	x ^= x * 0xc7afe638u;
	x ^= x * 0x8d22f6e6u;
	return x;
}
__host__ __device__ __forceinline__ uint32_t nested_uniform_scramble_base2(uint32_t x, uint32_t seed) { return brev32(laine_karras_permutation(brev32(x), seed)); }
__device__ __forceinline__ float ld_random_val(uint32_t index, uint32_t seed) {
	index = nested_uniform_scramble_base2(index, seed);
	return (float)nested_uniform_scramble_base2(sobol_dim(index, 0), hash_combine(seed, 0)) * 2.3283064365386963e-10f;   // 2^-32
}
__host__ __device__ __forceinline__ void ld_random_val_2d(uint32_t index, uint32_t seed, float out[2]) {
	index = nested_uniform_scramble_base2(index, seed);
	out[0] = (float)nested_uniform_scramble_base2(sobol_dim(index, 0), hash_combine(seed, 0)) * 2.3283064365386963e-10f;
	out[1] = (float)nested_uniform_scramble_base2(sobol_dim(index, 1), hash_combine(seed, 1)) * 2.3283064365386963e-10f;
}

struct RenderView {
	int32_t width, height;
	float focal[2], screen_center[2];
	float camera[12];
	float pixel_offset[2];      // ld_random_pixel_offset(snap ? 0 : spp): the same for every pixel of a frame, evaluated once on the host (integer + exact float ops)
	uint32_t sample_index;
	int32_t x0, y0, w, h;       // tile
	int32_t row_block, row_stride;
};

struct RayPayload { float o[3], d[3], t; uint32_t idx; };   // 32 bytes

// ---- k_render_init ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_render_init(RenderView v, Aabb render_aabb, GlobalMove gmv, const uint8_t* __restrict__ bitfield, float cone_angle,
              RayPayload* __restrict__ queue, uint32_t* __restrict__ counters, float4* __restrict__ ray_rgba, float* __restrict__ ray_depth) {
	const uint32_t tp = blockIdx.x * blockDim.x + threadIdx.x;   // tile pixel
	const bool in_tile = tp < (uint32_t)(v.w * v.h);
	bool alive = false;
	RayPayload P;
	if (in_tile) {
		ray_rgba[tp] = make_float4(0.f, 0.f, 0.f, 0.f);
		ray_depth[tp] = 0.f;
		const int ry = (int)(tp / (uint32_t)v.w);
		const int px = v.x0 + (int)(tp % (uint32_t)v.w), py = v.y0 + (v.row_block > 0 ? (ry / v.row_block) * v.row_block * v.row_stride + ry % v.row_block : ry);
		const uint32_t idx = (uint32_t)px + (uint32_t)v.width * (uint32_t)py;
		// pixel_to_ray (common_device.cuh:246-299)
		const float off[2] = {v.pixel_offset[0], v.pixel_offset[1]};
		const float uv[2] = {((float)px + off[0]) / (float)v.width, ((float)py + off[1]) / (float)v.height};
		const float dl[3] = {(uv[0] - v.screen_center[0]) * (float)v.width / v.focal[0], (uv[1] - v.screen_center[1]) * (float)v.height / v.focal[1], 1.0f};
		float o[3], du[3], d[3];
		#pragma unroll
		for (int k = 0; k < 3; ++k) {
			du[k] = fmaf(v.camera[k * 4 + 2], dl[2], fmaf(v.camera[k * 4 + 1], dl[1], v.camera[k * 4 + 0] * dl[0]));
			o[k] = v.camera[k * 4 + 3];
		}
		normalize3(du, d);
		#pragma unroll
		for (int k = 0; k < 3; ++k) o[k] -= gmv.first_frame_offset[k];
		global_move(gmv, o, d);
		float tmin, tmax;
		ray_intersect(render_aabb, o, d, &tmin, &tmax);
		float t = fmaxf(tmin, 0.2f) + 1e-6f;   // NERF_RENDERING_NEAR_DISTANCE
		float pos[3];
		#pragma unroll
		for (int k = 0; k < 3; ++k) pos[k] = fmaf(d[k], t, o[k]);
		if (aabb_contains(render_aabb, pos)) {
			// advance_pos_nerf (:797-845)
			const float idir[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
			float dt = calc_dt(t, cone_angle);
			t = fmaf(ld_random_val(v.sample_index, idx * 786433u), dt, t);
			alive = true;
			while (true) {
				#pragma unroll
				for (int k = 0; k < 3; ++k) pos[k] = fmaf(d[k], t, o[k]);
				if (!aabb_contains(render_aabb, pos)) { alive = false; break; }
				dt = calc_dt(t, cone_angle);
				const uint32_t mip = (uint32_t)mip_from_dt(dt, pos);
				if (density_grid_occupied_at(pos, bitfield, mip)) break;
				t = advance_to_next_voxel(t, cone_angle, pos, d, idir, NERF_GRIDSIZE >> mip);
			}
			#pragma unroll
			for (int k = 0; k < 3; ++k) { P.o[k] = o[k]; P.d[k] = d[k]; }
			P.t = t; P.idx = tp;
		}
	}
	// warp-aggregated append
	const uint32_t m = __ballot_sync(0xffffffffu, alive);
	if (m) {
		const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
		uint32_t base = 0;
		if (lane == leader) base = atomicAdd(counters, (uint32_t)__popc(m));
		base = __shfl_sync(0xffffffffu, base, leader);
		if (alive) {
			const uint32_t slot = base + __popc(m & ((1u << lane) - 1u));
			float4* q = reinterpret_cast<float4*>(queue + slot);
			q[0] = make_float4(P.o[0], P.o[1], P.o[2], P.d[0]);
			q[1] = make_float4(P.d[1], P.d[2], P.t, __uint_as_float(P.idx));
		}
	}
}

// ---- k_render_march --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RENDER_WARPS * 32, 2)
k_render_march(GridMeta gm, ParamLayout pl, const __half* __restrict__ params, float sdf_bias, Aabb render_aabb, Aabb train_aabb,
               const uint8_t* __restrict__ bitfield, float cone_angle, float cos_anneal_ratio, float min_transmittance,
               float cam_fwd0, float cam_fwd1, float cam_fwd2, float cam_pos0, float cam_pos1, float cam_pos2,
               const RayPayload* __restrict__ queue, uint32_t* __restrict__ counters /* [0] = queue length, [1] = queue head, [2] = samples evaluated */,
               float4* __restrict__ ray_rgba, float* __restrict__ ray_depth, uint32_t* __restrict__ ray_samples) {
	extern __shared__ __align__(16) unsigned char smem[];
	__half* sw = reinterpret_cast<__half*>(smem);
	stage_weights(sw, params, pl, false, threadIdx.x, blockDim.x);
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	unsigned char* slab = smem + WeightSmem::FWD_END * 2 + warp * WarpSlab::FWD_END;
	float* sCoord = reinterpret_cast<float*>(slab + WarpSlab::COORD);
	const __half* sOUT = reinterpret_cast<const __half*>(slab + WarpSlab::OUT);
	const uint32_t n_queue = counters[0];
	const float tdiag[3] = {train_aabb.mx[0] - train_aabb.mn[0], train_aabb.mx[1] - train_aabb.mn[1], train_aabb.mx[2] - train_aabb.mn[2]};

	// per-lane ray slot: ALL 32 lanes own a ray (marching and compositing are per-lane work); the network runs twice per
	// iteration, on the samples of lanes 0..15 and then of lanes 16..31 (a half with no live ray is skipped)
	bool have = false;
	float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 1.f}, idir[3] = {0.f, 0.f, 1.f}, t = 0.f;
	float rgba[4] = {0.f, 0.f, 0.f, 0.f}, depth = 0.f, max_weight = 0.f;
	uint32_t idx = 0, n_samples = 0, total_samples = 0;
	bool exhausted = false;
	FwdState st;

	while (true) {
		float c[7];   // this lane's sample (NerfCoordinate), kept in registers until its half's turn
		while (true) {
			if (!have) {
				if (exhausted) break;
				const uint32_t r = atomicAdd(counters + 1, 1u);
				if (r >= n_queue) { exhausted = true; break; }
				const float4* q = reinterpret_cast<const float4*>(queue + r);
				const float4 a = q[0], b = q[1];
				o[0] = a.x; o[1] = a.y; o[2] = a.z; d[0] = a.w; d[1] = b.x; d[2] = b.y; t = b.z; idx = __float_as_uint(b.w);
				idir[0] = 1.0f / d[0]; idir[1] = 1.0f / d[1]; idir[2] = 1.0f / d[2];
				rgba[0] = rgba[1] = rgba[2] = rgba[3] = 0.f; depth = 0.f; max_weight = 0.f; n_samples = 0;
				have = true;
			}
			// generate_next_nerf_network_inputs, one step (:919-936)
			float pos[3], dt = 0.f;
			bool inside = true;
			while (true) {
				#pragma unroll
				for (int k = 0; k < 3; ++k) pos[k] = fmaf(d[k], t, o[k]);
				if (!aabb_contains(render_aabb, pos)) { inside = false; break; }
				dt = calc_dt(t, cone_angle);
				const uint32_t mip = (uint32_t)mip_from_dt(dt, pos);
				if (density_grid_occupied_at(pos, bitfield, mip)) break;
				t = advance_to_next_voxel(t, cone_angle, pos, d, idir, NERF_GRIDSIZE >> mip);
			}
			if (!inside) {   // the ray left the box: it is compacted away (compact_kernel_nerf :2183-2204)
				if (rgba[3] > 0.001f) { ray_rgba[idx] = make_float4(rgba[0], rgba[1], rgba[2], rgba[3]); ray_depth[idx] = depth; }
				if (ray_samples) ray_samples[idx] = n_samples;
				have = false;
				continue;
			}
			#pragma unroll
			for (int k = 0; k < 3; ++k) c[k] = (pos[k] - train_aabb.mn[k]) / tdiag[k];
			c[3] = warp_dt(dt);
			#pragma unroll
			for (int k = 0; k < 3; ++k) c[4 + k] = (d[k] + 1.0f) * 0.5f;
			t += dt;
			break;
		}
		if (!have) {
			#pragma unroll
			for (int k = 0; k < 7; ++k) c[k] = 0.5f;
		}
		const uint32_t active = __ballot_sync(0xffffffffu, have);
		if (!active) break;
		#pragma unroll 1
		for (int half = 0; half < 2; ++half) {
			if (!((active >> (16 * half)) & 0xffffu)) continue;        // warp-uniform
			const bool mine = (lane >> 4) == half;
			if (mine) {
				float* cs = sCoord + (lane & 15) * 7;
				#pragma unroll
				for (int k = 0; k < 7; ++k) cs[k] = c[k];
			}
			__syncwarp();
			forward_tile<false>(gm, params, pl, sw, slab, sdf_bias, lane, st);
			// composite_kernel_nerf (:939-1168), one sample
			if (mine && have) {
				const __half* orow = sOUT + (lane & 15) * 16;
				float dv[3] = {__half2float(orow[8]) * 2.0f - 1.0f, __half2float(orow[9]) * 2.0f - 1.0f, __half2float(orow[10]) * 2.0f - 1.0f}, dir[3];   // BENT_DIR
				normalize3(dv, dir);
				const SampleEval e = eval_sample(orow, c, dir, cos_anneal_ratio);
				const float T = 1.f - rgba[3];
				const float weight = e.alpha * T;
				#pragma unroll
				for (int k = 0; k < 3; ++k) rgba[k] = fmaf(e.rgb[k], weight, rgba[k]);
				rgba[3] += weight;
				++n_samples; ++total_samples;
				if (weight > max_weight) {
					max_weight = weight;
					float pos[3];
					#pragma unroll
					for (int k = 0; k < 3; ++k) pos[k] = fmaf(c[k], tdiag[k], train_aabb.mn[k]);
					depth = fmaf(cam_fwd2, pos[2] - cam_pos2, fmaf(cam_fwd1, pos[1] - cam_pos1, cam_fwd0 * (pos[0] - cam_pos0)));
				}
				if (rgba[3] > (1.0f - min_transmittance)) {
					const float w = rgba[3];
					if (w > 0.001f) { ray_rgba[idx] = make_float4(rgba[0] / w, rgba[1] / w, rgba[2] / w, w / w); ray_depth[idx] = depth; }
					if (ray_samples) ray_samples[idx] = n_samples;
					have = false;
				}
			}
			__syncwarp();
		}
	}
	// statistics: samples pushed through the network
	total_samples += __shfl_xor_sync(0xffffffffu, total_samples, 1);
	total_samples += __shfl_xor_sync(0xffffffffu, total_samples, 2);
	total_samples += __shfl_xor_sync(0xffffffffu, total_samples, 4);
	total_samples += __shfl_xor_sync(0xffffffffu, total_samples, 8);
	total_samples += __shfl_xor_sync(0xffffffffu, total_samples, 16);
	if (lane == 0 && total_samples) atomicAdd(counters + 2, total_samples);
}

// ---- k_render_finish -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_render_finish(uint32_t n, const float4* __restrict__ ray_rgba, const float* __restrict__ ray_depth, float sample_count, float4* __restrict__ accumulate,
                float bg0, float bg1, float bg2, float bg3, float exposure, int to_srgb, float4* __restrict__ out, float* __restrict__ depth_out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float4 r = ray_rgba[i];
	float tmp[4] = {srgb_to_linear(r.x), srgb_to_linear(r.y), srgb_to_linear(r.z), r.w};   // shade_kernel_nerf: accumulate in linear colours
	if (depth_out) depth_out[i] = tmp[3] > 0.2f ? ray_depth[i] : 1e10f;
	float4 a4 = sample_count > 0.f ? accumulate[i] : make_float4(0.f, 0.f, 0.f, 0.f);
	float acc[4] = {a4.x, a4.y, a4.z, a4.w};
	#pragma unroll
	for (int k = 0; k < 4; ++k) acc[k] = fmaf(acc[k], sample_count, tmp[k]) / (sample_count + 1);
	accumulate[i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
	const float bg[3] = {srgb_to_linear(bg0), srgb_to_linear(bg1), srgb_to_linear(bg2)};
	const float weight = (1 - acc[3]) * bg3;
	float col[4] = {fmaf(bg[0], weight, acc[0]), fmaf(bg[1], weight, acc[1]), fmaf(bg[2], weight, acc[2]), acc[3] + weight};
	const float scale = powf(2.0f, exposure);
	#pragma unroll
	for (int k = 0; k < 3; ++k) { col[k] *= scale; if (to_srgb) col[k] = linear_to_srgb(col[k]); }
	out[i] = make_float4(col[0], col[1], col[2], col[3]);
}

// ---- launcher --------------------------------------------------------------------------------------------------------
int launch_render(cudaStream_t stream, const GridMeta& gm, const ParamLayout& pl, const __half* params, float sdf_bias, const RenderLaunch& r,
                  const Aabb& render_aabb, const Aabb& train_aabb, const GlobalMove& gmv, const uint8_t* bitfield, float cone_angle, float cos_anneal_ratio,
                  void* queue, uint32_t* counters, float* ray_rgba, float* ray_depth, uint32_t* ray_samples, float* accumulate, float* out, float* depth_out) {
	static bool attr[NEUS_MAX_DEVICES] = {};
	if (ensure_dynamic_smem(k_render_march, RENDER_SMEM, attr)) return -1;
	const uint32_t n = (uint32_t)r.tile_w * (uint32_t)r.tile_h;
	if (n == 0) return 0;
	RenderView v;
	v.width = r.width; v.height = r.height;
	v.focal[0] = r.focal[0]; v.focal[1] = r.focal[1]; v.screen_center[0] = r.screen_center[0]; v.screen_center[1] = r.screen_center[1];
	for (int k = 0; k < 12; ++k) v.camera[k] = r.camera[k];
	v.sample_index = r.sample_index;
	{   // ld_random_pixel_offset (random_val.cuh:316-321)
		float a[2], b[2];
		ld_random_val_2d(0, 0xdeadbeefu, a);
		ld_random_val_2d(r.snap_to_pixel_centers ? 0u : r.sample_index, 0xdeadbeefu, b);
		for (int k = 0; k < 2; ++k) { const float o = (0.5f - a[k]) + b[k]; v.pixel_offset[k] = o - floorf(o); }
	}
	v.x0 = r.tile_x0; v.y0 = r.tile_y0; v.w = r.tile_w; v.h = r.tile_h;
	v.row_block = r.row_block; v.row_stride = r.row_stride;
	CUDA_CHECK(cudaMemsetAsync(counters, 0, 4 * sizeof(uint32_t), stream));
	k_render_init<<<(n + 127) / 128, 128, 0, stream>>>(v, render_aabb, gmv, bitfield, cone_angle, (RayPayload*)queue, counters, (float4*)ray_rgba, ray_depth);
	const int sms = device_sm_count();
	k_render_march<<<sms * 2, RENDER_WARPS * 32, RENDER_SMEM, stream>>>(gm, pl, params, sdf_bias, render_aabb, train_aabb, bitfield, cone_angle, cos_anneal_ratio,
		r.min_transmittance, r.camera[2], r.camera[6], r.camera[10], r.camera[3], r.camera[7], r.camera[11],
		(const RayPayload*)queue, counters, (float4*)ray_rgba, ray_depth, ray_samples);
	k_render_finish<<<(n + 255) / 256, 256, 0, stream>>>(n, (const float4*)ray_rgba, ray_depth, (float)r.sample_index, (float4*)accumulate,
		r.background_color[0], r.background_color[1], r.background_color[2], r.background_color[3], r.exposure, r.to_srgb, (float4*)out, depth_out);
	COUNT_LAUNCH(3);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}
