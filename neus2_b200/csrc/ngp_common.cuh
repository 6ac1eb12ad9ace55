// ngp_common.cuh -- device helpers for the ray-marching side (sample generation, loss, occupancy grid, render).
// Files including this header are compiled with -fmad=false: a fused multiply-add happens exactly where fmaf()
// is written, which is what makes per-ray sample counts bit-identical to oracle/neus_oracle.cpp.
// Restates src/testbed_nerf.cu:57-65,149-164,356-378,445-520,624-638,1226-1261 and
// include/neural-graphics-primitives/{bounding_box.cuh:163-221, common_device.cuh:30-61,591-656}.
#pragma once
#include "common.cuh"

#define NERF_GRIDSIZE 128u
#define NERF_STEPS 1024u
#define NERF_CASCADES 8u
#define N_MAX_RANDOM_SAMPLES_PER_RAY 8u

struct DatasetDev {
	uint32_t n_images;
	int32_t width, height;
	const uint8_t* pixels;     // [n][h][w][4] RGBA8
	const float* focal;        // [n][2]
	const float* principal;    // [n][2] relative
	const float* xforms;       // [n][3][4] row-major
};

struct Aabb { float mn[3], mx[3]; };

struct GlobalMove {            // accumulated rotation (3x3) / translation, stored in half in the reference (nerf_network.h:1084-1090)
	float R[9];
	float T[3];
	float first_frame_offset[3];
};

#ifdef __CUDACC__
__host__ __device__ __forceinline__ float min_cone_stepsize() { return 1.73205080757f / 1024; }
__host__ __device__ __forceinline__ float max_cone_stepsize() { return min_cone_stepsize() * (1 << (NERF_CASCADES - 1)) * NERF_STEPS / NERF_GRIDSIZE; }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); }
__device__ __forceinline__ float calc_dt(float t, float cone_angle) { return clampf(t * cone_angle, min_cone_stepsize(), max_cone_stepsize()); }

__host__ __device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}
__host__ __device__ __forceinline__ uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) { return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2); }
__host__ __device__ __forceinline__ uint32_t morton3D_invert(uint32_t x) {
	x = x & 0x49249249;
	x = (x | (x >> 2)) & 0xc30c30c3;
	x = (x | (x >> 4)) & 0x0f00f00f;
	x = (x | (x >> 8)) & 0xff0000ff;
	x = (x | (x >> 16)) & 0x0000ffff;
	return x;
}

__device__ __forceinline__ int mip_from_pos(const float p[3], uint32_t max_cascade = NERF_CASCADES - 1) {
	int exponent;
	float maxval = fmaxf(fmaxf(fabsf(p[0] - 0.5f), fabsf(p[1] - 0.5f)), fabsf(p[2] - 0.5f));
	frexpf(maxval, &exponent);
	return min((int)max_cascade, max(0, exponent + 1));
}
__device__ __forceinline__ int mip_from_dt(float dt, const float p[3], uint32_t max_cascade = NERF_CASCADES - 1) {
	int mip = mip_from_pos(p, max_cascade);
	dt *= 2 * NERF_GRIDSIZE;
	if (dt < 1.f) return mip;
	int exponent;
	frexpf(dt, &exponent);
	return min((int)max_cascade, max(exponent, mip));
}
__device__ __forceinline__ uint32_t cascaded_grid_idx_at(const float pin[3], uint32_t mip) {
	float mip_scale = scalbnf(1.0f, -(int)mip);
	int i[3];
	#pragma unroll
	for (int d = 0; d < 3; ++d) {
		float p = pin[d] - 0.5f;
		p *= mip_scale;
		p += 0.5f;
		i[d] = (int)(p * NERF_GRIDSIZE);
		i[d] = min(max(i[d], 0), (int)NERF_GRIDSIZE - 1);
	}
	return morton3D(i[0], i[1], i[2]);
}
__device__ __forceinline__ bool density_grid_occupied_at(const float p[3], const uint8_t* __restrict__ bitfield, uint32_t mip) {
	uint32_t idx = cascaded_grid_idx_at(p, mip);
	return bitfield[idx / 8 + (NERF_GRIDSIZE * NERF_GRIDSIZE * NERF_GRIDSIZE * mip) / 8] & (1 << (idx % 8));
}
__device__ __forceinline__ float signf1(float x) { return copysignf(1.0f, x); }
__device__ __forceinline__ float distance_to_next_voxel(const float pos[3], const float dir[3], const float idir[3], uint32_t res) {
	float p[3] = {(float)res * pos[0], (float)res * pos[1], (float)res * pos[2]};
	float tx = (floorf(p[0] + 0.5f + 0.5f * signf1(dir[0])) - p[0]) * idir[0];
	float ty = (floorf(p[1] + 0.5f + 0.5f * signf1(dir[1])) - p[1]) * idir[1];
	float tz = (floorf(p[2] + 0.5f + 0.5f * signf1(dir[2])) - p[2]) * idir[2];
	float t = fminf(fminf(tx, ty), tz);
	return fmaxf(t * (1.0f / (float)res), 0.0f);   // res is a power of two (128 >> mip): the product is the quotient, bit for bit, without the IEEE division sequence
}
__device__ __forceinline__ float advance_to_next_voxel(float t, float cone_angle, const float pos[3], const float dir[3], const float idir[3], uint32_t res) {
	float t_target = t + distance_to_next_voxel(pos, dir, idir, res);
	do { t += calc_dt(t, cone_angle); } while (t < t_target);
	return t;
}
__device__ __forceinline__ float srgb_to_linear(float s) { return s <= 0.04045f ? s / 12.92f : powf((s + 0.055f) / 1.055f, 2.4f); }
__device__ __forceinline__ float linear_to_srgb(float l) { return l < 0.0031308f ? 12.92f * l : 1.055f * powf(l, 0.41666f) - 0.055f; }

__device__ __forceinline__ void read_rgba(const DatasetDev& ds, uint32_t img, int px, int py, float out[4]) {
	const uint32_t word = reinterpret_cast<const uint32_t*>(ds.pixels)[((size_t)img * ds.height + py) * ds.width + px];
	if (word == 0x00FF00FFu) { out[0] = out[1] = out[2] = out[3] = -1.0f; return; }
	const float alpha = (float)(word >> 24) * (1.0f / 255.0f);
	out[0] = srgb_to_linear((float)(word & 0xff) * (1.0f / 255.0f)) * alpha;
	out[1] = srgb_to_linear((float)((word >> 8) & 0xff) * (1.0f / 255.0f)) * alpha;
	out[2] = srgb_to_linear((float)((word >> 16) & 0xff) * (1.0f / 255.0f)) * alpha;
	out[3] = alpha;
}
__device__ __forceinline__ void image_pos(const float xy[2], int w, int h, int* px, int* py) {
	*px = max(min((int)(xy[0] * (float)w), w - 1), 0);
	*py = max(min((int)(xy[1] * (float)h), h - 1), 0);
}
__device__ __forceinline__ uint32_t image_idx(uint32_t base_idx, uint32_t n_rays, uint32_t n_rays_total, uint32_t n_images) {
	return (((base_idx + n_rays_total) * n_images) / n_rays) % n_images;
}
__device__ __forceinline__ void random_image_pos_training(Pcg32& rng, int w, int h, bool snap, float xy[2]) {
	xy[0] = rng.next_float(); xy[1] = rng.next_float();
	if (snap) {
		int ix = min(max((int)(xy[0] * (float)w), 0), w - 1);
		int iy = min(max((int)(xy[1] * (float)h), 0), h - 1);
		xy[0] = ((float)ix + 0.5f) / (float)w;
		xy[1] = ((float)iy + 0.5f) / (float)h;
	}
}
__device__ __forceinline__ void make_ray(const DatasetDev& ds, uint32_t img, const float xy[2], float o[3], float d[3]) {
	const float* X = ds.xforms + (size_t)img * 12;
	const float fx = ds.focal[img * 2], fy = ds.focal[img * 2 + 1];
	const float cx = ds.principal[img * 2], cy = ds.principal[img * 2 + 1];
	const float dl[3] = {(xy[0] - cx) * (float)ds.width / fx, (xy[1] - cy) * (float)ds.height / fy, 1.0f};
	#pragma unroll
	for (int k = 0; k < 3; ++k) {
		o[k] = X[k * 4 + 3];
		d[k] = fmaf(X[k * 4 + 2], dl[2], fmaf(X[k * 4 + 1], dl[1], X[k * 4 + 0] * dl[0]));
	}
}
__device__ __forceinline__ void normalize3(const float v[3], float out[3]) {
	float n2 = fmaf(v[2], v[2], fmaf(v[1], v[1], v[0] * v[0]));
	float n = sqrtf(n2);
	if (n2 > 0) { out[0] = v[0] / n; out[1] = v[1] / n; out[2] = v[2] / n; } else { out[0] = v[0]; out[1] = v[1]; out[2] = v[2]; }
}
__device__ __forceinline__ void global_move(const GlobalMove& gmv, float o[3], float d[3]) {
	float no[3], nd[3];
	#pragma unroll
	for (int k = 0; k < 3; ++k) {
		no[k] = fmaf(gmv.R[k * 3 + 2], o[2], fmaf(gmv.R[k * 3 + 1], o[1], gmv.R[k * 3] * o[0])) + gmv.T[k];
		nd[k] = fmaf(gmv.R[k * 3 + 2], d[2], fmaf(gmv.R[k * 3 + 1], d[1], gmv.R[k * 3] * d[0]));
	}
	#pragma unroll
	for (int k = 0; k < 3; ++k) { o[k] = no[k]; d[k] = nd[k]; }
}
__device__ __forceinline__ void ray_intersect(const Aabb& bb, const float o[3], const float d[3], float* tmin_out, float* tmax_out) {
	const float FMAXV = 3.402823466e+38f;
	float tmin = (bb.mn[0] - o[0]) / d[0], tmax = (bb.mx[0] - o[0]) / d[0];
	if (tmin > tmax) { float s = tmin; tmin = tmax; tmax = s; }
	float tymin = (bb.mn[1] - o[1]) / d[1], tymax = (bb.mx[1] - o[1]) / d[1];
	if (tymin > tymax) { float s = tymin; tymin = tymax; tymax = s; }
	if (tmin > tymax || tymin > tmax) { *tmin_out = FMAXV; *tmax_out = FMAXV; return; }
	if (tymin > tmin) tmin = tymin;
	if (tymax < tmax) tmax = tymax;
	float tzmin = (bb.mn[2] - o[2]) / d[2], tzmax = (bb.mx[2] - o[2]) / d[2];
	if (tzmin > tzmax) { float s = tzmin; tzmin = tzmax; tzmax = s; }
	if (tmin > tzmax || tzmin > tmax) { *tmin_out = FMAXV; *tmax_out = FMAXV; return; }
	if (tzmin > tmin) tmin = tzmin;
	if (tzmax < tmax) tmax = tzmax;
	*tmin_out = tmin; *tmax_out = tmax;
}
__device__ __forceinline__ bool aabb_contains(const Aabb& bb, const float p[3]) {
	return p[0] >= bb.mn[0] && p[0] <= bb.mx[0] && p[1] >= bb.mn[1] && p[1] <= bb.mx[1] && p[2] >= bb.mn[2] && p[2] <= bb.mx[2];
}
__device__ __forceinline__ float warp_dt(float dt) {
	float max_stepsize = min_cone_stepsize() * (1 << (NERF_CASCADES - 1));
	return (dt - min_cone_stepsize()) / (max_stepsize - min_cone_stepsize());
}
__device__ __forceinline__ float unwarp_dt(float dt) {
	float max_stepsize = min_cone_stepsize() * (1 << (NERF_CASCADES - 1));
	return dt * (max_stepsize - min_cone_stepsize()) + min_cone_stepsize();
}

// One ray's march through the occupancy bitfield (testbed_nerf.cu:1391-1449): returns the number of samples;
// emits up to `limit` NerfCoordinates (7 floats) when coords != nullptr.
__device__ __forceinline__ uint32_t march_ray(const Aabb& bb, const float ray_o[3], const float dir[3], float startt, float cone_angle,
                                              const uint8_t* __restrict__ bitfield, uint32_t limit, float* __restrict__ coords) {
	const float idir[3] = {1.0f / dir[0], 1.0f / dir[1], 1.0f / dir[2]};
	const float wdir[3] = {(dir[0] + 1.0f) * 0.5f, (dir[1] + 1.0f) * 0.5f, (dir[2] + 1.0f) * 0.5f};
	const float diag[3] = {bb.mx[0] - bb.mn[0], bb.mx[1] - bb.mn[1], bb.mx[2] - bb.mn[2]};
	uint32_t j = 0;
	float t = startt;
	float pos[3];
	while (true) {
		#pragma unroll
		for (int k = 0; k < 3; ++k) pos[k] = fmaf(t, dir[k], ray_o[k]);
		if (!(aabb_contains(bb, pos) && j < limit)) break;
		const float dt = calc_dt(t, cone_angle);
		const uint32_t mip = (uint32_t)mip_from_dt(dt, pos);
		if (density_grid_occupied_at(pos, bitfield, mip)) {
			if (coords) {
				float* c = coords + (size_t)j * 7;
				#pragma unroll
				for (int k = 0; k < 3; ++k) c[k] = (pos[k] - bb.mn[k]) / diag[k];
				c[3] = warp_dt(dt);
				c[4] = wdir[0]; c[5] = wdir[1]; c[6] = wdir[2];
			}
			++j;
			t += dt;
		} else {
			const uint32_t res = NERF_GRIDSIZE >> mip;
			t = advance_to_next_voxel(t, cone_angle, pos, dir, idir, res);
		}
	}
	return j;
}

// ---- NeuS per-sample evaluation shared by the training loss (loss.cu) and the inference render (render.cu) ----
// testbed_nerf.cu:1575-1625 (training) == :1003-1037 (composite_kernel_nerf): rgb = logistic, inv_s = exp(10*var) with the
// product taken in half, alpha from the logistic CDF of the sdf half a step before / after the sample.
__device__ __forceinline__ float logistic_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float relu_f(float v) { return v > 0.0f ? v : 0.0f; }

struct SampleEval { float alpha, p_div_c, true_cos, iter_cos, inv_s, sdf, next_sdf, dt; float rgb[3]; float g[3]; };

__device__ __forceinline__ SampleEval eval_sample(const __half* __restrict__ o_ptr, const float* __restrict__ coord, const float dir[3], float cos_anneal_ratio) {
	// one 32-byte row of network output
	uint4 lo = reinterpret_cast<const uint4*>(o_ptr)[0], hi = reinterpret_cast<const uint4*>(o_ptr)[1];
	const __half* o = reinterpret_cast<const __half*>(&lo);
	const __half* o2 = reinterpret_cast<const __half*>(&hi);
	(void)o2;
	SampleEval e;
	#pragma unroll
	for (int k = 0; k < 3; ++k) e.rgb[k] = logistic_f(__half2float(o[k]));
	e.dt = unwarp_dt(coord[3]);
	e.inv_s = expf(__half2float(__hmul(__float2half_rn(10.0f), o[7])));     // :1590
	e.sdf = __half2float(o[3]);
	#pragma unroll
	for (int k = 0; k < 3; ++k) e.g[k] = __half2float(o[4 + k]);
	e.true_cos = (dir[0] * e.g[0] + dir[1] * e.g[1] + dir[2] * e.g[2]);
	e.iter_cos = (float)-(relu_f((float)(-e.true_cos * 0.5 + 0.5)) * (1.0 - cos_anneal_ratio) + relu_f(-e.true_cos) * cos_anneal_ratio);
	e.next_sdf = (float)(e.sdf + e.iter_cos * e.dt * 0.5);
	const float prev_sdf = (float)(e.sdf - e.iter_cos * e.dt * 0.5);
	const float next_cdf = logistic_f(e.next_sdf * e.inv_s);
	const float prev_cdf = logistic_f(prev_sdf * e.inv_s);
	const float p = prev_cdf - next_cdf;
	const float c = prev_cdf;
	e.p_div_c = (p + 1e-5f) / (c + 1e-5f);
	e.alpha = clampf(e.p_div_c, 0.0f, 1.0f);
	return e;
}

#endif
