// mesh.cu -- SDF lattice evaluation and marching cubes, compiled with -fmad=false.
//
// Reference: Testbed::marching_cubes (src/testbed_nerf.cu:4175-4226) = get_density_on_grid (:4096-4143;
// generate_grid_samples_nerf_uniform :596-608 writes 12 B/point of positions, NerfNetwork::sdf runs in 2^20 batches,
// grid_samples_half_to_float :692-708) + marching_cubes_gpu (src/marching_cubes.cu:797-822; gen_vertices :276-331 and
// gen_faces :382-715 run twice each and number vertices / triangles by atomicAdd arrival; 3 x int32 per cell of
// vertex-index workspace).
//
// B200 design:
//   k_sdf_lattice   positions are computed from the lattice index inside the SDF kernel (no position buffer, no batches);
//                   a warp owns 16 consecutive points along x, so its gathers hit neighbouring grid cells.
//   marching cubes  deterministic: vertices are numbered in ascending (cell, axis) order and triangles in ascending cell
//                   order by a block-count / scan / emit scheme, so the mesh is reproducible bit for bit; the per-cell
//                   workspace is one u32 (vertex prefix) instead of three int32, written only where the surface passes.
#include "network.cuh"
#include "ngp_common.cuh"
#include "ngp_kernels.h"
#include "mc_table.cuh"

#define LATTICE_WARPS 8
static constexpr int LATTICE_SMEM = WeightSmem::FWD_END * 2 + LATTICE_WARPS * WarpSlab::FWD_END;

__global__ void __launch_bounds__(LATTICE_WARPS * 32, 2)
k_sdf_lattice(GridMeta gm, ParamLayout pl, const __half* __restrict__ params, float sdf_bias, uint32_t res0, uint32_t res1, uint32_t res2,
              Aabb render_aabb, Aabb train_aabb, uint32_t z0, uint64_t n, float* __restrict__ out) {
	extern __shared__ __align__(16) unsigned char smem[];
	__half* sw = reinterpret_cast<__half*>(smem);
	stage_matrix(sw + WeightSmem::W0, LD_X0, params + pl.off_w0, 64, SDF_IN, threadIdx.x, blockDim.x);
	stage_matrix(sw + WeightSmem::W1, LD_H, params + pl.off_w1, 16, 64, threadIdx.x, blockDim.x);
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	unsigned char* slab = smem + WeightSmem::FWD_END * 2 + warp * WarpSlab::FWD_END;
	float* sCoord = reinterpret_cast<float*>(slab + WarpSlab::COORD);
	const __half* sSDF = reinterpret_cast<const __half*>(slab + WarpSlab::SDF);
	const float inv_res[3] = {1.f / (float)res0, 1.f / (float)res1, 1.f / (float)res2};
	const uint64_t n_tiles = (n + TILE - 1) / TILE;
	const uint64_t plane = (uint64_t)res0 * res1;
	for (uint64_t tile = (uint64_t)blockIdx.x * LATTICE_WARPS + warp; tile < n_tiles; tile += (uint64_t)gridDim.x * LATTICE_WARPS) {
		const uint64_t row0 = tile * TILE;
		if (lane < 16) {
			const uint64_t i = row0 + lane;
			float* c = sCoord + lane * 7;
			if (i < n) {   // generate_grid_samples_nerf_uniform (:596-608)
				const uint32_t zi = (uint32_t)(i / plane);
				const uint32_t rem = (uint32_t)(i - (uint64_t)zi * plane);
				const uint32_t xi[3] = {rem % res0, rem / res0, z0 + zi};
				#pragma unroll
				for (int k = 0; k < 3; ++k) {
					float p = (float)xi[k] * inv_res[k];
					p = fmaf(p, render_aabb.mx[k] - render_aabb.mn[k], render_aabb.mn[k]);
					c[k] = (p - train_aabb.mn[k]) / (train_aabb.mx[k] - train_aabb.mn[k]);
				}
			} else { c[0] = c[1] = c[2] = 0.f; }
		}
		__syncwarp();
		sdf_tile(gm, params + pl.off_grid, sw, slab, lane);
		if (lane < 16 && row0 + lane < n) out[row0 + lane] = __half2float(__hadd(sSDF[lane], __float2half_rn(sdf_bias)));
		__syncwarp();
	}
}

int launch_sdf_lattice(cudaStream_t stream, const GridMeta& gm, const ParamLayout& pl, const __half* params, float sdf_bias, const uint32_t res[3],
                       const Aabb& render_aabb, const Aabb& train_aabb, uint32_t z0, uint32_t z1, float* out) {
	static bool attr[NEUS_MAX_DEVICES] = {};
	if (ensure_dynamic_smem(k_sdf_lattice, LATTICE_SMEM, attr)) return -1;
	if (z1 <= z0) return 0;
	const uint64_t n = (uint64_t)res[0] * res[1] * (z1 - z0);
	const int sms = device_sm_count();
	uint64_t grid = (uint64_t)sms * 2;
	const uint64_t need = ((n + TILE - 1) / TILE + LATTICE_WARPS - 1) / LATTICE_WARPS;
	if (need < grid) grid = need;
	k_sdf_lattice<<<(uint32_t)grid, LATTICE_WARPS * 32, LATTICE_SMEM, stream>>>(gm, pl, params, sdf_bias, res[0], res[1], res[2], render_aabb, train_aabb, z0, n, out);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}

// ---- marching cubes --------------------------------------------------------------------------------------------------
#define MC_BLOCK 512

struct McGeom { uint32_t rx, ry, rz; uint64_t res1, res2, res3; };

// crossing flags of the three edges that start at cell (x,y,z): bit a set <=> edge along axis a carries a vertex
// (gen_vertices, marching_cubes.cu:276-331)
__device__ __forceinline__ uint32_t cell_edges(const McGeom& g, const float* __restrict__ density, float thresh, uint32_t x, uint32_t y, uint32_t z, uint64_t idx, float f0) {
	const bool inside = f0 > thresh;
	uint32_t m = 0;
	if (x < g.rx - 1 && inside != (density[idx + 1] > thresh)) m |= 1u;
	if (y < g.ry - 1 && inside != (density[idx + g.res1] > thresh)) m |= 2u;
	if (z < g.rz - 1 && inside != (density[idx + g.res2] > thresh)) m |= 4u;
	return m;
}
// case index of the cube whose lowest corner is cell idx (gen_faces, marching_cubes.cu:670-683)
__device__ __forceinline__ int cube_mask(const McGeom& g, const float* __restrict__ density, float thresh, uint64_t idx) {
	int mask = 0;
	if (density[idx] > thresh) mask |= 1;
	if (density[idx + 1] > thresh) mask |= 2;
	if (density[idx + 1 + g.res1] > thresh) mask |= 4;
	if (density[idx + g.res1] > thresh) mask |= 8;
	if (density[idx + g.res2] > thresh) mask |= 16;
	if (density[idx + g.res2 + 1] > thresh) mask |= 32;
	if (density[idx + g.res2 + 1 + g.res1] > thresh) mask |= 64;
	if (density[idx + g.res2 + g.res1] > thresh) mask |= 128;
	return mask;
}
__device__ __forceinline__ uint32_t tri_index_count(uint64_t tri) {
	uint32_t n = 0;
	for (; n < 15; n += 3) if (((tri >> (4 * n)) & 0xf) == 0xf) break;
	return n;
}
// per-cell counts: low 16 bits vertices (0..3), high 16 bits triangle indices (0..15)
__device__ __forceinline__ uint32_t cell_counts(const McGeom& g, const float* __restrict__ density, float thresh, uint64_t idx, uint32_t* edges_out, int* mask_out) {
	uint32_t packed = 0; *edges_out = 0; *mask_out = 0;
	if (idx < g.res3) {
		const uint32_t z = (uint32_t)(idx / g.res2);
		const uint32_t rem = (uint32_t)(idx - (uint64_t)z * g.res2);
		const uint32_t y = rem / g.rx, x = rem - y * g.rx;
		const float f0 = density[idx];
		const uint32_t e = cell_edges(g, density, thresh, x, y, z, idx, f0);
		*edges_out = e;
		packed = __popc(e);
		if (x < g.rx - 1 && y < g.ry - 1 && z < g.rz - 1) {
			const int mask = cube_mask(g, density, thresh, idx);
			if (mask && mask != 255) { *mask_out = mask; packed |= tri_index_count(c_mc_triangles[mask]) << 16; }
		}
	}
	return packed;
}
// block-wide exclusive scan of packed (16|16) counts; returns this thread's exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_scan_packed(uint32_t v, uint32_t* s_warp, uint32_t* total) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc = v;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
	if (lane == 31) s_warp[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		uint32_t w = lane < MC_BLOCK / 32 ? s_warp[lane] : 0u, winc = w;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= d) winc += o; }
		if (lane < MC_BLOCK / 32) s_warp[lane] = winc - w;
		if (lane == MC_BLOCK / 32 - 1) s_warp[32] = winc;
	}
	__syncthreads();
	*total = s_warp[32];
	return s_warp[warp] + inc - v;
}

__global__ void __launch_bounds__(MC_BLOCK)
k_mc_count(McGeom g, float thresh, const float* __restrict__ density, uint32_t* __restrict__ block_counts) {
	__shared__ uint32_t s_warp[33];
	const uint64_t idx = (uint64_t)blockIdx.x * MC_BLOCK + threadIdx.x;
	uint32_t e; int mask;
	const uint32_t packed = cell_counts(g, density, thresh, idx, &e, &mask);
	uint32_t total;
	block_scan_packed(packed, s_warp, &total);
	if (threadIdx.x == 0) { block_counts[2 * blockIdx.x] = total & 0xffffu; block_counts[2 * blockIdx.x + 1] = total >> 16; }
}

// exclusive scan of the per-block counts (in place), totals to counts[0..1]; one CTA
__global__ void __launch_bounds__(1024)
k_mc_scan(uint32_t n_blocks, uint32_t* __restrict__ block_counts, uint32_t* __restrict__ counts) {
	__shared__ uint32_t s_v[1024], s_i[1024];
	const uint32_t per = (n_blocks + 1023) / 1024;
	const uint32_t b0 = threadIdx.x * per, b1 = min(n_blocks, b0 + per);
	uint32_t sv = 0, si = 0;
	for (uint32_t b = b0; b < b1; ++b) { sv += block_counts[2 * b]; si += block_counts[2 * b + 1]; }
	s_v[threadIdx.x] = sv; s_i[threadIdx.x] = si;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t av = 0, ai = 0;
		for (int k = 0; k < 1024; ++k) { const uint32_t tv = s_v[k], ti = s_i[k]; s_v[k] = av; s_i[k] = ai; av += tv; ai += ti; }
		counts[0] = av; counts[1] = ai;
	}
	__syncthreads();
	uint32_t av = s_v[threadIdx.x], ai = s_i[threadIdx.x];
	for (uint32_t b = b0; b < b1; ++b) {
		const uint32_t tv = block_counts[2 * b], ti = block_counts[2 * b + 1];
		block_counts[2 * b] = av; block_counts[2 * b + 1] = ai;
		av += tv; ai += ti;
	}
}

// vertices + per-cell vertex prefix
__global__ void __launch_bounds__(MC_BLOCK)
k_mc_emit_verts(McGeom g, Aabb aabb, float thresh, const float* __restrict__ density, const uint32_t* __restrict__ block_offsets, uint32_t n_blocks,
                uint32_t n_verts_total, uint32_t* __restrict__ vert_prefix, float* __restrict__ verts) {
	__shared__ uint32_t s_warp[33];
	const uint32_t off = block_offsets[2 * blockIdx.x];
	const uint32_t next = blockIdx.x + 1 < n_blocks ? block_offsets[2 * blockIdx.x + 2] : n_verts_total;
	if (next == off) return;   // no surface in this block
	const uint64_t idx = (uint64_t)blockIdx.x * MC_BLOCK + threadIdx.x;
	uint32_t e; int mask;
	const uint32_t packed = cell_counts(g, density, thresh, idx, &e, &mask);
	uint32_t total;
	const uint32_t pre = block_scan_packed(packed & 0xffffu, s_warp, &total);
	if (!e) return;
	uint32_t vidx = off + pre;
	vert_prefix[idx] = vidx;
	const uint32_t z = (uint32_t)(idx / g.res2);
	const uint32_t rem = (uint32_t)(idx - (uint64_t)z * g.res2);
	const uint32_t y = rem / g.rx, x = rem - y * g.rx;
	const float scale[3] = {(aabb.mx[0] - aabb.mn[0]) / (float)g.rx, (aabb.mx[1] - aabb.mn[1]) / (float)g.ry, (aabb.mx[2] - aabb.mn[2]) / (float)g.rz};
	const float f0 = density[idx];
	const uint64_t nb[3] = {idx + 1, idx + g.res1, idx + g.res2};
	#pragma unroll
	for (int a = 0; a < 3; ++a) {
		if (!((e >> a) & 1u)) continue;
		const float f1 = density[nb[a]];
		const float dt = (thresh - f0) / (f1 - f0);
		const float p[3] = {(float)x + (a == 0 ? dt : 0.f), (float)y + (a == 1 ? dt : 0.f), (float)z + (a == 2 ? dt : 0.f)};
		#pragma unroll
		for (int k = 0; k < 3; ++k) verts[(size_t)vidx * 3 + k] = fmaf(p[k], scale[k], aabb.mn[k]);
		++vidx;
	}
}

// vertex id of the edge along `axis` that starts at cell c (= vertidx_grid[c + res3*axis] - 1 in the reference)
__device__ __forceinline__ uint32_t edge_vertex(const McGeom& g, const float* __restrict__ density, float thresh, const uint32_t* __restrict__ vert_prefix,
                                                uint32_t x, uint32_t y, uint32_t z, int axis) {
	const uint64_t c = x + (uint64_t)y * g.res1 + (uint64_t)z * g.res2;
	uint32_t v = vert_prefix[c];
	if (axis > 0) {
		const uint32_t e = cell_edges(g, density, thresh, x, y, z, c, density[c]);
		v += __popc(e & ((1u << axis) - 1u));
	}
	return v;
}

__global__ void __launch_bounds__(MC_BLOCK)
k_mc_emit_faces(McGeom g, float thresh, const float* __restrict__ density, const uint32_t* __restrict__ block_offsets, uint32_t n_blocks,
                uint32_t n_indices_total, const uint32_t* __restrict__ vert_prefix, uint32_t* __restrict__ indices) {
	__shared__ uint32_t s_warp[33];
	const uint32_t off = block_offsets[2 * blockIdx.x + 1];
	const uint32_t next = blockIdx.x + 1 < n_blocks ? block_offsets[2 * blockIdx.x + 3] : n_indices_total;
	if (next == off) return;
	const uint64_t idx = (uint64_t)blockIdx.x * MC_BLOCK + threadIdx.x;
	uint32_t e; int mask;
	const uint32_t packed = cell_counts(g, density, thresh, idx, &e, &mask);
	uint32_t total;
	const uint32_t pre = block_scan_packed(packed >> 16, s_warp, &total);
	const uint32_t n = packed >> 16;
	if (!n) return;
	const uint32_t z = (uint32_t)(idx / g.res2);
	const uint32_t rem = (uint32_t)(idx - (uint64_t)z * g.res2);
	const uint32_t y = rem / g.rx, x = rem - y * g.rx;
	// cube edge id -> (dx, dy, dz, axis): local_edges[] of gen_faces (marching_cubes.cu:686-703), 2 bits per field
	// e0 x@(0,0,0) e1 y@(1,0,0) e2 x@(0,1,0) e3 y@(0,0,0) e4 x@(0,0,1) e5 y@(1,0,1) e6 x@(0,1,1) e7 y@(0,0,1) e8..e11 z@(0,0,0),(1,0,0),(1,1,0),(0,1,0)
	const uint64_t tri = c_mc_triangles[mask];
	for (uint32_t i = 0; i < n; ++i) {
		const int ed = (int)((tri >> (4 * i)) & 0xf);
		int dx, dy, dz, axis;
		if (ed < 8) { axis = ed & 1; dz = ed >> 2; dx = (ed & 3) == 1; dy = (ed & 3) == 2; }
		else { axis = 2; dz = 0; dx = (ed == 9 || ed == 10); dy = (ed == 10 || ed == 11); }
		indices[off + pre + i] = edge_vertex(g, density, thresh, vert_prefix, x + dx, y + dy, z + dz, axis);
	}
}

uint32_t marching_cubes_blocks(const uint32_t res[3]) {
	const uint64_t res3 = (uint64_t)res[0] * res[1] * res[2];
	return (uint32_t)((res3 + MC_BLOCK - 1) / MC_BLOCK);
}
static McGeom make_geom(const uint32_t res[3]) {
	McGeom g; g.rx = res[0]; g.ry = res[1]; g.rz = res[2];
	g.res1 = res[0]; g.res2 = (uint64_t)res[0] * res[1]; g.res3 = g.res2 * res[2];
	return g;
}

int launch_marching_cubes_count(cudaStream_t stream, const uint32_t res[3], float thresh, const float* density, uint32_t* block_counts, uint32_t* counts_dev) {
	const uint32_t nb = marching_cubes_blocks(res);
	if (!nb) return 0;
	k_mc_count<<<nb, MC_BLOCK, 0, stream>>>(make_geom(res), thresh, density, block_counts);
	k_mc_scan<<<1, 1024, 0, stream>>>(nb, block_counts, counts_dev);
	COUNT_LAUNCH(2);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}

int launch_marching_cubes_emit(cudaStream_t stream, const uint32_t res[3], const Aabb& aabb, float thresh, const float* density, const uint32_t* block_offsets,
                               uint32_t n_verts, uint32_t n_indices, uint32_t* vert_prefix, float* verts, uint32_t* indices) {
	const uint32_t nb = marching_cubes_blocks(res);
	if (!nb || (!n_verts && !n_indices)) return 0;
	const McGeom g = make_geom(res);
	k_mc_emit_verts<<<nb, MC_BLOCK, 0, stream>>>(g, aabb, thresh, density, block_offsets, nb, n_verts, vert_prefix, verts);
	k_mc_emit_faces<<<nb, MC_BLOCK, 0, stream>>>(g, thresh, density, block_offsets, nb, n_indices, vert_prefix, indices);
	COUNT_LAUNCH(2);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}

// ---- mesh attributes ---------------------------------------------------------------------------------------
// compute_mesh_1ring's normal accumulation (marching_cubes.cu:333-370): area-weighted face normals summed per vertex.
__global__ void __launch_bounds__(256)
k_mesh_vertex_normals(uint32_t n_tris, const uint32_t* __restrict__ indices, const float* __restrict__ verts, float* __restrict__ normals) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_tris) return;
	const uint32_t ia = indices[i * 3 + 0], ib = indices[i * 3 + 1], ic = indices[i * 3 + 2];
	const float pa[3] = {verts[ia * 3], verts[ia * 3 + 1], verts[ia * 3 + 2]};
	const float pb[3] = {verts[ib * 3], verts[ib * 3 + 1], verts[ib * 3 + 2]};
	const float pc[3] = {verts[ic * 3], verts[ic * 3 + 1], verts[ic * 3 + 2]};
	const float u[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]}, v[3] = {pa[0] - pc[0], pa[1] - pc[1], pa[2] - pc[2]};
	const float n[3] = {fmaf(u[1], v[2], -(u[2] * v[1])), fmaf(u[2], v[0], -(u[0] * v[2])), fmaf(u[0], v[1], -(u[1] * v[0]))};   // (pb-pa) x (pa-pc), not normalised
	const uint32_t id[3] = {ia, ib, ic};
	#pragma unroll
	for (int k = 0; k < 3; ++k) { atomicAdd(normals + id[k] * 3 + 0, n[0]); atomicAdd(normals + id[k] * 3 + 1, n[1]); atomicAdd(normals + id[k] * 3 + 2, n[2]); }
}
// generate_nerf_network_inputs_from_positions (testbed_nerf.cu:847-853): outward-pointing view direction
__global__ void __launch_bounds__(256)
k_mesh_color_inputs(uint32_t n, const float* __restrict__ verts, Aabb train_aabb, float* __restrict__ coords) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float p[3] = {verts[i * 3], verts[i * 3 + 1], verts[i * 3 + 2]};
	const float c[3] = {p[0] - 0.5f, p[1] - 0.5f, p[2] - 0.5f};
	float d[3];
	normalize3(c, d);
	float* o = coords + (size_t)i * 7;
	#pragma unroll
	for (int k = 0; k < 3; ++k) o[k] = (p[k] - train_aabb.mn[k]) / (train_aabb.mx[k] - train_aabb.mn[k]);
	o[3] = warp_dt(min_cone_stepsize());
	#pragma unroll
	for (int k = 0; k < 3; ++k) o[4 + k] = (d[k] + 1.0f) * 0.5f;
}
// extract_srgb_with_activation (testbed_nerf.cu:532-545): logistic, optional linear -> sRGB
__global__ void __launch_bounds__(256)
k_mesh_extract_colors(uint32_t n, const __half* __restrict__ out16, int from_linear, float* __restrict__ colors) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n * 3) return;
	const uint32_t e = i / 3, d = i - e * 3;
	float c = logistic_f(__half2float(out16[(size_t)e * 16 + d]));
	if (from_linear) c = linear_to_srgb(c);
	colors[i] = c;
}

int launch_mesh_normals(cudaStream_t stream, uint32_t n_verts, uint32_t n_tris, const uint32_t* indices, const float* verts, float* normals) {
	CUDA_CHECK(cudaMemsetAsync(normals, 0, (size_t)n_verts * 3 * sizeof(float), stream));
	if (n_tris) { k_mesh_vertex_normals<<<(n_tris + 255) / 256, 256, 0, stream>>>(n_tris, indices, verts, normals); COUNT_LAUNCH(1); }
	CUDA_CHECK(cudaGetLastError());
	return 0;
}
int launch_mesh_color_inputs(cudaStream_t stream, uint32_t n_verts, const float* verts, const Aabb& train_aabb, float* coords) {
	if (n_verts) { k_mesh_color_inputs<<<(n_verts + 255) / 256, 256, 0, stream>>>(n_verts, verts, train_aabb, coords); COUNT_LAUNCH(1); }
	CUDA_CHECK(cudaGetLastError());
	return 0;
}
int launch_mesh_extract_colors(cudaStream_t stream, uint32_t n_verts, const __half* out16, int from_linear, float* colors) {
	if (n_verts) { k_mesh_extract_colors<<<(n_verts * 3 + 255) / 256, 256, 0, stream>>>(n_verts, out16, from_linear, colors); COUNT_LAUNCH(1); }
	CUDA_CHECK(cudaGetLastError());
	return 0;
}
