// ngp_kernels.h -- launcher declarations for the ray-marching side (sampling, loss, occupancy grid, optimiser).
#pragma once
#include "ngp_common.cuh"
#include "lookback.cuh"

struct LossParamsDev {
	float loss_scale;            // LOSS_SCALE = 128 (testbed_nerf.cu)
	float background_color[3];
	int random_bg_color;
	int snap_to_pixel_centers;
	float mask_loss_weight;
	float ek_loss_weight;
	float cos_anneal_ratio;
};

int launch_generate_training_samples(cudaStream_t stream, uint32_t n_rays, uint32_t ray_offset, uint32_t ray_stride, uint32_t n_rays_global, uint32_t n_rays_total, const uint64_t rng_state[2], const DatasetDev& ds,
                                     const Aabb& bb, const GlobalMove& gmv, const uint8_t* bitfield, int snap, float cone_angle, uint32_t max_samples,
                                     uint32_t* numsteps_all, uint32_t* base, uint32_t* slot, uint32_t* ray_indices, float* rays, uint32_t* numsteps,
                                     float* coords, uint32_t* counters, void* windows, uint32_t* n_windows, float* ray_setup,
                                     uint32_t* ray_done /* nullable: per kept ray slot, set to 0xffffffff for the early-terminating first pass */);
size_t sampling_window_bytes_per_ray();

// counters: [0] = kept rays (device, from sample generation), out_counter = total compacted samples (un-clamped)
// Scratch of the by-samples loss (loss.cu): per marched sample the compositing state behind it, per compacted slot its sample and ray, per ray a 48-byte record
struct LossScratch { float4* prefix4; float* prefix1; uint32_t* slot_sample; uint32_t* slot_ray; float* ray_rec; };
int launch_compute_loss(cudaStream_t stream, uint32_t n_rays_local, uint32_t n_rays /* global: image mapping + normalisers */, uint32_t n_rays_total, const uint64_t rng_state[2], const DatasetDev& ds, const Aabb& bb,
                        const GlobalMove& gmv, const LossParamsDev& lp, uint32_t max_samples_compacted, const uint32_t* n_kept_ptr,
                        const __half* network_output, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords,
                        float* ray_tmp, uint32_t* scan_tmp, float* coords_compacted, __half* dloss_doutput,
                        float* loss_output, float* ek_loss_output, float* mask_loss_output, uint32_t* out_counter,
                        // all non-null: the loss by samples (k_loss_rays + k_loss_samples).  composite_in: {alpha, r, g, b} per sample as stored by
                        // launch_network_forward_rays; lookback desc: 64-bit word per 8 rays, never cleared; ticket: one zeroed word; epoch: 1 .. 2^30 - 1,
                        // different from the previous launch's on the same desc.  Otherwise: composite / scan / gradient kernels, one warp per ray.
                        const float4* composite_in, const LossScratch* scratch, unsigned long long* lookback_desc, uint32_t* lookback_ticket, uint32_t lookback_epoch);

int launch_adam_ema(cudaStream_t stream, uint64_t n, uint64_t n_matrix, float loss_scale, float lr, float beta1, float beta2, float eps, float l2_reg,
                    float ema_decay, uint32_t ema_step, float* grads, float* master, __half* params, float* m1, float* m2, uint32_t* steps, __half* ema,
                    float* bias_table, uint32_t bias_table_size, int zero_grads, uint32_t* const zero_ptrs[3], const uint32_t zero_counts[3]);
// table[k] = sqrt(1 - beta2^k) / (1 - beta1^k) for k in [first, last] (api.cu: compiled with FMA contraction like the reference's adam_step)
int launch_adam_bias_table(cudaStream_t stream, float beta1, float beta2, uint32_t first, uint32_t last, float* bias_table, uint32_t bias_table_size);

#define NEUS2_MAX_PEERS 8
#define OPTIMIZER_BLOCK_QUADS 256   // the sharded optimizer deals blocks of 256 quads (1024 parameters) round-robin over the ranks
uint32_t optimizer_block_owner(uint64_t param_index, uint32_t world);
// ---- cross-GPU ordering points of the data-parallel step over peer memory (peer_sync.cu) ----
#define PEER_PAYLOAD_WORDS 20   // {kept rays, samples, compacted, 0 | loss sums 4 | DeltaNetwork gradient sums 12}
struct PeerSlot { uint32_t payload[PEER_PAYLOAD_WORDS]; uint32_t pad[11]; uint32_t flag; };   // one 128-byte line; written by exactly one peer
struct PeerSyncBlock { PeerSlot exchange[2][NEUS2_MAX_PEERS]; PeerSlot barrier[NEUS2_MAX_PEERS]; };   // exchange slots alternate with the epoch's parity
struct PeerSyncPtrs { PeerSyncBlock* block[NEUS2_MAX_PEERS]; };   // indexed by RANK (every rank's block as mapped into this process)
struct PeerZero { uint32_t* p[3]; uint32_t n[3]; };
struct PeerXfer { PeerSyncPtrs peers; uint32_t epoch, rank, world; };   // world == 0: no exchange (passed by value to the kernels that can carry one)
int launch_peer_exchange(cudaStream_t stream, const PeerXfer& x, const uint32_t* counts_local, uint32_t* counts_global, float* loss_sums, float* delta_acc);   // null = not part of this call's payload
int launch_peer_barrier(cudaStream_t stream, const PeerXfer& x, const PeerZero& z);

// zero_grads: every gradient quad this rank consumed (its slice of EVERY rank's buffer) is cleared through the same peer mapping, touched quads only
int launch_adam_ema_p2p(cudaStream_t stream, uint64_t n, uint64_t n_matrix, uint32_t rank, uint32_t world, float loss_scale, float lr, float beta1, float beta2,
                        float eps, float l2_reg, float ema_decay, uint32_t ema_step, float* const* grads_peers, __half* const* params_peers,
                        __half* const* ema_peers, float* master, float* m1, float* m2, uint32_t* steps, float* bias_table, uint32_t bias_table_size, int zero_grads,
                        const PeerXfer* barrier = nullptr /* the kernel's last block runs this barrier over the peers and then clears `zero` */, const PeerZero* zero = nullptr,
                        uint32_t* done_counter = nullptr);

int launch_density_grid_samples(cudaStream_t stream, uint32_t n, const uint64_t rng_state[2], uint32_t step, const Aabb& bb, const float* grid_in,
                                uint32_t n_cascades, float thresh, float* positions, uint32_t* indices,
                                // non-null: the uniform samples (thresh < 0) of a one-cascade grid are written in cell order (density_grid.cu); desc: 128^3 / 256
                                // 64-bit words, never cleared; ticket: one zeroed word; epoch as for launch_compute_loss
                                unsigned long long* lookback_desc = nullptr, uint32_t* lookback_ticket = nullptr, uint32_t lookback_epoch = 0);
int launch_density_grid_splat_decay(cudaStream_t stream, uint32_t n_samples, const uint32_t* indices, const __half* density, uint32_t n_elements,
                                    float decay, float* grid_tmp, float* grid);
int launch_density_grid_bitfield(cudaStream_t stream, const float* grid, uint32_t n_cascades_used, float* mean_tmp, uint8_t* bitfield);
int launch_reduce_sum(cudaStream_t stream, const float* in, uint32_t n, float* out /* zeroed by the launcher */);

// ---- inference render (render.cu) ----
struct RenderLaunch {
	int32_t width, height;
	float focal[2], screen_center[2];
	float camera[12];
	uint32_t sample_index;
	int32_t snap_to_pixel_centers;
	float min_transmittance;
	float background_color[4];
	float exposure;
	int32_t to_srgb;
	int32_t tile_x0, tile_y0, tile_w, tile_h;
	int32_t row_block, row_stride;   // interleaved rows, see neus2_render_params
};
// queue: 32 B per tile pixel; counters: 4 x u32; ray_rgba / accumulate / out: float4 per tile pixel; ray_depth, depth_out (nullable): float;
// ray_samples (nullable): u32 per tile pixel = samples composited along the ray.
int launch_render(cudaStream_t stream, const GridMeta& gm, const ParamLayout& pl, const __half* params, float sdf_bias, const RenderLaunch& r,
                  const Aabb& render_aabb, const Aabb& train_aabb, const GlobalMove& gmv, const uint8_t* bitfield, float cone_angle, float cos_anneal_ratio,
                  void* queue, uint32_t* counters, float* ray_rgba, float* ray_depth, uint32_t* ray_samples, float* accumulate, float* out, float* depth_out);

// ---- SDF lattice + marching cubes (mesh.cu) ----
// SDF (with bias, half -> float) at lattice points x in [0,res0), y in [0,res1), z in [z0,z1): out[x + y*res0 + (z-z0)*res0*res1].
int launch_sdf_lattice(cudaStream_t stream, const GridMeta& gm, const ParamLayout& pl, const __half* params, float sdf_bias, const uint32_t res[3],
                       const Aabb& render_aabb, const Aabb& train_aabb, uint32_t z0, uint32_t z1, float* out);
// Deterministic marching cubes.  Pass 1 (count): fills block_counts (2 x n_blocks u32) and counts_dev[0..1] = {vertices, indices}.
// Pass 2 (emit): needs vert_prefix (u32 per cell), verts (3 floats per vertex) and indices.
uint32_t marching_cubes_blocks(const uint32_t res[3]);
int launch_marching_cubes_count(cudaStream_t stream, const uint32_t res[3], float thresh, const float* density, uint32_t* block_counts, uint32_t* counts_dev);
int launch_marching_cubes_emit(cudaStream_t stream, const uint32_t res[3], const Aabb& aabb, float thresh, const float* density, const uint32_t* block_offsets,
                               uint32_t n_verts, uint32_t n_indices, uint32_t* vert_prefix, float* verts, uint32_t* indices);
int launch_mesh_normals(cudaStream_t stream, uint32_t n_verts, uint32_t n_tris, const uint32_t* indices, const float* verts, float* normals);
int launch_mesh_color_inputs(cudaStream_t stream, uint32_t n_verts, const float* verts, const Aabb& train_aabb, float* coords);
int launch_mesh_extract_colors(cudaStream_t stream, uint32_t n_verts, const __half* out16, int from_linear, float* colors);

// first network pass of the training step, ray by ray with the loss kernel's early termination (forward_rays.cu)
int launch_network_forward_rays(cudaStream_t stream, const GridMeta& gm, const ParamLayout& pl, const __half* params, float sdf_bias, float cos_anneal_ratio,
                                const uint32_t* n_kept_ptr, const uint32_t* longest_ptr, const uint32_t* numsteps, const float* coords, uint32_t max_samples,
                                uint32_t n_rays_max, __half* out, float4* composite_in /* nullable: {alpha, r, g, b} per evaluated sample, for launch_compute_loss */,
                                uint32_t* ray_done, uint32_t* work_counter);
