// ngp_kernels.h -- launcher declarations for the ray-marching side (sampling, loss, occupancy grid, optimiser).
#pragma once
#include "ngp_common.cuh"

struct LossParamsDev {
	float loss_scale;            // LOSS_SCALE = 128 (testbed_nerf.cu)
	float background_color[3];
	int random_bg_color;
	int snap_to_pixel_centers;
	float mask_loss_weight;
	float ek_loss_weight;
	float cos_anneal_ratio;
};

int launch_generate_training_samples(cudaStream_t stream, uint32_t n_rays, uint32_t ray_offset, uint32_t n_rays_global, uint32_t n_rays_total, const uint64_t rng_state[2], const DatasetDev& ds,
                                     const Aabb& bb, const GlobalMove& gmv, const uint8_t* bitfield, int snap, float cone_angle, uint32_t max_samples,
                                     uint32_t* numsteps_all, uint32_t* base, uint32_t* slot, uint32_t* ray_indices, float* rays, uint32_t* numsteps,
                                     float* coords, uint32_t* counters, void* windows, uint32_t* n_windows, float* ray_setup);
size_t sampling_window_bytes_per_ray();

// counters: [0] = kept rays (device, from sample generation), out_counter = total compacted samples (un-clamped)
int launch_compute_loss(cudaStream_t stream, uint32_t n_rays_local, uint32_t n_rays /* global: image mapping + normalisers */, uint32_t n_rays_total, const uint64_t rng_state[2], const DatasetDev& ds, const Aabb& bb,
                        const GlobalMove& gmv, const LossParamsDev& lp, uint32_t max_samples_compacted, const uint32_t* n_kept_ptr,
                        const __half* network_output, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords,
                        float* ray_tmp, uint32_t* scan_tmp, float* coords_compacted, __half* dloss_doutput,
                        float* loss_output, float* ek_loss_output, float* mask_loss_output, uint32_t* out_counter);

int launch_adam_ema(cudaStream_t stream, uint64_t n, uint64_t n_matrix, float loss_scale, float lr, float beta1, float beta2, float eps, float l2_reg,
                    float ema_decay, uint32_t ema_step, const float* grads, float* master, __half* params, float* m1, float* m2, uint32_t* steps, __half* ema);

int launch_density_grid_samples(cudaStream_t stream, uint32_t n, const uint64_t rng_state[2], uint32_t step, const Aabb& bb, const float* grid_in,
                                uint32_t n_cascades, float thresh, float* positions, uint32_t* indices);
int launch_density_grid_splat_decay(cudaStream_t stream, uint32_t n_samples, const uint32_t* indices, const __half* density, uint32_t n_elements,
                                    float decay, float* grid_tmp, float* grid);
int launch_density_grid_bitfield(cudaStream_t stream, const float* grid, uint32_t n_cascades_used, float* mean_tmp, uint8_t* bitfield);
int launch_reduce_sum(cudaStream_t stream, const float* in, uint32_t n, float* out /* zeroed by the launcher */);
