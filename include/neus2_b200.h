/* neus2_b200.h -- C ABI of the B200-native NeuS2 training / rendering hot path.
 *
 * Drop-in boundary for the path the reference drives through
 *   tcnn::DifferentiableObject<float, half, half> as implemented by NerfNetwork<half>
 *   (dependencies/my_tcnn/include/tiny-cuda-nn/object.h:96-303, include/neural-graphics-primitives/nerf_network.h:100-145,330)
 * and its callers in Testbed (src/testbed_nerf.cu:3293-3397, 3440-4001; src/testbed.cu:2640-2740).
 *
 * Conventions: every function returns 0 on success, non-zero on failure (message via neus2_last_error());
 * nothing throws across the ABI.  Pointers named *_dev are CUDA device pointers on the context's device,
 * *_host are host pointers.  All work is enqueued on the caller's stream (a cudaStream_t passed as void*,
 * NULL = default stream); functions that return host values synchronise that stream themselves.
 * Matrices use the reference's layouts: network input = NerfCoordinate AoS, 7 floats (nerf.h:76-102);
 * network output = 16 halves per sample AoS (nerf_network.h:935; [0..2] rgb logits, [3] sdf, [4..6] d sdf/dx,
 * [7] variance, [8..10] warped view dir).  Parameters live in ONE contiguous buffer in the reference order
 * [sdf-MLP | colour-MLP | hash grid | variance(4)] (nerf_network.h:741-785).
 */
#ifndef NEUS2_B200_H
#define NEUS2_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct neus2_ctx neus2_ctx;

/* Mirrors the keys of configs/nerf/base.json that the hot path consumes (src/testbed.cu:2084-2199). */
typedef struct neus2_config {
	/* "encoding" */
	uint32_t n_levels;               /* 14 */
	uint32_t log2_hashmap_size;      /* 19 */
	uint32_t base_resolution;        /* 16 */
	float top_resolution;            /* 2048 */
	float per_level_scale;           /* 0 = derive: exp(log(top*aabb/base)/(L-1)), testbed.cu:2175-2185 */
	float valid_level_scale;         /* 0.02  progressive levels, grid.h:2427-2440 */
	float base_valid_level_scale;    /* 0.2 */
	uint32_t base_training_step;     /* 100 */
	/* "network" / "rgb_network": 64 wide, ReLU, n_hidden_layers 1 / 2 (fixed shape of this build) */
	float sdf_bias;                  /* -0.1, nerf_network.h:87 */
	/* "optimizer" (Ema{ExponentialDecay{Adam}}) */
	float learning_rate;             /* 1e-3 */
	float beta1, beta2, epsilon;     /* 0.9, 0.99, 1e-15 */
	float l2_reg;                    /* 1e-6 */
	float ema_decay;                 /* 0.95 */
	/* "hyperparams" */
	float mask_loss_weight;          /* 0.0 */
	float ek_loss_weight;            /* 0.01 */
	int32_t anneal_end;              /* 0 -> cos_anneal_ratio = 1 */
	uint32_t batch_size;             /* 1<<18 (testbed.cu:2120) */
	/* dataset / testbed */
	int32_t aabb_scale;              /* 1 */
	uint32_t seed;                   /* 1337 (testbed.h:524) */
	int32_t random_bg_color;         /* 1 (nerf.training.random_bg_color) */
	int32_t snap_to_pixel_centers;   /* 1 (testbed.h:623) */
	float background_color[3];       /* used when random_bg_color == 0 */
	float density_grid_decay;        /* 0.95 (testbed.h:644) */
	float loss_scale;                /* 128 (LOSS_SCALE) */
	/* data-parallel build-out (not in the reference, which is single-GPU): buffers are sized for this rank's share */
	uint32_t world_size;             /* 1 */
	uint32_t max_rays_per_batch;     /* 1<<18 (testbed_nerf.cu:3435); raise to world_size<<18 for weak scaling */
	/* 1 (default): the step's first network pass stops along each ray where the loss stops reading (T < 1e-4, testbed_nerf.cu:1567);
	 * 0: evaluate every marched sample like the reference.  Same loss, gradients and counters either way (tests/test_gpu_step.py).
	 * With 1 the first pass also leaves {alpha, r, g, b} per evaluated sample and the loss runs by samples (loss.cu: a light per-ray replay
	 * of the compositing recurrences, then one thread per kept sample); with 0 the loss kernels work one warp per ray.  Same bits. */
	int32_t forward_early_termination;
	/* 1 (default, single-GPU contexts): pipelined step.  The step's only host read-back (the sample counts, final after the loss kernel)
	 * is taken BEFORE the backward pass is queued; while the GPU runs backward pass + optimizer the host computes the next step's
	 * rays_per_batch and enqueues its sample generation on a second stream behind the fused forward+backward kernel, so it overlaps the
	 * weight-gradient GEMMs and Adam, and the GPU never waits for the host between steps.  Same arithmetic, same order of random
	 * numbers, same results (tests/test_gpu_step.py); the sample buffers then hold the NEXT step's rays after a step returns.
	 * 0: the reference's order (generate, forward, loss, backward, optimizer, read back). */
	int32_t prefetch_next_samples;
	/* "globalmove" optimizer (configs/nerf/base.json:135-158: Ema{ExponentialDecay{Adam}}) of the DeltaNetwork's own trainer
	 * (src/testbed.cu:2320-2328).  The Ema wrapper's inference copy has no consumer (the network always reads params()), so it is not kept. */
	float delta_learning_rate;       /* 1e-4 */
	float delta_beta1, delta_beta2;  /* 0.9, 0.99 */
	float delta_epsilon;             /* 1e-10 */
	uint32_t delta_decay_start;      /* 50 */
	uint32_t delta_decay_interval;   /* 25 (0 = no decay) */
	float delta_decay_base;          /* 0.99 */
	/* ExponentialDecay around the main Adam (configs/nerf/base.json:7-12; tcnn optimizers/exponential_decay.h:60-71): from optimizer step
	 * decay_start on, every decay_interval steps (up to decay_end), the learning rate is multiplied by decay_base. */
	uint32_t decay_start;            /* 20000 */
	uint32_t decay_interval;         /* 10000 (0 = no decay) */
	uint32_t decay_end;              /* 10000000 (the reference's default; base.json does not set it) */
	float decay_base;                /* 0.33 */
	/* "reset_density_grid_after_global_movement" (src/testbed.cu:2123, 2664-2666; reset_density_grid_nerf, src/testbed_nerf.cu:3205-3291): on
	 * frames >= 1, at the step where the canonical model starts training again (training_step == frozen_steps), the occupancy grid is rebuilt
	 * from zero with one uniform pass over every cell.  The reference's code default is true, configs/nerf/base.json:130 sets false (the default here). */
	int32_t reset_density_grid_after_global_movement;
} neus2_config;

typedef struct neus2_step_stats {
	uint32_t training_step;              /* after the step */
	uint32_t rays_per_batch;             /* rays traced this step (global) */
	uint32_t rays_per_batch_next;        /* adapted value for the next step (testbed_nerf.cu:3434-3435) */
	uint32_t n_rays_kept;                /* rays with >0 samples that fit */
	uint32_t measured_batch_size_before_compaction;  /* N_u */
	uint32_t measured_batch_size;        /* compacted sample count (un-clamped) */
	float loss, ek_loss, mask_loss;      /* valid when loss_valid (refreshed every 16 steps, testbed.cu:2714) */
	int32_t loss_valid;
	int32_t density_grid_updated;
	uint32_t first_pass_samples;         /* samples this rank's first network pass evaluated (<= N_u with forward_early_termination) */
	int32_t samples_prefetched;          /* 1: this step's samples had been generated during the previous step (prefetch_next_samples) */
} neus2_step_stats;

const char* neus2_last_error(void);
void neus2_default_config(neus2_config* cfg);                       /* base.json values */
/* m_learning_rate_factor of ExponentialDecayOptimizer (exponential_decay.h:60-71) when the nested Adam has done `optimizer_steps_done` steps and
 * is about to do the next one: decay_base multiplied in once per trigger, in float like the reference.  Pure host arithmetic (no GPU needed);
 * the optimizer entry points use it, so the factor needs no state of its own and survives snapshots through the optimizer step count. */
float neus2_learning_rate_factor(const neus2_config* cfg, uint32_t optimizer_steps_done);
/* rays_per_batch of the next step, Counters::update_after_training (src/testbed_nerf.cu:3435-3436).  Pure host arithmetic; the step uses it. */
uint32_t neus2_next_rays_per_batch(uint32_t rays_per_batch, uint32_t target_batch_size, uint32_t measured_batch_size, uint32_t world_size, uint32_t max_rays_per_batch);

/* Testbed::reset_network (src/testbed.cu:2084) for TestbedMode::Nerf: builds tables, allocates every buffer. */
int neus2_ctx_create(const neus2_config* cfg, int device, neus2_ctx** out);
void neus2_ctx_destroy(neus2_ctx* ctx);

/* ParametricObject::n_params / layer sizes (object.h:60-94).  layout[8] = {sdf_in_width, off_W0, off_W1, off_C0, off_C1, off_C2, off_grid, off_variance}. */
uint64_t neus2_n_params(const neus2_ctx* ctx);
int neus2_param_layout(const neus2_ctx* ctx, uint32_t layout[8]);
/* Grid level tables (grid.h:1472-1508): offsets[n_levels+1], resolutions[n_levels], scales[n_levels]. */
int neus2_grid_tables(const neus2_ctx* ctx, uint32_t* offsets, uint32_t* resolutions, float* scales);
/* Trainer-owned buffers (trainer.h:72-109): fp32 master, fp16 params, fp16 EMA ("inference") params, gradients (fp32 here). */
int neus2_buffers(neus2_ctx* ctx, float** master_dev, void** params_f16_dev, void** ema_f16_dev, float** grads_f32_dev);

/* Trainer::initialize_params (trainer.h:70-108) + NerfNetwork::initialize_params (nerf_network.h:815-905): same RNG
 * stream as the reference.  sdf_mlp_weights_host: contents of utils/mlp_weights_hidden_layer_num_1_hidden_size_32.txt
 * (3072 floats, row-major [out][in], first layer then last; nerf_network.h:787-813) or NULL to generate the geometric
 * init (dependencies/my_tcnn/scripts/geometry_init_save_weights.py:292-389). */
int neus2_init_params(neus2_ctx* ctx, const float* sdf_mlp_weights_host);
/* Trainer::set_params_full_precision: upload n fp32 params, cast to fp16; reset=1 also zeroes Adam/EMA state. */
int neus2_set_params(neus2_ctx* ctx, const float* master_host, uint64_t n, int reset_optimizer);
int neus2_get_params(neus2_ctx* ctx, float* master_host, uint64_t n);

/* Encoding::set_training_step (grid.h:2427-2440). */
int neus2_set_training_step(neus2_ctx* ctx, int32_t step);

/* NerfNetwork::inference_mixed_precision_impl (nerf_network.h:100-103): coords [n][7] float -> out [n][16] half. */
int neus2_forward(neus2_ctx* ctx, void* stream, const float* coords_dev, uint32_t n, void* out_f16_dev, int use_inference_params);
/* NerfNetwork::sdf / ::density (nerf_network.h:656-739): pos [n][stride] float -> half [n]. */
int neus2_sdf(neus2_ctx* ctx, void* stream, const float* pos_dev, uint32_t stride, uint32_t n, void* out_f16_dev, int as_density, int use_inference_params);
/* NerfNetwork::forward + ::backward (nerf_network.h:145,330) with EGradientMode::Overwrite on the compacted batch:
 * rows [0,n_rows) read source row i % n_in (fill_rollover, testbed_nerf.cu:3922-3930); gradients land in grads_f32. */
int neus2_backward(neus2_ctx* ctx, void* stream, const float* coords_dev, const void* dL_dout_f16_dev, const uint32_t* n_in_dev,
                   uint32_t n_rows, uint32_t indeed_batch_size);
/* Trainer::optimizer_step (trainer.h:170): fused Adam + EMA (adam.h:52-160, ema.h:64-77). */
int neus2_optimizer_step(neus2_ctx* ctx, void* stream, float loss_scale);

/* Training data (NerfDataset, nerf_loader.h): n RGBA8 images of one resolution, per-image focal length (pixels),
 * principal point (relative), camera-to-world 3x4 row-major in ngp convention (nerf_matrix_to_ngp already applied). */
int neus2_set_dataset(neus2_ctx* ctx, uint32_t n_images, int32_t width, int32_t height, const uint8_t* pixels_host,
                      const float* focal_host, const float* principal_host, const float* xforms_host);

/* Testbed::update_density_grid_nerf via training_prep_nerf (testbed_nerf.cu:3293-3397, 4003-4016). */
int neus2_update_density_grid(neus2_ctx* ctx, void* stream);

/* Testbed::train(batch) for one step (src/testbed.cu:2640-2740 -> testbed_nerf.cu:3440-4001), split so that a
 * data-parallel caller can all-reduce grads_f32 between the halves:
 *   begin: progressive level, occupancy refresh, sample generation for the global rays rank, rank + world_size, ... (n_local of them)
 *          of this step's rays_per_batch, forward, loss, backward on n_rows_local compacted rows.
 *   end:   optimizer step, counters read-back, rays_per_batch adaptation, rng advance.
 * world_size = 1: ray_offset = 0, n_local = 0 (= all rays), n_rows_local = 0 (= batch_size). */
int neus2_train_step_begin(neus2_ctx* ctx, void* stream, uint32_t rank, uint32_t world_size);
/* counts_global: {kept rays, N_u, compacted} summed over ranks, or NULL = take them from the device buffer returned by
 * neus2_counts_device (begin() leaves this rank's counts there; a data-parallel caller all-reduces the 4 x uint32 in place,
 * stream-ordered, between begin and end -- no host round trip; with world_size 1 they are simply the local counts). */
int neus2_train_step_end(neus2_ctx* ctx, void* stream, const uint32_t* counts_global_host, neus2_step_stats* stats);
/* Data-parallel pipelining (config.prefetch_next_samples): call after the all-reduce of neus2_counts_device() has been enqueued on
 * `stream`.  Takes the step's host read-back there (ahead of the optimizer and its collectives); neus2_train_step_end then does not
 * synchronise at the end of the step and generates the next step's samples on a second stream beside the optimizer. */
int neus2_train_step_counts_ready(neus2_ctx* ctx, void* stream);
int neus2_local_counts(neus2_ctx* ctx, void* stream, uint32_t counts_host[3]);
int neus2_counts_device(neus2_ctx* ctx, uint32_t** counts_dev);
/* The step's loss sums {rgb, eikonal, mask, -} over THIS rank's rays (fp32, device), reduced right behind the loss kernel on steps
 * with training_step % 16 == 0 (*due = 1): a data-parallel caller all-reduces them in place before neus2_train_step_counts_ready /
 * neus2_train_step_end reads them back; otherwise the reported losses are this rank's share. */
int neus2_loss_scalars_device(neus2_ctx* ctx, float** sums_dev, int* due);
/* convenience: begin + end on one GPU */
int neus2_train_step(neus2_ctx* ctx, void* stream, neus2_step_stats* stats);

/* Test / integration access to the step's device buffers by name: "numsteps_all", "ray_indices", "rays", "numsteps",
 * "coords", "mlp_out", "coords_compacted", "dloss", "loss", "ek_loss", "mask_loss", "counters", "density_grid",
 * "bitfield", "density_grid_mean", "render_queue", "render_counters", "render_ray_rgba", "render_ray_depth", "render_ray_samples". */
int neus2_debug_buffer(neus2_ctx* ctx, const char* name, void** ptr_dev, uint64_t* n_bytes);
/* Testbed state the drivers read (python_api.cu:317-471): training_step, rays_per_batch, rng state. */
int neus2_get_state(neus2_ctx* ctx, uint32_t* training_step, uint32_t* rays_per_batch, uint64_t rng_state[2]);
int neus2_set_state(neus2_ctx* ctx, uint32_t training_step, uint32_t rays_per_batch, uint32_t n_rays_total, uint32_t measured_before_compaction);
int neus2_set_bitfield(neus2_ctx* ctx, const uint8_t* bitfield_host, uint64_t n_bytes);
/* Per-stage device timing (CUDA events on the caller's stream): stages = {occupancy grid, sample generation, forward,
 * loss, fused forward+backward, weight-gradient GEMMs, optimizer, whole step}; sums in ms and number of samples. */
int neus2_set_profiling(neus2_ctx* ctx, int enable);
int neus2_get_stage_times(neus2_ctx* ctx, float ms_sum[8], uint32_t count[8]);
/* The same 8 stages followed by the data-parallel sync points: [8] counts exchange (behind the loss kernel), [9] gradient-ready exchange
 * (behind the backward pass), [10] barrier behind the sharded optimizer, [11] unused.  Their time is almost entirely WAITING for the
 * slowest peer. */
int neus2_get_stage_times_ext(neus2_ctx* ctx, float ms_sum[12], uint32_t count[12]);
/* Profiling only (process-wide; results are WRONG with any bit set): switches parts of the training kernel off so that their cost can be
 * timed on a frozen state (tools/bench_fwdbwd.py).  1 = no hash-grid atomics, 4 = no read-back of the tensor-memory accumulators,
 * 8 = weight-gradient operands staged but never multiplied, 16 = no waits for tcgen05 completion.  0 = normal operation. */
int neus2_set_debug_bits(int bits);
/* number of kernels this library launched on the context since creation (bench.py's gpu_launches) */
uint64_t neus2_launch_count(const neus2_ctx* ctx);

/* ---- inference rendering ------------------------------------------------------------------------------
 * One call = one Testbed::render_frame (src/testbed.cu:2820-2990) in Nerf mode, ERenderMode::Shade, for one sample index of
 * one image tile: NerfTracer::init_rays_from_camera + ::trace (src/testbed_nerf.cu:2397-2600), shade_kernel_nerf (:2148-2181),
 * CudaRenderBuffer::accumulate / ::tonemap (src/render_buffer.cu:540-598).  Testbed::render_to_cpu (src/python_api.cu:123-170)
 * is the loop `for s in range(spp): neus2_render(sample_index = s)`; sample 0 resets the accumulation. */
typedef struct neus2_render_params {
	int32_t width, height;           /* full image resolution */
	float focal_length[2];           /* pixels: relative_focal_length * resolution[fov_axis] * zoom (testbed.cu:2738) */
	float screen_center[2];          /* Testbed::render_screen_center (testbed.cu:2742); 0.5, 0.5 = centred */
	float camera_matrix[12];         /* 3x4 row-major camera-to-world, ngp convention (m_smoothed_camera) */
	uint32_t sample_index;           /* render_buffer.spp() */
	int32_t snap_to_pixel_centers;   /* testbed.snap_to_pixel_centers */
	float min_transmittance;         /* nerf.rendering_min_transmittance (0.01; the eval scripts set 1e-4) */
	float background_color[4];       /* sRGB + alpha (testbed.h:842) */
	float exposure;                  /* 0 */
	int32_t to_srgb;                 /* render(linear=False) */
	int32_t use_inference_params;    /* 1 = EMA weights, as the reference renders */
	int32_t tile_x0, tile_y0, tile_w, tile_h;   /* 0, 0, width, height = whole image; tiles of one image are independent */
	/* interleaved rows (one call renders a GPU's share of a frame that is dealt in row blocks): tile row ry is image row
	 * tile_y0 + (ry / row_block) * row_block * row_stride + ry % row_block.  row_block = 0: contiguous rows. */
	int32_t row_block, row_stride;
} neus2_render_params;
void neus2_default_render_params(neus2_render_params* p, int32_t width, int32_t height);
/* rgba_out_dev: [tile_h][tile_w][4] float (the surface render_to_cpu copies out); depth_out_dev: [tile_h][tile_w] float or NULL. */
int neus2_render(neus2_ctx* ctx, void* stream, const neus2_render_params* p, float* rgba_out_dev, float* depth_out_dev);
/* after a render on `stream`: {rays that entered the occupancy grid, samples pushed through the network} (synchronises) */
int neus2_render_stats(neus2_ctx* ctx, void* stream, uint32_t stats_host[2]);

/* ---- SDF lattice + marching cubes ------------------------------------------------------------------------
 * Testbed::get_density_on_grid (src/testbed_nerf.cu:4096-4143): SDF (inference params) at lattice points x in [0,res[0]),
 * y in [0,res[1]), z in [z0,z1) of `aabb` (6 floats min,max; Testbed::marching_cubes rounds res up to multiples of 16 itself,
 * :4176-4178).  out_dev[x + y*res0 + (z-z0)*res0*res1].  z-slabs are independent (one per GPU). */
int neus2_sdf_lattice(neus2_ctx* ctx, void* stream, const uint32_t res[3], const float aabb[6], uint32_t z0, uint32_t z1, float* out_dev, int use_inference_params);
/* marching_cubes_gpu (src/marching_cubes.cu:797-822) on a float lattice (inside = density > thresh).  Vertices are numbered in
 * ascending (cell, axis) order and triangles in ascending cell order (the reference: atomicAdd arrival order).  Results stay
 * in context-owned device buffers until the next call; synchronises `stream` once to size them, like the reference. */
int neus2_marching_cubes(neus2_ctx* ctx, void* stream, const uint32_t res[3], const float aabb[6], float thresh, const float* density_dev,
                         uint32_t* n_verts, uint32_t* n_indices);
int neus2_mesh_buffers(neus2_ctx* ctx, float** verts_dev, uint32_t** indices_dev, uint32_t* n_verts, uint32_t* n_indices);
/* compute_mesh_1ring's area-weighted vertex normals (src/marching_cubes.cu:333-370) and compute_mesh_vertex_colors
 * (src/testbed_nerf.cu:4071-4093: network colour seen from outside, logistic, sRGB) of that mesh; [n_verts][3] floats each, either NULL. */
int neus2_mesh_attributes(neus2_ctx* ctx, void* stream, float* normals_out_dev, float* colors_out_dev, int use_inference_params);

/* ---- data-parallel training over one NVSwitch domain (not in the reference, which is single-GPU) -----------------
 * The optimizer step fused with its collective: rank r owns every W-th block of 1024 parameters (1/W of them), sums those of every
 * rank's gradient buffer through NVLink peer loads, applies Adam + EMA to it and stores the new fp16 parameters / EMA values
 * into every rank's buffers.  Set-up: each rank exports CUDA IPC handles of its {gradient, fp16 parameter, EMA} buffers, the
 * host side exchanges them (torch.distributed all_gather), every rank attaches all of them.  A fourth mapped buffer is a small
 * per-rank sync block (flags + payload slots written by the peers), which carries the step's two cross-rank ordering points
 * without a collective library call.  Per step:
 *   neus2_train_step_begin -> neus2_peer_exchange (all backward passes done; counts / loss sums / DeltaNetwork gradient sums
 *   summed over ranks, in rank order) -> neus2_train_step_counts_ready -> neus2_optimizer_step_sharded -> neus2_peer_barrier
 *   (all slices stored everywhere) -> neus2_train_step_end (which then skips its own optimizer call).
 * A caller without peer mappings all-reduces neus2_counts_device (+ neus2_loss_scalars_device, neus2_delta_gradients_device and
 * the gradient buffer) with its collective library instead and lets neus2_train_step_end run the replicated optimizer.
 * The fp32 master copy / Adam moments of a rank are current only for its own slice (neus2_optimizer_shard): gather the slices
 * before serialising optimizer state (neus2_b200.parallel.sync_optimizer_state). */
#define NEUS2_IPC_BUFFERS 4
#define NEUS2_IPC_HANDLE_BYTES 64
int neus2_ipc_export(neus2_ctx* ctx, unsigned char* handles /* NEUS2_IPC_BUFFERS x NEUS2_IPC_HANDLE_BYTES */);
int neus2_ipc_attach(neus2_ctx* ctx, uint32_t rank, uint32_t world_size, const unsigned char* all_handles /* [world_size][NEUS2_IPC_BUFFERS][64] */);
/* Ownership of the sharded optimizer: blocks of *block_params consecutive parameters are dealt round-robin, parameter i belongs to rank
 * (i / *block_params) % world_size; *n_owned = how many parameters that gives `rank`. */
int neus2_optimizer_shard(const neus2_ctx* ctx, uint32_t rank, uint32_t world_size, uint64_t* block_params, uint64_t* n_owned);
int neus2_optimizer_step_sharded(neus2_ctx* ctx, void* stream, float loss_scale);
int neus2_peer_exchange(neus2_ctx* ctx, void* stream);
int neus2_peer_barrier(neus2_ctx* ctx, void* stream);

/* ---- snapshot support: the state Testbed::save_snapshot serialises (src/testbed.cu:3144-3178; tcnn trainer.h:281-304,
 * optimizers/adam.h:424-445, ema.h:182-194) beyond what neus2_buffers / neus2_get_state / neus2_debug_buffer already expose. */
int neus2_optimizer_buffers(neus2_ctx* ctx, float** first_moments_dev, float** second_moments_dev, uint32_t** param_steps_dev, uint32_t* current_step);
int neus2_set_optimizer_step(neus2_ctx* ctx, uint32_t current_step);
/* {n_rays_total, measured_batch_size_before_compaction, density_grid_ema_step, optimizer current_step, measured_batch_size,
 * m_canonical_training_step} (testbed.h:598-612, 640, 877) */
int neus2_get_counters(neus2_ctx* ctx, uint32_t out[6]);
/* load_snapshot's density-grid restore (testbed.cu:3222-3240): uploads the grid and rebuilds mean + occupancy bitfield. */
int neus2_set_density_grid(neus2_ctx* ctx, void* stream, const float* density_grid_host, uint64_t n, uint32_t ema_step);

/* ---- sequences (the part that needs no global-movement network) -------------------------------------------
 * Testbed::training_network_next_frame (src/testbed.cu:2001-2082) with "predict_global_movement": false: call after
 * neus2_set_dataset installed the next frame's images.  The model continues from its inference (EMA) weights with a fresh
 * optimizer, step counters at zero, rays_per_batch 4096, m_rng re-seeded; the parameters stay frozen for the first
 * `frozen_steps` steps (predict_global_movement_training_step) and the learning rate becomes `after_learning_rate` (> 0). */
int neus2_next_frame(neus2_ctx* ctx, void* stream, float after_learning_rate, uint32_t frozen_steps);
int neus2_get_time_frame(const neus2_ctx* ctx, uint32_t* current_training_time_frame, float* learning_rate);
/* m_canonical_training_step BEFORE the step with index `training_step` of time frame `time_frame` runs (src/testbed_nerf.cu:3512-3527;
 * reset by training_network_next_frame, src/testbed.cu:2015, 2050); frozen_steps = predict_global_movement_training_step (0 without
 * global-movement prediction).  Host arithmetic.  It drives the occupancy-grid cadence (testbed.cu:2669-2670), the uniform-sampling
 * phase (testbed_nerf.cu:4011) and the restart of the image round-robin (testbed_nerf.cu:3785). */
uint32_t neus2_canonical_training_step(uint32_t time_frame, uint32_t frozen_steps, uint32_t training_step);

/* "incremental_reinit_sdf_mlp" (src/testbed.cu:2065-2068; NerfNetwork::initialize_sdf_mlp_params, nerf_network.h:888-908): every
 * incremental_reinit_sdf_mlp_iters frames the fp32 master copy of the SDF MLP goes back to its geometric initialisation
 * (sdf_mlp_weights: 64*32 + 16*64 floats as for neus2_init_params, or NULL for the built-in generator).  Call after neus2_next_frame. */
int neus2_reinit_sdf_mlp(neus2_ctx* ctx, void* stream, const float* sdf_mlp_weights);

/* ---- global rigid motion of dynamic scenes ("predict_global_movement": true) -------------------------------
 * The reference keeps an ACCUMULATED rotation (3x3) + transition applied to every ray (global_movement_with_rotation_6d,
 * src/testbed_nerf.cu:194-213, 1377-1387, 2287-2294) and, for frames >= 1, a DeltaNetwork (include/neural-graphics-primitives/
 * transform_network.h:23-480): 6d rotation + transition applied to every SAMPLE inside the network (add_global_movement_with_rotation_6d,
 * common_operation.cuh:416-470; nerf_network.h:162-181, 663-674) and trained by its own Adam from the position gradients
 * (nerf_network.h:604-651; transform_network.h:110-250).
 *
 * neus2_begin_global_movement: Testbed::training_network_next_frame's global-movement part (src/testbed.cu:2054, 2076-2078 and the
 * trainer reset at :2569-2577): accumulated <- delta o accumulated (accumulate_global_movement_rotation_6d_kernel,
 * common_operation.cuh:551-580), delta <- identity with a fresh optimizer, m_use_delta = m_train_delta = 1.  Call it next to
 * neus2_next_frame.  While the mode is on, every entry point that evaluates the network (train step, forward, sdf, render,
 * sdf_lattice, mesh attributes, density-grid refresh) deforms its sample positions / directions first, the backward pass also
 * produces the 9 delta gradients and neus2_optimizer_step / _sharded also steps the delta parameters.
 * neus2_set_delta_mode(0): "finetune_global_movement": false at step predict_global_movement_training_step (src/testbed.cu:2659-2663). */
int neus2_begin_global_movement(neus2_ctx* ctx, void* stream);
int neus2_set_delta_mode(neus2_ctx* ctx, int use_delta);
/* save_global_movement / load_global_movement / load_local_movement (nerf_network.h:1180-1278), Testbed::save_transform
 * (src/testbed.cu:3119-3142).  Null pointers are skipped. */
int neus2_get_global_movement(neus2_ctx* ctx, void* stream, float accumulated_rotation[9], float accumulated_transition[3],
                              float delta_rotation_6d[6], float delta_transition[3], int* use_delta);
int neus2_set_global_movement(neus2_ctx* ctx, void* stream, const float accumulated_rotation[9], const float accumulated_transition[3],
                              const float delta_rotation_6d[6], const float delta_transition[3], const float first_frame_offset[3]);
/* Device address of the 12-float gradient accumulator of the delta parameters ([transition 3 | rotation-6d 6 | 0 0 0]): a
 * data-parallel caller all-reduces it between neus2_train_step_begin and the optimizer step. */
int neus2_delta_gradients_device(neus2_ctx* ctx, float** acc_dev);

#ifdef __cplusplus
}
#endif
#endif
