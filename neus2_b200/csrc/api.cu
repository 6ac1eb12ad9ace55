// api.cu -- context, buffers and the C ABI declared in include/neus2_b200.h.
// Host-side restatement of the Testbed/Trainer plumbing around the hot path:
//   reset_network (src/testbed.cu:2084-2349), Trainer::initialize_params (tcnn trainer.h:70-108),
//   NerfNetwork::initialize_params (nerf_network.h:815-905), Testbed::train (testbed.cu:2640-2740),
//   train_nerf / train_nerf_step (testbed_nerf.cu:3440-4001), update_density_grid_nerf (:3293-3397),
//   Counters::update_after_training (:3412-3438).
#include "../../include/neus2_b200.h"
#include "kernels.h"
#include "ngp_kernels.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

static thread_local std::string g_error;
unsigned long long g_neus_launches = 0;
void neus_set_error(const char* msg, const char* file, int line) {
	char buf[512];
	snprintf(buf, sizeof(buf), "%s (%s:%d)", msg, file, line);
	g_error = buf;
}
#define NEUS2_N_STAGES 12         // the first 8 are what neus2_get_stage_times reports; neus2_get_stage_times_ext adds the data-parallel sync points
#define NEUS2_N_STAGES_BASE 8
#define ADAM_BIAS_TABLE (1u << 20)
enum { ST_GRID = 0, ST_GEN = 1, ST_FWD = 2, ST_LOSS = 3, ST_FWDBWD = 4, ST_WGRAD = 5, ST_OPT = 6, ST_STEP = 7,
       ST_X_COUNTS = 8, ST_X_GRADS = 9, ST_BARRIER = 10, ST_SPARE = 11 };
#define FAIL(msg) do { neus_set_error(msg, __FILE__, __LINE__); return -1; } while (0)
#define CHECK(x) do { if ((x) != 0) return -1; } while (0)

struct neus2_ctx {
	neus2_config cfg;
	int device;
	GridMeta gm;
	ParamLayout pl;
	uint64_t n_params;
	uint32_t B;                 // target batch size (global)
	uint32_t B_local;           // rows this rank back-propagates (B / world_size)
	uint32_t max_rays;          // 1<<18
	uint32_t max_samples;       // 16*B
	uint32_t n_cascades;        // max_cascade + 1
	Aabb aabb;
	GlobalMove gmv;
	float cone_angle_constant;

	// Trainer buffers
	float* master; __half* params; __half* ema; float* grads; float* m1; float* m2; uint32_t* steps;
	uint32_t adam_step;
	float* adam_bias_table;     // sqrt(1-b2^k)/(1-b1^k) for per-parameter step counters k < ADAM_BIAS_TABLE (optimizer.cu)

	// dataset
	DatasetDev ds; uint8_t* d_pixels; float* d_focal; float* d_principal; float* d_xforms;

	// training scratch
	uint32_t *numsteps_all, *base, *slot, *ray_indices, *numsteps, *scan_tmp, *counters, *n_windows;
	uint32_t* counts_global;    // {kept rays, N_u, compacted, 0}: copy of counters[0..2] after the backward pass; a data-parallel caller all-reduces it in place
	unsigned char* windows; float* ray_setup;
	float *rays, *coords, *ray_tmp, *coords_compacted, *loss, *ek_loss, *mask_loss;
	__half *mlp_out, *dloss;
	StagedOperands so;
	float *partials, *var_acc, *w1row0_acc, *scalar_tmp;
	uint32_t* fwdbwd_done;      // CTA arrival counter of the training kernel (self-resetting)
	// prefix sums across the blocks of one launch (lookback.cuh): descriptors (never cleared, validated by epoch), ticket counters, launch epoch
	unsigned long long *lb_loss_desc, *lb_grid_desc; uint32_t* lb_tickets; uint32_t lb_epoch;   // tickets: [0] loss, [1] occupancy-grid samples
	// the loss by samples (loss.cu): {alpha, r, g, b} per marched sample from the first pass, and the loss kernels' own scratch
	float4* composite_in; LossScratch loss_ws;

	// occupancy grid
	float *density_grid, *density_grid_tmp, *grid_positions, *density_grid_mean;
	uint32_t* grid_indices; __half* grid_density; uint8_t* bitfield;
	uint32_t density_grid_ema_step;

	// Testbed state
	Pcg32 rng, density_grid_rng;
	uint32_t training_step, rays_per_batch, n_rays_total, measured_before;
	uint32_t measured_after;    // counters_rgb.measured_batch_size of the last step (compacted samples)
	uint32_t step_rays, step_rays_local, step_n_rays_total;   // of the step in flight
	// incremental (per-frame) training of a sequence: Testbed::training_network_next_frame (src/testbed.cu:2001-2082)
	uint32_t time_frame;            // current_training_time_frame
	uint32_t frozen_steps;          // predict_global_movement_training_step: canonical parameters are frozen for the first steps of frames >= 1
	float learning_rate;            // current Adam learning rate (after_learning_rate from frame 1 on)
	// DeltaNetwork: the per-frame rigid motion (transform_network.h:23-480) with its own trainer (testbed.cu:2320-2328).
	// Parameter order [transition 4 | rotation-6d 8] (DeltaNetwork::set_params).  d_xform is what the kernels read.
	float *d_xform, *d_acc, *d_master, *d_m1, *d_m2; uint32_t* d_steps; __half* d_params;
	bool use_delta;             // m_use_delta == m_train_delta for frames >= 1 (testbed.cu:2704-2710)
	uint32_t delta_opt_step;    // the global-move Adam's step count (drives its ExponentialDecay)
	float delta_lr_factor;
	bool step_density_updated;
	uint32_t* h_counters;       // pinned
	int profiling;              // record CUDA events around the stages of a step
	cudaEvent_t ev[2][2 * NEUS2_N_STAGES];   // two sets, alternating per step: a set is read back one step later, without a sync
	int ev_set;
	float stage_ms_sum[NEUS2_N_STAGES];
	uint32_t stage_count[NEUS2_N_STAGES];
	bool stage_pending[2][NEUS2_N_STAGES];
	// Pipelined single-GPU step (config.prefetch_next_samples): the step's one host read-back (sample counts) happens right after the
	// loss kernel, so while the GPU runs backward pass + optimizer the host already knows rays_per_batch of the NEXT step and
	// enqueues its sample generation on aux_stream, ordered behind this step's fused forward+backward kernel (which owns every
	// SM's register file) -- it overlaps the weight-gradient GEMMs and Adam.  The counters ping-pong between two buffers.
	struct StepPlan { uint32_t R, n_local, ray_offset, ray_stride, max_inference, n_rays_total, training_step; uint64_t rng_state, rng_inc; };
	cudaStream_t aux_stream;
	cudaEvent_t ev_counts, ev_fwdbwd_done, ev_gen_done, ev_loss_done;
	uint32_t* counters_buf[2]; int counters_cur;
	bool prefetched; StepPlan prefetch_plan; uint64_t prefetch_epoch;
	bool step_pipelined;        // the step in flight took the pipelined path (its counts were read back after the loss kernel)
	bool step_used_prefetch;    // ... and its samples had been generated during the previous step
	uint32_t step_rank, step_world;
	bool fuse_zero_into_adam;   // set by train_step_end around its optimizer call (single-GPU pipelined step)
	bool aux_pending;           // work is in flight on aux_stream that the next step must wait for (ev_gen_done)
	uint64_t state_epoch;       // bumped by every call that changes what sample generation reads (dataset, bitfield, rigid motion, step state)
	float* h_scalars;           // pinned
	// data-parallel peers (CUDA IPC mappings of every rank's gradient / fp16 parameter / EMA buffers; slot [rank] = own)
	uint32_t peer_world, peer_rank;
	float* peer_grads[NEUS2_MAX_PEERS]; __half* peer_params[NEUS2_MAX_PEERS]; __half* peer_ema[NEUS2_MAX_PEERS];
	PeerSyncBlock* sync_block;  // this rank's flags + payload slots, written by the peers (peer_sync.cu)
	PeerSyncPtrs peer_sync;     // every rank's block
	uint32_t exchange_epoch, barrier_epoch;
	uint32_t* sync_done;        // arrival counters (self-resetting) of the kernels that carry a sync point in their last block: [0] sharded optimizer
	bool grads_exchange_fused;  // the training kernel of the step in flight carried the gradient-ready exchange: neus2_peer_exchange has nothing left to do
	bool barrier_fused;         // the sharded optimizer of the step in flight carried the barrier: neus2_peer_barrier has nothing left to do
	bool sharded_zeroed;        // the sharded optimizer of the step in flight cleared the gradient quads it consumed (on every rank)
	bool barrier_zeroed;        // ... and the peer barrier behind it cleared the step's small accumulators and counters
	bool optimizer_done;        // the sharded optimizer already ran for the step in flight
	bool scratch_zeroed;        // train_step_end already enqueued the next step's zero-fills (counters, per-ray losses, gradients) behind its read-back
	cudaEvent_t readback_event; // marks the end of train_step_end's device->host copies
	// inference render scratch (sized on first use / growth, then reused)
	size_t render_cap;          // tile pixels
	void* r_queue; uint32_t* r_counters; float *r_ray_rgba, *r_ray_depth, *r_accumulate; uint32_t* r_ray_samples;
	// marching cubes
	size_t mc_cells_cap, mc_blocks_cap, mc_verts_cap, mc_indices_cap;
	uint32_t *mc_block_counts, *mc_counts, *mc_vert_prefix, *mc_indices; float* mc_verts;
	uint32_t mc_n_verts, mc_n_indices;
};

// Trainer::params_inference(): the optimizer's EMA copy once it has stepped, the training parameters before that (trainer.h:78-109, 170-178:
// m_params_inference is only redirected to the optimizer's custom weights by the first optimizer_step)
static const __half* inference_or_training_params(const neus2_ctx* ctx, int use_inference_params);
static void stage_begin(neus2_ctx* ctx, int st, cudaStream_t s);
static void stage_end(neus2_ctx* ctx, int st, cudaStream_t s);
static const __half* inference_or_training_params(const neus2_ctx* ctx, int use_inference_params) {
	return (use_inference_params && ctx->adam_step > 0) ? ctx->ema : ctx->params;
}
// m_train_canonical == false (testbed.cu:2075, 2659-2667): frames >= 1 train only the DeltaNetwork for their first frozen_steps steps
static bool canonical_frozen(const neus2_ctx* ctx) { return ctx->time_frame >= 1 && ctx->training_step < ctx->frozen_steps; }
// m_canonical_training_step (testbed_nerf.cu:3512-3527; reset to 0 by training_network_next_frame, testbed.cu:2015, 2050): equal to the
// training step except on frames >= 1 with global-movement prediction, where it stays 0 through the delta-only phase and counts from
// there.  It drives the occupancy-grid cadence (testbed.cu:2669), the uniform-sampling phase (testbed_nerf.cu:4011) and the restart of
// the image round-robin (testbed_nerf.cu:3785).
extern "C" uint32_t neus2_canonical_training_step(uint32_t time_frame, uint32_t frozen_steps, uint32_t training_step) {
	if (time_frame >= 1 && frozen_steps > 0) return training_step > frozen_steps ? training_step - frozen_steps : 0u;
	return training_step;
}
static uint32_t canonical_training_step(const neus2_ctx* ctx) { return neus2_canonical_training_step(ctx->time_frame, ctx->frozen_steps, ctx->training_step); }
static int dmalloc(void** p, size_t bytes) {
	CUDA_CHECK(cudaMalloc(p, bytes));
	CUDA_CHECK(cudaMemset(*p, 0, bytes));
	return 0;
}
#define DMALLOC(ptr, count) CHECK(dmalloc((void**)&(ptr), sizeof(*(ptr)) * (size_t)(count)))

static void stage_begin_set(neus2_ctx* ctx, int set, int st, cudaStream_t s) { if (ctx->profiling) cudaEventRecord(ctx->ev[set][2 * st], s); }
static void stage_end_set(neus2_ctx* ctx, int set, int st, cudaStream_t s) { if (ctx->profiling) { cudaEventRecord(ctx->ev[set][2 * st + 1], s); ctx->stage_pending[set][st] = true; } }
static void stage_begin(neus2_ctx* ctx, int st, cudaStream_t s) { stage_begin_set(ctx, ctx->ev_set, st, s); }
static void stage_end(neus2_ctx* ctx, int st, cudaStream_t s) { stage_end_set(ctx, ctx->ev_set, st, s); }
static void stage_collect(neus2_ctx* ctx, int set) {   // call once the set's events have completed
	if (!ctx->profiling) return;
	for (int st = 0; st < NEUS2_N_STAGES; ++st) {
		if (!ctx->stage_pending[set][st]) continue;
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, ctx->ev[set][2 * st], ctx->ev[set][2 * st + 1]) == cudaSuccess) { ctx->stage_ms_sum[st] += ms; ctx->stage_count[st]++; }
		ctx->stage_pending[set][st] = false;
	}
}
// Anything that changes what sample generation reads makes samples generated ahead of time stale; wait for them before the
// caller touches the buffers they read.
static void invalidate_prefetch(neus2_ctx* ctx) {
	++ctx->state_epoch;
	if (ctx->prefetched) { cudaEventSynchronize(ctx->ev_gen_done); ctx->prefetched = false; }
}

template <typename T> static int regrow(T** p, size_t* cap, size_t need) {
	if (need <= *cap && *p) return 0;
	if (*p) cudaFree(*p);
	*p = nullptr; *cap = 0;
	if (need == 0) need = 1;
	CUDA_CHECK(cudaMalloc((void**)p, need * sizeof(T)));
	*cap = need;
	return 0;
}

// Bias correction of adam_step (adam.h:149-150), sqrt(1 - beta2^k) / (1 - beta1^k), tabulated per step count k.  Lives here because
// this translation unit is compiled with nvcc's default FMA contraction, as the reference's optimizer is: libdevice's powf then
// compiles to the same instruction sequence as inside the reference's adam_step.
__global__ void k_adam_bias_table(float beta1, float beta2, uint32_t first, uint32_t last, float* __restrict__ table) {
	const uint32_t cs = first + blockIdx.x * blockDim.x + threadIdx.x;
	if (cs <= last) table[cs] = sqrtf(1 - powf(beta2, (float)cs)) / (1 - powf(beta1, (float)cs));
}
int launch_adam_bias_table(cudaStream_t stream, float beta1, float beta2, uint32_t first, uint32_t last, float* bias_table, uint32_t bias_table_size) {
	if (!bias_table || !bias_table_size || last < first) return 0;
	if (last >= bias_table_size) last = bias_table_size - 1;
	if (first > last) return 0;
	k_adam_bias_table<<<(last - first + 256) / 256, 256, 0, stream>>>(beta1, beta2, first, last, bias_table);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}

// ---- DeltaNetwork parameters -------------------------------------------------------------------------------------------
// One warp.  Optionally one Adam step of the global-move trainer on the 12 parameters (adam_step, tcnn optimizers/adam.h:52-160:
// no matrix parameters, so zero gradients are skipped and there is no L2 term; the gradient is the fp32 batch sum rounded to half,
// transform_network.h:214-243), then the float image the kernels read is rebuilt from the half parameters
// (rotation_6d_to_matrix, common_operation.cuh:37-60) and the gradient accumulators are cleared.
__global__ void k_delta_update(int do_step, float loss_scale, float lr, float beta1, float beta2, float eps, float off0, float off1, float off2,
                               float* __restrict__ acc, float* __restrict__ master, __half* __restrict__ params, float* __restrict__ m1, float* __restrict__ m2,
                               uint32_t* __restrict__ steps, float* __restrict__ xform) {
	const int k = threadIdx.x;
	if (do_step && k < 12) {
		const int ai = k < 3 ? k : (k >= 4 && k < 10 ? k - 1 : -1);
		const float g_half = ai >= 0 ? __half2float(__float2half_rn(acc[ai])) : 0.f;
		const float gradient = g_half / loss_scale;
		if (gradient != 0) {
			const float w = master[k];
			const float gsq = __fmul_rn(gradient, gradient);
			const float fm = fmaf(beta1, m1[k], __fmul_rn(1 - beta1, gradient));
			const float sm = fmaf(beta2, m2[k], __fmul_rn(1 - beta2, gsq));
			m1[k] = fm; m2[k] = sm;
			const uint32_t cs = ++steps[k];
			const float learning_rate = __fmul_rn(lr, sqrtf(1 - powf(beta2, (float)cs)) / (1 - powf(beta1, (float)cs)));
			const float eff = fminf(fmaxf(learning_rate / (sqrtf(sm) + eps), 0.f), 3.402823466e+38f);
			const float nw = fmaf(-eff, fm, w);
			master[k] = nw;
			params[k] = __float2half_rn(nw);
		}
	}
	__syncwarp();
	if (k == 0) {
		float a1[3], a2[3], b1[3], u2[3], b2[3], b3[3];
		#pragma unroll
		for (int d = 0; d < 3; ++d) { a1[d] = __half2float(params[4 + d]); a2[d] = __half2float(params[7 + d]); }
		normalize3(a1, b1);
		const float dd = fmaf(b1[2], a2[2], fmaf(b1[1], a2[1], __fmul_rn(b1[0], a2[0])));
		#pragma unroll
		for (int d = 0; d < 3; ++d) u2[d] = fmaf(-dd, b1[d], a2[d]);
		normalize3(u2, b2);
		b3[0] = fmaf(b1[1], b2[2], -__fmul_rn(b1[2], b2[1]));
		b3[1] = fmaf(b1[2], b2[0], -__fmul_rn(b1[0], b2[2]));
		b3[2] = fmaf(b1[0], b2[1], -__fmul_rn(b1[1], b2[0]));
		const float off[3] = {off0, off1, off2};
		#pragma unroll
		for (int d = 0; d < 3; ++d) {
			xform[DELTA_R + d * 3 + 0] = b1[d]; xform[DELTA_R + d * 3 + 1] = b2[d]; xform[DELTA_R + d * 3 + 2] = b3[d];
			const float t = __half2float(params[d]);
			xform[DELTA_T + d] = t; xform[DELTA_TMO + d] = __fsub_rn(t, off[d]); xform[DELTA_OFF + d] = off[d];
			xform[DELTA_A1 + d] = a1[d]; xform[DELTA_A2 + d] = a2[d]; xform[DELTA_U2 + d] = u2[d];
		}
	}
	if (k < 12) acc[k] = 0.f;
}

static int launch_delta_update(neus2_ctx* ctx, cudaStream_t stream, int do_step, float loss_scale, float lr) {
	const neus2_config& c = ctx->cfg;
	k_delta_update<<<1, 32, 0, stream>>>(do_step, loss_scale, lr, c.delta_beta1, c.delta_beta2, c.delta_epsilon, ctx->gmv.first_frame_offset[0],
	                                     ctx->gmv.first_frame_offset[1], ctx->gmv.first_frame_offset[2], ctx->d_acc, ctx->d_master, ctx->d_params, ctx->d_m1, ctx->d_m2,
	                                     ctx->d_steps, ctx->d_xform);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}
// A fresh global-move trainer (reset_network_incremental, testbed.cu:2569-2577 -> DeltaNetwork::initialize_params,
// transform_network.h:352-396): transition 0, rotation-6d (1,0,0, 0,1,0), optimizer state cleared.
static int reset_delta(neus2_ctx* ctx, cudaStream_t stream) {
	const float init[12] = {0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
	__half init_h[12];
	for (int k = 0; k < 12; ++k) init_h[k] = __float2half_rn(init[k]);
	CUDA_CHECK(cudaMemcpyAsync(ctx->d_master, init, sizeof(init), cudaMemcpyHostToDevice, stream));
	CUDA_CHECK(cudaMemcpyAsync(ctx->d_params, init_h, sizeof(init_h), cudaMemcpyHostToDevice, stream));
	CUDA_CHECK(cudaMemsetAsync(ctx->d_m1, 0, 12 * sizeof(float), stream));
	CUDA_CHECK(cudaMemsetAsync(ctx->d_m2, 0, 12 * sizeof(float), stream));
	CUDA_CHECK(cudaMemsetAsync(ctx->d_steps, 0, 12 * sizeof(uint32_t), stream));
	CUDA_CHECK(cudaMemsetAsync(ctx->d_acc, 0, 12 * sizeof(float), stream));
	CUDA_CHECK(cudaStreamSynchronize(stream));   // the host arrays above live on this stack frame
	ctx->delta_opt_step = 0; ctx->delta_lr_factor = 1.0f;
	return launch_delta_update(ctx, stream, 0, 1.f, 0.f);
}
// ExponentialDecayOptimizer::step (tcnn optimizers/exponential_decay.h:60-71) around the Adam step of the global-move trainer
static int delta_optimizer_step(neus2_ctx* ctx, cudaStream_t stream, float loss_scale) {
	const neus2_config& c = ctx->cfg;
	const uint32_t st = ctx->delta_opt_step;
	if (st == 0) ctx->delta_lr_factor = 1.0f;
	if (c.delta_decay_interval && st >= c.delta_decay_start && (st - c.delta_decay_start) % c.delta_decay_interval == 0) ctx->delta_lr_factor *= c.delta_decay_base;
	++ctx->delta_opt_step;
	return launch_delta_update(ctx, stream, 1, loss_scale, c.delta_learning_rate * ctx->delta_lr_factor);
}

extern "C" {

const char* neus2_last_error(void) { return g_error.c_str(); }

void neus2_default_config(neus2_config* c) {
	memset(c, 0, sizeof(*c));
	c->n_levels = 14; c->log2_hashmap_size = 19; c->base_resolution = 16; c->top_resolution = 2048.f; c->per_level_scale = 0.f;
	c->valid_level_scale = 0.02f; c->base_valid_level_scale = 0.2f; c->base_training_step = 100;
	c->sdf_bias = -0.1f;
	c->learning_rate = 1e-3f; c->beta1 = 0.9f; c->beta2 = 0.99f; c->epsilon = 1e-15f; c->l2_reg = 1e-6f; c->ema_decay = 0.95f;
	c->mask_loss_weight = 0.f; c->ek_loss_weight = 0.01f; c->anneal_end = 0; c->batch_size = 1u << 18;
	c->aabb_scale = 1; c->seed = 1337; c->random_bg_color = 1; c->snap_to_pixel_centers = 1;
	c->background_color[0] = c->background_color[1] = c->background_color[2] = 0.f;
	c->density_grid_decay = 0.95f; c->loss_scale = 128.f;
	c->world_size = 1; c->max_rays_per_batch = 1u << 18; c->forward_early_termination = 1; c->prefetch_next_samples = 1;
	// "globalmove" optimizer of configs/nerf/base.json:135-158 (Ema{ExponentialDecay{Adam}})
	c->delta_learning_rate = 1e-4f; c->delta_beta1 = 0.9f; c->delta_beta2 = 0.99f; c->delta_epsilon = 1e-10f;
	c->delta_decay_start = 50; c->delta_decay_interval = 25; c->delta_decay_base = 0.99f;
	c->decay_start = 20000; c->decay_interval = 10000; c->decay_end = 10000000; c->decay_base = 0.33f;
}

// Counters::update_after_training (src/testbed_nerf.cu:3435-3436): rays for the next step so that it yields target_batch_size compacted samples,
// rounded up to the batch granularity (128 rays per rank here, tcnn::batch_size_granularity there) and capped (1 << 18 in the reference).
uint32_t neus2_next_rays_per_batch(uint32_t rays_per_batch, uint32_t target_batch_size, uint32_t measured_batch_size, uint32_t world_size, uint32_t max_rays_per_batch) {
	uint32_t rpb = (uint32_t)((float)rays_per_batch * (float)target_batch_size / (float)measured_batch_size);
	const uint32_t gran = 128u * (world_size ? world_size : 1u);
	return std::min((rpb + gran - 1) / gran * gran, max_rays_per_batch);
}

// ExponentialDecayOptimizer::step (tcnn optimizers/exponential_decay.h:60-71): with s = nested->step() BEFORE the nested step,
// "if (s >= decay_start && (s - decay_start) % decay_interval == 0 && s <= decay_end) factor *= decay_base".  The factor after all triggers up to s.
float neus2_learning_rate_factor(const neus2_config* c, uint32_t s) {
	float factor = 1.0f;
	if (!c->decay_interval || s < c->decay_start) return factor;
	const uint32_t last = s < c->decay_end ? s : c->decay_end;
	if (last < c->decay_start) return factor;
	const uint32_t triggers = (last - c->decay_start) / c->decay_interval + 1;
	for (uint32_t k = 0; k < triggers; ++k) factor *= c->decay_base;
	return factor;
}

static int build_tables(neus2_ctx* ctx) {
	const neus2_config& c = ctx->cfg;
	if (c.n_levels < 1 || c.n_levels > NEUS_MAX_LEVELS) FAIL("n_levels out of range (1..16)");
	if (3 + 2 * c.n_levels > 32) FAIL("this build fixes the sdf-MLP input width at 32: need 3 + 2*n_levels <= 32");
	float pls = c.per_level_scale;
	if (pls <= 0.f && c.n_levels > 1) pls = std::exp(std::log(c.top_resolution * (float)c.aabb_scale / (float)c.base_resolution) / (c.n_levels - 1));   // testbed.cu:2181
	if (pls <= 0.f) pls = 1.f;
	GridMeta& gm = ctx->gm;
	memset(&gm, 0, sizeof(gm));
	gm.n_levels = c.n_levels;
	gm.valid_level = c.n_levels;
	uint32_t offset = 0;
	for (uint32_t i = 0; i < c.n_levels; ++i) {   // grid.h:1472-1505
		const float scale = exp2f(i * std::log2(pls)) * c.base_resolution - 1.0f;
		const uint32_t resolution = (uint32_t)(ceilf(scale)) + 1;
		gm.scale[i] = (float)(resolution - 1);
		gm.res[i] = resolution;
		const uint32_t max_params = std::numeric_limits<uint32_t>::max() / 2;
		uint32_t params_in_level = std::pow((float)resolution, 3) > (float)max_params ? max_params : resolution * resolution * resolution;
		params_in_level = (params_in_level + 7u) / 8u * 8u;
		params_in_level = std::min(params_in_level, (1u << c.log2_hashmap_size));
		gm.offset[i] = offset;
		if ((uint64_t)resolution * resolution * resolution > (uint64_t)params_in_level) gm.hashed_mask |= 1u << i;
		offset += params_in_level;
	}
	gm.offset[c.n_levels] = offset;
	ParamLayout& pl = ctx->pl;
	pl.sdf_in_w = 32;
	pl.off_w0 = 0; pl.off_w1 = 64 * 32;
	pl.off_c0 = pl.off_w1 + 16 * 64; pl.off_c1 = pl.off_c0 + 64 * 48; pl.off_c2 = pl.off_c1 + 64 * 64;
	pl.off_grid = pl.off_c2 + 16 * 64;
	pl.off_var = pl.off_grid + offset * 2;
	pl.n_params = pl.off_var + 4;
	ctx->n_params = pl.n_params;
	return 0;
}

int neus2_ctx_create(const neus2_config* cfg, int device, neus2_ctx** out) {
	if (!cfg || !out) FAIL("null argument");
	CUDA_CHECK(cudaSetDevice(device));
	neus2_ctx* ctx = new neus2_ctx();
	memset(ctx, 0, sizeof(*ctx));
	ctx->cfg = *cfg;
	ctx->device = device;
	if (build_tables(ctx)) { delete ctx; return -1; }
	ctx->B = cfg->batch_size;
	if (ctx->B == 0 || ctx->B % 128 != 0) { delete ctx; FAIL("batch_size must be a positive multiple of 128 (tcnn batch_size_granularity)"); }
	if (ctx->cfg.world_size == 0) ctx->cfg.world_size = 1;
	if (ctx->cfg.max_rays_per_batch == 0) ctx->cfg.max_rays_per_batch = 1u << 18;
	const uint32_t W = ctx->cfg.world_size;
	if (ctx->B % (128 * W) != 0 || ctx->cfg.max_rays_per_batch % (128 * W) != 0) { delete ctx; FAIL("batch_size and max_rays_per_batch must be multiples of 128*world_size"); }
	ctx->max_rays = ctx->cfg.max_rays_per_batch / W;     // this rank's share
	ctx->B_local = ctx->B / W;
	ctx->max_samples = ctx->B_local * 16;
	uint32_t max_cascade = 0;
	while ((1 << max_cascade) < cfg->aabb_scale) ++max_cascade;
	ctx->n_cascades = max_cascade + 1;
	const float half_ext = 0.5f * std::min(1 << (NERF_CASCADES - 1), cfg->aabb_scale);   // testbed_nerf.cu:3075-3076
	for (int k = 0; k < 3; ++k) { ctx->aabb.mn[k] = 0.5f - half_ext; ctx->aabb.mx[k] = 0.5f + half_ext; }
	ctx->cone_angle_constant = cfg->aabb_scale <= 1 ? 0.0f : (1.0f / 256.0f);            // :3090
	const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
	memcpy(ctx->gmv.R, I, sizeof(I));

	const uint64_t N = ctx->n_params;
	const uint32_t B = ctx->B_local, R = ctx->max_rays, S = ctx->max_samples;
	#define CTXCHECK(x) do { if ((x) != 0) { neus2_ctx_destroy(ctx); return -1; } } while (0)
	#define CDM(ptr, count) CTXCHECK(dmalloc((void**)&(ptr), sizeof(*(ptr)) * (size_t)(count)))
	CDM(ctx->master, N); CDM(ctx->params, N); CDM(ctx->ema, N); CDM(ctx->grads, N); CDM(ctx->m1, N); CDM(ctx->m2, N); CDM(ctx->steps, N); CDM(ctx->adam_bias_table, ADAM_BIAS_TABLE);
	CDM(ctx->numsteps_all, R); CDM(ctx->base, R); CDM(ctx->slot, R); CDM(ctx->ray_indices, R); CDM(ctx->numsteps, 2 * (size_t)R); CDM(ctx->scan_tmp, 2 * (size_t)R);
	CDM(ctx->n_windows, R); CDM(ctx->windows, sampling_window_bytes_per_ray() * (size_t)R); CDM(ctx->ray_setup, 12 * (size_t)R);
	CDM(ctx->counters_buf[0], 12); CDM(ctx->counters_buf[1], 12);
	ctx->counters = ctx->counters_buf[0]; ctx->counts_global = ctx->counters + 8;   // one 48-byte read-back per step covers both
	if (cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_counts, cudaEventDisableTiming) != cudaSuccess ||
	    cudaEventCreateWithFlags(&ctx->ev_fwdbwd_done, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_gen_done, cudaEventDisableTiming) != cudaSuccess ||
	    cudaEventCreateWithFlags(&ctx->ev_loss_done, cudaEventDisableTiming) != cudaSuccess) {
		neus2_ctx_destroy(ctx); FAIL("stream / event creation failed");
	}
	CDM(ctx->rays, 6 * (size_t)R); CDM(ctx->coords, 7 * (size_t)S); CDM(ctx->ray_tmp, 12 * (size_t)R);
	CDM(ctx->coords_compacted, 7 * (size_t)B); CDM(ctx->loss, R); CDM(ctx->ek_loss, R); CDM(ctx->mask_loss, R);
	CDM(ctx->mlp_out, 16 * (size_t)S); CDM(ctx->dloss, 16 * (size_t)B);
	if (!weight_gradients_in_kernel()) {   // NEUS2_WGRAD=staged: operand matrices for the separate weight-gradient GEMM kernel (243 MB at B = 2^18)
		CDM(ctx->so.dOUT, 16 * (size_t)B); CDM(ctx->so.H2, 64 * (size_t)B); CDM(ctx->so.dH2, 64 * (size_t)B); CDM(ctx->so.H1, 64 * (size_t)B); CDM(ctx->so.dH1, 64 * (size_t)B);
		CDM(ctx->so.CIN, 48 * (size_t)B); CDM(ctx->so.dSO, 16 * (size_t)B); CDM(ctx->so.H0, 64 * (size_t)B); CDM(ctx->so.dH0, 64 * (size_t)B);
		CDM(ctx->so.X0, 32 * (size_t)B); CDM(ctx->so.dI, 32 * (size_t)B);
		CDM(ctx->partials, (size_t)weight_gradient_partials() * ctx->pl.off_grid);
	}
	CDM(ctx->var_acc, 1); CDM(ctx->w1row0_acc, 64); CDM(ctx->scalar_tmp, 4); CDM(ctx->fwdbwd_done, 1);
	if (ctx->cfg.forward_early_termination) {
		CDM(ctx->composite_in, S); CDM(ctx->loss_ws.prefix4, S); CDM(ctx->loss_ws.prefix1, S);
		CDM(ctx->loss_ws.slot_sample, B); CDM(ctx->loss_ws.slot_ray, B); CDM(ctx->loss_ws.ray_rec, 12 * (size_t)R);
	}
	CDM(ctx->lb_loss_desc, (size_t)R / 8 + 2); CDM(ctx->lb_grid_desc, (size_t)NERF_GRIDSIZE * NERF_GRIDSIZE * NERF_GRIDSIZE / 256 + 2); CDM(ctx->lb_tickets, 4);
	const size_t G3n = (size_t)NERF_GRIDSIZE * NERF_GRIDSIZE * NERF_GRIDSIZE;
	CDM(ctx->density_grid, G3n * ctx->n_cascades); CDM(ctx->density_grid_tmp, G3n * ctx->n_cascades);
	CDM(ctx->grid_positions, 3 * G3n * ctx->n_cascades); CDM(ctx->grid_indices, G3n * ctx->n_cascades); CDM(ctx->grid_density, G3n * ctx->n_cascades);
	CDM(ctx->bitfield, G3n / 8 * NERF_CASCADES); CDM(ctx->density_grid_mean, 1);
	CTXCHECK(launch_adam_bias_table(nullptr, ctx->cfg.beta1, ctx->cfg.beta2, 1, ADAM_BIAS_TABLE - 1, ctx->adam_bias_table, ADAM_BIAS_TABLE));   // once: the optimizer only looks it up
	CDM(ctx->sync_block, 1); CDM(ctx->sync_done, 4);
	CDM(ctx->d_xform, DELTA_XFORM_FLOATS); CDM(ctx->d_acc, 12); CDM(ctx->d_master, 12); CDM(ctx->d_m1, 12); CDM(ctx->d_m2, 12); CDM(ctx->d_steps, 12); CDM(ctx->d_params, 12);
	CTXCHECK(reset_delta(ctx, 0));
	if (cudaMallocHost((void**)&ctx->h_counters, 16 * sizeof(uint32_t)) != cudaSuccess || cudaMallocHost((void**)&ctx->h_scalars, 4 * sizeof(float)) != cudaSuccess) {
		neus2_ctx_destroy(ctx); FAIL("cudaMallocHost failed");
	}
	// reset_network state (testbed.cu:2087-2101)
	ctx->rng = Pcg32(cfg->seed);
	ctx->density_grid_rng = Pcg32(ctx->rng.next_uint());
	(void)ctx->rng.next_uint();   // tv_loss_rng
	ctx->rays_per_batch = std::min(1u << 12, ctx->cfg.max_rays_per_batch);   // testbed.h:606, within what the per-ray arrays hold
	ctx->measured_before = 0;
	ctx->training_step = 0;
	ctx->n_rays_total = 0;
	ctx->time_frame = 0; ctx->frozen_steps = 0; ctx->learning_rate = cfg->learning_rate;
	*out = ctx;
	return 0;
}

void neus2_ctx_destroy(neus2_ctx* ctx) {
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
	for (cudaEvent_t e : {ctx->ev_counts, ctx->ev_fwdbwd_done, ctx->ev_gen_done, ctx->ev_loss_done}) if (e) cudaEventDestroy(e);
	ctx->counters = nullptr;   // freed through counters_buf below
	void* ptrs[] = {ctx->master, ctx->params, ctx->ema, ctx->grads, ctx->m1, ctx->m2, ctx->steps, ctx->adam_bias_table, ctx->d_pixels, ctx->d_focal, ctx->d_principal, ctx->d_xforms,
		ctx->numsteps_all, ctx->base, ctx->slot, ctx->ray_indices, ctx->numsteps, ctx->scan_tmp, ctx->counters, ctx->n_windows, ctx->windows, ctx->ray_setup, ctx->rays, ctx->coords, ctx->ray_tmp,
		ctx->coords_compacted, ctx->loss, ctx->ek_loss, ctx->mask_loss, ctx->mlp_out, ctx->dloss, ctx->so.dOUT, ctx->so.H2, ctx->so.dH2, ctx->so.H1, ctx->so.dH1,
		ctx->so.CIN, ctx->so.dSO, ctx->so.H0, ctx->so.dH0, ctx->so.X0, ctx->so.dI, ctx->partials, ctx->var_acc, ctx->w1row0_acc, ctx->scalar_tmp, ctx->fwdbwd_done,
		ctx->density_grid, ctx->density_grid_tmp, ctx->grid_positions, ctx->grid_indices, ctx->grid_density, ctx->bitfield, ctx->density_grid_mean,
		ctx->d_xform, ctx->d_acc, ctx->d_master, ctx->d_m1, ctx->d_m2, ctx->d_steps, ctx->d_params, ctx->counters_buf[0], ctx->counters_buf[1], ctx->sync_block, ctx->sync_done, ctx->lb_loss_desc, ctx->lb_grid_desc, ctx->lb_tickets, ctx->composite_in, ctx->loss_ws.prefix4, ctx->loss_ws.prefix1, ctx->loss_ws.slot_sample, ctx->loss_ws.slot_ray, ctx->loss_ws.ray_rec};
	for (uint32_t r = 0; r < ctx->peer_world; ++r) if (r != ctx->peer_rank) {   // unmap the peers' buffers before this rank's own are freed
		if (ctx->peer_grads[r]) cudaIpcCloseMemHandle(ctx->peer_grads[r]);
		if (ctx->peer_params[r]) cudaIpcCloseMemHandle(ctx->peer_params[r]);
		if (ctx->peer_ema[r]) cudaIpcCloseMemHandle(ctx->peer_ema[r]);
		if (ctx->peer_sync.block[r]) cudaIpcCloseMemHandle(ctx->peer_sync.block[r]);
	}
	for (void* p : ptrs) if (p) cudaFree(p);
	void* more[] = {ctx->r_queue, ctx->r_counters, ctx->r_ray_rgba, ctx->r_ray_depth, ctx->r_accumulate, ctx->r_ray_samples,
		ctx->mc_block_counts, ctx->mc_counts, ctx->mc_vert_prefix, ctx->mc_indices, ctx->mc_verts};
	for (void* p : more) if (p) cudaFree(p);
	if (ctx->readback_event) cudaEventDestroy(ctx->readback_event);
	if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
	if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
	delete ctx;
}

uint64_t neus2_n_params(const neus2_ctx* ctx) { return ctx->n_params; }

int neus2_param_layout(const neus2_ctx* ctx, uint32_t layout[8]) {
	const ParamLayout& p = ctx->pl;
	layout[0] = p.sdf_in_w; layout[1] = p.off_w0; layout[2] = p.off_w1; layout[3] = p.off_c0; layout[4] = p.off_c1; layout[5] = p.off_c2;
	layout[6] = p.off_grid; layout[7] = p.off_var;
	return 0;
}

int neus2_grid_tables(const neus2_ctx* ctx, uint32_t* offsets, uint32_t* resolutions, float* scales) {
	for (uint32_t i = 0; i <= ctx->gm.n_levels; ++i) offsets[i] = ctx->gm.offset[i];
	for (uint32_t i = 0; i < ctx->gm.n_levels; ++i) { resolutions[i] = ctx->gm.res[i]; scales[i] = ctx->gm.scale[i]; }
	return 0;
}

int neus2_buffers(neus2_ctx* ctx, float** master, void** params, void** ema, float** grads) {
	if (master) *master = ctx->master;
	if (params) *params = ctx->params;
	if (ema) *ema = ctx->ema;
	if (grads) *grads = ctx->grads;
	return 0;
}

__global__ void k_cast_params(uint64_t n, const float* __restrict__ master, __half* __restrict__ params) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) params[i] = __float2half_rn(master[i]);
}

__global__ void k_widen_params(uint64_t n, const __half* __restrict__ params, float* __restrict__ master) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) master[i] = __half2float(params[i]);
}

int neus2_set_params(neus2_ctx* ctx, const float* master_host, uint64_t n, int reset_optimizer) {
	if (n != ctx->n_params) FAIL("parameter count mismatch");
	ctx->scratch_zeroed = false;
	CUDA_CHECK(cudaSetDevice(ctx->device));
	CUDA_CHECK(cudaMemcpy(ctx->master, master_host, n * sizeof(float), cudaMemcpyHostToDevice));
	k_cast_params<<<(uint32_t)((n + 255) / 256), 256>>>(n, ctx->master, ctx->params);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	if (reset_optimizer) {
		CUDA_CHECK(cudaMemset(ctx->m1, 0, n * sizeof(float)));
		CUDA_CHECK(cudaMemset(ctx->m2, 0, n * sizeof(float)));
		CUDA_CHECK(cudaMemset(ctx->steps, 0, n * sizeof(uint32_t)));
		CUDA_CHECK(cudaMemset(ctx->ema, 0, n * sizeof(__half)));
		CUDA_CHECK(cudaMemset(ctx->grads, 0, n * sizeof(float)));
		ctx->adam_step = 0;
	}
	CUDA_CHECK(cudaDeviceSynchronize());
	return 0;
}

int neus2_get_params(neus2_ctx* ctx, float* master_host, uint64_t n) {
	if (n != ctx->n_params) FAIL("parameter count mismatch");
	CUDA_CHECK(cudaSetDevice(ctx->device));
	CUDA_CHECK(cudaMemcpy(master_host, ctx->master, n * sizeof(float), cudaMemcpyDeviceToHost));
	return 0;
}

// geometric init of the SDF MLP (geometry_init_save_weights.py:318-331) with our own generator (the reference's weight file is not in the snapshot)
static void geometric_init_sdf_mlp(float* w0 /*[64][32]*/, float* w1 /*[16][64]*/) {
	Pcg32 g{0x9e0u};
	auto normal = [&]() {
		float u1 = g.next_float(), u2 = g.next_float();
		if (u1 < 1e-7f) u1 = 1e-7f;
		return std::sqrt(-2.0f * std::log(u1)) * std::cos(6.283185307179586f * u2);
	};
	for (uint32_t o = 0; o < 64; ++o) for (uint32_t i = 0; i < 32; ++i) w0[o * 32 + i] = i < 3 ? normal() * (std::sqrt(2.0f) / std::sqrt(64.0f)) : 0.f;
	for (uint32_t i = 0; i < 16 * 64; ++i) w1[i] = std::sqrt(3.14159265358979f) / std::sqrt(64.0f) + 1e-5f * normal();
}

// NerfNetwork::initialize_sdf_mlp_params (nerf_network.h:888-908), used by "incremental_reinit_sdf_mlp" every
// incremental_reinit_sdf_mlp_iters frames (src/testbed.cu:2065-2068): the fp32 MASTER copy of the SDF MLP is reset to its geometric
// initialisation; like the reference, the fp16 training / inference copies are left alone and follow at the next optimizer step.
int neus2_reinit_sdf_mlp(neus2_ctx* ctx, void* stream_, const float* sdf_mlp_weights_host) {
	cudaStream_t stream = (cudaStream_t)stream_;
	CUDA_CHECK(cudaSetDevice(ctx->device));
	const ParamLayout& pl = ctx->pl;
	if (pl.off_w1 != pl.off_w0 + 64 * 32) FAIL("unexpected SDF-MLP layout");
	std::vector<float> w(64 * 32 + 16 * 64);
	if (sdf_mlp_weights_host) memcpy(w.data(), sdf_mlp_weights_host, w.size() * sizeof(float));
	else geometric_init_sdf_mlp(w.data(), w.data() + 64 * 32);
	CUDA_CHECK(cudaMemcpyAsync(ctx->master + pl.off_w0, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
	CUDA_CHECK(cudaStreamSynchronize(stream));
	return 0;
}

int neus2_init_params(neus2_ctx* ctx, const float* sdf_mlp_weights_host) {
	const ParamLayout& pl = ctx->pl;
	std::vector<float> p(ctx->n_params, 0.f);
	// Trainer rng (trainer.h:56-59)
	std::seed_seq seq{ctx->cfg.seed};
	std::vector<uint32_t> seeds(2);
	seq.generate(seeds.begin(), seeds.end());
	Pcg32 rng{seeds.front()};
	auto xavier = [&](float* dst, uint32_t fan_out, uint32_t fan_in) {   // gpu_matrix.h:292-307
		float scale = 1.0f;
		scale *= std::sqrt(6.0f / (float)(fan_in + fan_out));
		for (uint32_t i = 0; i < fan_out * fan_in; ++i) dst[i] = rng.next_float() * 2.0f * scale - scale;
	};
	xavier(&p[pl.off_w0], 64, 32); xavier(&p[pl.off_w1], 16, 64);      // consumed, then overwritten (nerf_network.h:819-833)
	if (sdf_mlp_weights_host) {
		memcpy(&p[pl.off_w0], sdf_mlp_weights_host, sizeof(float) * (64 * 32 + 16 * 64));
	} else {
		geometric_init_sdf_mlp(&p[pl.off_w0], &p[pl.off_w1]);
	}
	xavier(&p[pl.off_c0], 64, 48); xavier(&p[pl.off_c1], 64, 64); xavier(&p[pl.off_c2], 16, 64);
	{   // grid: generate_random_uniform(-1e-4, 1e-4) (grid.h:2375-2380; random.h:67-99)
		const uint64_t n = (uint64_t)ctx->gm.offset[ctx->gm.n_levels] * 2;
		const uint64_t n_threads = ((n + 3) / 4 + 127) / 128 * 128;
		float* out = &p[pl.off_grid];
		#pragma omp parallel for schedule(static)
		for (int64_t i = 0; i < (int64_t)n_threads; ++i) {
			Pcg32 r = rng; r.advance(i * 4);
			for (uint64_t j = 0; j < 4; ++j) {
				const uint64_t idx = (uint64_t)i + n_threads * j;
				if (idx >= n) break;
				out[idx] = fmaf(r.next_float(), 1e-4f - (-1e-4f), -1e-4f);
			}
		}
		rng.advance((int64_t)n);
	}
	for (int k = 0; k < 4; ++k) p[pl.off_var + k] = 0.3f;   // nerf_network.h:897-899
	return neus2_set_params(ctx, p.data(), ctx->n_params, 1);
}

int neus2_set_training_step(neus2_ctx* ctx, int32_t step) {   // grid.h:2427-2440
	const neus2_config& c = ctx->cfg;
	if (step <= 0) { ctx->gm.valid_level = c.n_levels; return 0; }
	const uint32_t v = (uint32_t)ceil(c.base_valid_level_scale * c.n_levels + c.valid_level_scale * std::max(0, (int)(step - (int)c.base_training_step)));
	ctx->gm.valid_level = std::min(c.n_levels, v);
	return 0;
}

int neus2_forward(neus2_ctx* ctx, void* stream, const float* coords, uint32_t n, void* out, int use_inference_params) {
	CUDA_CHECK(cudaSetDevice(ctx->device));
	return launch_network_forward((cudaStream_t)stream, ctx->gm, ctx->pl, inference_or_training_params(ctx, use_inference_params), coords, nullptr, n, ctx->cfg.sdf_bias, (__half*)out);
}

int neus2_sdf(neus2_ctx* ctx, void* stream, const float* pos, uint32_t stride, uint32_t n, void* out, int as_density, int use_inference_params) {
	CUDA_CHECK(cudaSetDevice(ctx->device));
	return launch_network_sdf((cudaStream_t)stream, ctx->gm, ctx->pl, inference_or_training_params(ctx, use_inference_params), pos, stride, n, ctx->cfg.sdf_bias, as_density, (__half*)out, nullptr);
}

static int backward_on_stream(neus2_ctx* ctx, cudaStream_t stream, const float* coords, const void* dL_dout, const uint32_t* n_in_dev, uint32_t n_rows,
                              uint32_t indeed_batch_size, bool grads_zeroed, bool fuse_grads_exchange = false) {
	if (n_rows > ctx->B_local) FAIL("n_rows exceeds the context's (local) batch size");
	if (!grads_zeroed) {
		CUDA_CHECK(cudaMemsetAsync(ctx->grads, 0, ctx->n_params * sizeof(float), stream));   // EGradientMode::Overwrite
		CUDA_CHECK(cudaMemsetAsync(ctx->var_acc, 0, sizeof(float), stream));
		CUDA_CHECK(cudaMemsetAsync(ctx->w1row0_acc, 0, 64 * sizeof(float), stream));
	}
	stage_begin(ctx, ST_FWDBWD, stream);
	PeerXfer gx = {};
	ctx->grads_exchange_fused = false;
	if (fuse_grads_exchange) {   // data-parallel with peer mappings: the kernel's last CTA tells the peers that this rank's gradients are complete and waits for theirs
		gx.peers = ctx->peer_sync; gx.epoch = ++ctx->exchange_epoch; gx.rank = ctx->peer_rank; gx.world = ctx->peer_world;
		ctx->grads_exchange_fused = true;
	}
	CHECK(launch_network_fwdbwd(stream, ctx->gm, ctx->pl, ctx->params, coords, (const __half*)dL_dout, n_in_dev, n_rows, ctx->cfg.sdf_bias, indeed_batch_size,
	                            ctx->grads, ctx->var_acc, ctx->w1row0_acc, ctx->so, ctx->use_delta ? ctx->d_acc : nullptr, ctx->fwdbwd_done, fuse_grads_exchange ? &gx : nullptr));
	stage_end(ctx, ST_FWDBWD, stream);
	CUDA_CHECK(cudaEventRecord(ctx->ev_fwdbwd_done, stream));
	if (!weight_gradients_in_kernel()) {   // default: the training kernel has accumulated the MLP weight gradients itself (tensor memory)
		stage_begin(ctx, ST_WGRAD, stream);
		CHECK(launch_weight_gradients(stream, ctx->pl, ctx->so, ctx->params, n_in_dev, n_rows, ctx->partials, weight_gradient_partials(), ctx->var_acc, ctx->w1row0_acc, ctx->grads));
		stage_end(ctx, ST_WGRAD, stream);
	}
	return 0;
}

// the zero-fills a step starts from: Counters::prepare_for_training_steps (testbed_nerf.cu:3399-3410) + EGradientMode::Overwrite
static int zero_step_scratch(neus2_ctx* ctx, cudaStream_t stream, uint32_t n_rays) {
	CUDA_CHECK(cudaMemsetAsync(ctx->counters, 0, 8 * sizeof(uint32_t), stream));
	CUDA_CHECK(cudaMemsetAsync(ctx->loss, 0, sizeof(float) * n_rays, stream));
	CUDA_CHECK(cudaMemsetAsync(ctx->ek_loss, 0, sizeof(float) * n_rays, stream));
	CUDA_CHECK(cudaMemsetAsync(ctx->mask_loss, 0, sizeof(float) * n_rays, stream));
	return 0;
}

int neus2_backward(neus2_ctx* ctx, void* stream_, const float* coords, const void* dL_dout, const uint32_t* n_in_dev, uint32_t n_rows, uint32_t indeed_batch_size) {
	CUDA_CHECK(cudaSetDevice(ctx->device));
	ctx->scratch_zeroed = false;   // an explicit operator call owns the gradient buffer again
	return backward_on_stream(ctx, (cudaStream_t)stream_, coords, dL_dout, n_in_dev, n_rows, indeed_batch_size, false);
}

int neus2_optimizer_step(neus2_ctx* ctx, void* stream, float loss_scale) {
	CUDA_CHECK(cudaSetDevice(ctx->device));
	stage_begin(ctx, ST_OPT, (cudaStream_t)stream);
	if (canonical_frozen(ctx)) {
		// "if (m_train_canonical) m_trainer->optimizer_step(...)" (testbed_nerf.cu:3503-3505): while only the DeltaNetwork trains, the canonical
		// optimizer does not run at all -- Adam's step count, the learning-rate decay and the EMA (and with it the switch of the inference
		// weights to the EMA copy) start with the first joint step.  The caller's zero-fills clear the unused gradients.
		int rc_f = 0;
		if (ctx->use_delta) rc_f = delta_optimizer_step(ctx, (cudaStream_t)stream, loss_scale);
		stage_end(ctx, ST_OPT, (cudaStream_t)stream);
		return rc_f;
	}
	++ctx->adam_step;
	const neus2_config& c = ctx->cfg;
	// inside a pipelined train step the kernel also clears what the next step starts from (consumed gradients, the small accumulators,
	// this step's counters): no zero-fill launches are left between the optimizer and the next step's first kernel
	const int fuse = ctx->fuse_zero_into_adam ? 1 : 0;
	uint32_t* const zp[3] = {(uint32_t*)ctx->var_acc, (uint32_t*)ctx->w1row0_acc, ctx->counters};
	const uint32_t zn[3] = {1u, 64u, 8u};
	const float lr = ctx->learning_rate * neus2_learning_rate_factor(&c, ctx->adam_step - 1);   // x 1.0f before decay_start: the same float
	const int rc_ = launch_adam_ema((cudaStream_t)stream, ctx->n_params, ctx->pl.off_grid, loss_scale, lr, c.beta1, c.beta2, c.epsilon, c.l2_reg,
	                                c.ema_decay, ctx->adam_step, ctx->grads, ctx->master, ctx->params, ctx->m1, ctx->m2, ctx->steps, ctx->ema, ctx->adam_bias_table, ADAM_BIAS_TABLE,
	                                fuse, zp, zn);
	if (rc_ == 0 && ctx->use_delta) CHECK(delta_optimizer_step(ctx, (cudaStream_t)stream, loss_scale));   // m_global_move.trainer->optimizer_step (testbed_nerf.cu:3506-3508)
	stage_end(ctx, ST_OPT, (cudaStream_t)stream);
	return rc_;
}

int neus2_set_dataset(neus2_ctx* ctx, uint32_t n_images, int32_t width, int32_t height, const uint8_t* pixels, const float* focal, const float* principal, const float* xforms) {
	invalidate_prefetch(ctx);
	CUDA_CHECK(cudaSetDevice(ctx->device));
	if (!n_images || width <= 0 || height <= 0) FAIL("empty dataset");
	if (ctx->d_pixels) { cudaFree(ctx->d_pixels); cudaFree(ctx->d_focal); cudaFree(ctx->d_principal); cudaFree(ctx->d_xforms); }
	const size_t npx = (size_t)n_images * width * height * 4;
	CUDA_CHECK(cudaMalloc((void**)&ctx->d_pixels, npx));
	CUDA_CHECK(cudaMalloc((void**)&ctx->d_focal, n_images * 2 * sizeof(float)));
	CUDA_CHECK(cudaMalloc((void**)&ctx->d_principal, n_images * 2 * sizeof(float)));
	CUDA_CHECK(cudaMalloc((void**)&ctx->d_xforms, n_images * 12 * sizeof(float)));
	CUDA_CHECK(cudaMemcpy(ctx->d_pixels, pixels, npx, cudaMemcpyHostToDevice));
	CUDA_CHECK(cudaMemcpy(ctx->d_focal, focal, n_images * 2 * sizeof(float), cudaMemcpyHostToDevice));
	CUDA_CHECK(cudaMemcpy(ctx->d_principal, principal, n_images * 2 * sizeof(float), cudaMemcpyHostToDevice));
	CUDA_CHECK(cudaMemcpy(ctx->d_xforms, xforms, n_images * 12 * sizeof(float), cudaMemcpyHostToDevice));
	ctx->ds.n_images = n_images; ctx->ds.width = width; ctx->ds.height = height;
	ctx->ds.pixels = ctx->d_pixels; ctx->ds.focal = ctx->d_focal; ctx->ds.principal = ctx->d_principal; ctx->ds.xforms = ctx->d_xforms;
	return 0;
}

static int update_density_grid_on_stream(neus2_ctx* ctx, cudaStream_t stream, bool reset = false);
int neus2_update_density_grid(neus2_ctx* ctx, void* stream_) {   // training_prep_nerf (:4003-4016) -> update_density_grid_nerf (:3293-3375)
	CUDA_CHECK(cudaSetDevice(ctx->device));
	invalidate_prefetch(ctx);   // the occupancy bitfield changes
	return update_density_grid_on_stream(ctx, (cudaStream_t)stream_);
}
// reset: Testbed::reset_density_grid_nerf (src/testbed_nerf.cu:3205-3291) -- the same update from a zeroed grid with one uniform sample in every cell
static int update_density_grid_on_stream(neus2_ctx* ctx, cudaStream_t stream, bool reset) {
	void* stream_ = (void*)stream; (void)stream_;
	const uint32_t G3n = NERF_GRIDSIZE * NERF_GRIDSIZE * NERF_GRIDSIZE;
	const uint32_t n_elements = G3n * ctx->n_cascades;
	uint32_t n_uniform, n_nonuniform;
	if (reset || canonical_training_step(ctx) < 256) { n_uniform = G3n * ctx->n_cascades; n_nonuniform = 0; }   // training_prep_nerf (:4011)
	else { n_uniform = G3n / 4 * ctx->n_cascades; n_nonuniform = G3n / 4 * ctx->n_cascades; }
	if (reset) {
		CUDA_CHECK(cudaMemsetAsync(ctx->density_grid, 0, sizeof(float) * n_elements, stream));   // the decay counter keeps running (:3249, 3287)
	} else if (ctx->training_step == 0) {
		ctx->density_grid_ema_step = 0;
		CUDA_CHECK(cudaMemsetAsync(ctx->density_grid, 0, sizeof(float) * n_elements, stream));
	}
	uint64_t st[2] = {ctx->density_grid_rng.state, ctx->density_grid_rng.inc};
	const char* ordered_env = getenv("NEUS2_GRID_SAMPLES_ORDERED");   // =0: A/B switch (uniform samples in the reference's sample order)
	const bool ordered = ordered_env ? atoi(ordered_env) != 0 : true;
	ctx->lb_epoch = ctx->lb_epoch % 0x3ffffffeu + 1u;
	CHECK(launch_density_grid_samples(stream, n_uniform, st, ctx->density_grid_ema_step, ctx->aabb, ctx->density_grid, ctx->n_cascades, -0.01f,
	                                  ctx->grid_positions, ctx->grid_indices, ordered ? ctx->lb_grid_desc : nullptr, ctx->lb_tickets + 1, ctx->lb_epoch));
	ctx->density_grid_rng.advance();
	st[0] = ctx->density_grid_rng.state;
	CHECK(launch_density_grid_samples(stream, n_nonuniform, st, ctx->density_grid_ema_step, ctx->aabb, ctx->density_grid, ctx->n_cascades, 0.1f,
	                                  ctx->grid_positions + (size_t)n_uniform * 3, ctx->grid_indices + n_uniform));
	ctx->density_grid_rng.advance();
	const uint32_t n = n_uniform + n_nonuniform;
	// density with the TRAINING params (use_inference_params=false, :3366)
	CHECK(launch_network_sdf(stream, ctx->gm, ctx->pl, ctx->params, ctx->grid_positions, 3, n, ctx->cfg.sdf_bias, 1, ctx->grid_density, nullptr));
	CHECK(launch_density_grid_splat_decay(stream, n, ctx->grid_indices, ctx->grid_density, n_elements, ctx->cfg.density_grid_decay, ctx->density_grid_tmp, ctx->density_grid));
	++ctx->density_grid_ema_step;
	CHECK(launch_density_grid_bitfield(stream, ctx->density_grid, ctx->n_cascades, ctx->density_grid_mean, ctx->bitfield));
	return 0;
}

// Loss scalars of a step (every 16th, testbed.cu:2714): the three per-ray arrays are summed on the device right behind the loss kernel
// (reduce_loss_scalars); a data-parallel caller all-reduces the three partial sums in place (neus2_loss_scalars_device) before the
// read-back (read_back_loss_scalars) is enqueued.
static int reduce_loss_scalars(neus2_ctx* ctx, cudaStream_t stream) {
	// the per-ray arrays hold THIS rank's rays (max_rays = max_rays_per_batch / world_size entries)
	const uint32_t n = std::min(ctx->step_rays_local, ctx->max_rays);
	CHECK(launch_reduce_sum(stream, ctx->loss, n, ctx->scalar_tmp + 0));
	CHECK(launch_reduce_sum(stream, ctx->ek_loss, n, ctx->scalar_tmp + 1));
	CHECK(launch_reduce_sum(stream, ctx->mask_loss, n, ctx->scalar_tmp + 2));
	return 0;
}
static int read_back_loss_scalars(neus2_ctx* ctx, cudaStream_t stream) {
	CUDA_CHECK(cudaMemcpyAsync(ctx->h_scalars, ctx->scalar_tmp, 3 * sizeof(float), cudaMemcpyDeviceToHost, stream));
	return 0;
}

// Where the next step's sample generation may start: 0 = as soon as this step's loss kernel is done (nothing after it reads or writes
// the sample buffers); its blocks then fill the SMs the persistent forward+backward kernel frees at its tail and run on beside the
// weight-gradient GEMMs and Adam.  1 = only after the forward+backward kernel has completed.
#ifndef NEUS2_GEN_AFTER_FWDBWD
#define NEUS2_GEN_AFTER_FWDBWD 0
#endif
// What sample generation of the next train step will be launched with: a pure function of the context's step state, so that the
// end of step k can generate step k+1's samples ahead of time and step k+1 can verify that they are the ones it would have made.
static void plan_step(const neus2_ctx* ctx, uint32_t rank, uint32_t world_size, neus2_ctx::StepPlan& p) {
	memset(&p, 0, sizeof(p));
	p.R = ctx->rays_per_batch;
	p.n_local = p.R / world_size; p.ray_offset = rank; p.ray_stride = world_size;   // rank r traces the global rays r, r + W, r + 2 W, ...
	// :3770-3775: the first pass is sized by the previous step's sample count, and rays whose samples do not fit are dropped (:1422-1424).
	// Data-parallel, measured_before is the GLOBAL count and a rank's budget is its share of it (the scan over rays restarts on every rank).
	const uint32_t before_share = (ctx->measured_before + world_size - 1) / world_size;
	p.max_inference = ctx->measured_before == 0 ? ctx->max_samples : (std::min(before_share, ctx->max_samples) + 127u) / 128u * 128u;
	p.n_rays_total = (ctx->training_step == 0 || canonical_training_step(ctx) == 0) ? 0 : ctx->n_rays_total;   // :3785-3787
	p.training_step = ctx->training_step;
	p.rng_state = ctx->rng.state; p.rng_inc = ctx->rng.inc;
}
static int launch_gen(neus2_ctx* ctx, cudaStream_t stream, const neus2_ctx::StepPlan& p, uint32_t* counters) {
	const uint64_t st[2] = {p.rng_state, p.rng_inc};
	return launch_generate_training_samples(stream, p.n_local, p.ray_offset, p.ray_stride, p.R, p.n_rays_total, st, ctx->ds, ctx->aabb, ctx->gmv, ctx->bitfield, ctx->cfg.snap_to_pixel_centers,
	                                        ctx->cone_angle_constant, p.max_inference, ctx->numsteps_all, ctx->base, ctx->slot, ctx->ray_indices, ctx->rays,
	                                        ctx->numsteps, ctx->coords, counters, ctx->windows, ctx->n_windows, ctx->ray_setup,
	                                        ctx->cfg.forward_early_termination ? ctx->scan_tmp : nullptr);   // scan_tmp: free between the loss's scan of step k and of step k + 1
}
static bool density_grid_due(uint32_t canonical_training_step) {
	const uint32_t n_prep_to_skip = std::min(std::max(canonical_training_step / 16u, 1u), 16u);   // testbed.cu:2669-2670: m_canonical_training_step
	return canonical_training_step % n_prep_to_skip == 0;
}

int neus2_train_step_begin(neus2_ctx* ctx, void* stream_, uint32_t rank, uint32_t world_size) {
	cudaStream_t stream = (cudaStream_t)stream_;
	CUDA_CHECK(cudaSetDevice(ctx->device));
	if (!ctx->d_pixels) FAIL("no training data: call neus2_set_dataset first");
	if (world_size != ctx->cfg.world_size || rank >= world_size) FAIL("rank / world_size do not match the context's config.world_size");
	const neus2_config& c = ctx->cfg;
	// world > 1 without peer mappings: neus2_train_step_counts_ready switches the step to the pipelined end (behind the caller's all-reduce)
	const bool peers = world_size > 1 && ctx->peer_world == world_size;
	const bool pipelined = (world_size == 1 || peers) && c.prefetch_next_samples;
	ctx->ev_set ^= 1;
	neus2_set_training_step(ctx, ctx->time_frame >= 1 ? (int32_t)ctx->training_step - (int32_t)ctx->frozen_steps : (int32_t)ctx->training_step);   // testbed.cu:2651-2657
	ctx->step_density_updated = false;
	stage_begin(ctx, ST_STEP, stream);
	if (c.reset_density_grid_after_global_movement && ctx->time_frame >= 1 && ctx->training_step == ctx->frozen_steps) {
		// the canonical model starts training again (testbed.cu:2659-2666): rebuild the occupancy grid from zero before the regular refresh
		CHECK(update_density_grid_on_stream(ctx, stream, /*reset=*/true));
		ctx->step_density_updated = true;
	}
	if (density_grid_due(canonical_training_step(ctx))) {
		stage_begin(ctx, ST_GRID, stream);
		CHECK(update_density_grid_on_stream(ctx, stream));
		stage_end(ctx, ST_GRID, stream);
		ctx->step_density_updated = true;
	}

	// Counters::prepare_for_training_steps (:3399-3410)
	const bool pre_zeroed = ctx->scratch_zeroed;   // train_step_end of the previous step already queued these fills on this stream
	ctx->scratch_zeroed = false;

	// train_nerf_step (:3723-4001)
	neus2_ctx::StepPlan plan;
	plan_step(ctx, rank, world_size, plan);
	const uint32_t R = plan.R, n_local = plan.n_local, max_inference = plan.max_inference, n_rays_total = plan.n_rays_total;
	const uint32_t rows_local = ctx->B_local;
	bool have_samples = false;
	if (ctx->aux_pending) { CUDA_CHECK(cudaStreamWaitEvent(stream, ctx->ev_gen_done, 0)); ctx->aux_pending = false; }
	if (ctx->prefetched) {
		// step k-1 generated samples ahead of time on the aux stream; they are this step's iff nothing they depend on changed
		ctx->prefetched = false;
		if (c.prefetch_next_samples && pre_zeroed && !ctx->step_density_updated && ctx->prefetch_epoch == ctx->state_epoch && !memcmp(&plan, &ctx->prefetch_plan, sizeof(plan))) {
			ctx->counters_cur ^= 1;
			ctx->counters = ctx->counters_buf[ctx->counters_cur]; ctx->counts_global = ctx->counters + 8;
			have_samples = true;
		}
	}
	ctx->step_used_prefetch = have_samples;
	if (n_local > ctx->max_rays) FAIL("rays_per_batch exceeds max_rays_per_batch");
	if (!have_samples && !pre_zeroed) CHECK(zero_step_scratch(ctx, stream, n_local));
	if (ctx->measured_before == 0) ctx->measured_before = ctx->max_samples;
	ctx->n_rays_total = n_rays_total + R;   // reset first while the canonical step is 0 (:3785-3790)
	ctx->step_rays = R; ctx->step_rays_local = n_local; ctx->step_n_rays_total = n_rays_total;
	const uint64_t st[2] = {plan.rng_state, plan.rng_inc};
	if (!have_samples) {
		stage_begin(ctx, ST_GEN, stream);
		CHECK(launch_gen(ctx, stream, plan, ctx->counters));
		stage_end(ctx, ST_GEN, stream);
	}
	LossParamsDev lp;
	lp.loss_scale = c.loss_scale; memcpy(lp.background_color, c.background_color, sizeof(lp.background_color));
	lp.random_bg_color = c.random_bg_color; lp.snap_to_pixel_centers = c.snap_to_pixel_centers;
	lp.mask_loss_weight = c.mask_loss_weight; lp.ek_loss_weight = c.ek_loss_weight;
	lp.cos_anneal_ratio = c.anneal_end == 0 ? 1.0f : (float)std::min(1.0, (double)((float)ctx->training_step / c.anneal_end));   // nerf_network.h:1119-1128
	stage_begin(ctx, ST_FWD, stream);
	if (c.forward_early_termination) {
		CHECK(launch_network_forward_rays(stream, ctx->gm, ctx->pl, ctx->params, c.sdf_bias, lp.cos_anneal_ratio, ctx->counters + 0, ctx->counters + 6, ctx->numsteps,
		                                  ctx->coords, max_inference, n_local, ctx->mlp_out, ctx->composite_in, ctx->scan_tmp, ctx->counters + 4));   // scan_tmp: free until the loss's scan
	} else {
		CHECK(launch_network_forward(stream, ctx->gm, ctx->pl, ctx->params, ctx->coords, ctx->counters + 1, max_inference, c.sdf_bias, ctx->mlp_out));   // :3854
	}
	stage_end(ctx, ST_FWD, stream);
	stage_begin(ctx, ST_LOSS, stream);
	const char* loss_env = getenv("NEUS2_LOSS_BY_SAMPLES");   // =0: A/B switch (composite / scan / gradient kernels, one warp per ray); read per step
	const bool by_samples = c.forward_early_termination && (loss_env ? atoi(loss_env) != 0 : true);   // needs what the early-terminating first pass stores
	ctx->lb_epoch = ctx->lb_epoch % 0x3ffffffeu + 1u;
	CHECK(launch_compute_loss(stream, n_local, R, n_rays_total, st, ctx->ds, ctx->aabb, ctx->gmv, lp, rows_local, ctx->counters + 0, ctx->mlp_out, ctx->ray_indices,
	                          ctx->rays, ctx->numsteps, ctx->coords, ctx->ray_tmp, ctx->scan_tmp, ctx->coords_compacted, ctx->dloss, ctx->loss, ctx->ek_loss,
	                          ctx->mask_loss, ctx->counters + 2, by_samples ? ctx->composite_in : nullptr, &ctx->loss_ws, ctx->lb_loss_desc, ctx->lb_tickets + 0, ctx->lb_epoch));
	stage_end(ctx, ST_LOSS, stream);
	ctx->step_pipelined = pipelined;
	ctx->step_rank = rank; ctx->step_world = world_size;
	if (pipelined) {
		// every number the host needs from this step exists now: read it back BEFORE the backward pass is queued, so that the host's
		// turnaround (and the next step's sample generation) hides behind the backward pass instead of following it
		const bool loss_due = ctx->training_step % 16 == 0;
		if (loss_due) CHECK(reduce_loss_scalars(ctx, stream));
		if (peers) {
			// data-parallel: the counts (and loss sums) of all ranks meet here, over peer memory, and every rank reads back the same sums
			stage_begin(ctx, ST_X_COUNTS, stream);
			PeerXfer x1; x1.peers = ctx->peer_sync; x1.epoch = ++ctx->exchange_epoch; x1.rank = ctx->peer_rank; x1.world = ctx->peer_world;
			CHECK(launch_peer_exchange(stream, x1, ctx->counters, ctx->counts_global, loss_due ? ctx->scalar_tmp : nullptr, nullptr));
			stage_end(ctx, ST_X_COUNTS, stream);
		}   // one GPU: the global counts are the local ones, read back from counters[0..2] directly (no device-to-device copy)
		// ... on the second stream: a copy queued on the main stream would hold the training kernel back until it has landed in host memory
		static int readback_main = -1;   // NEUS2_READBACK_MAIN=1: A/B switch (copies on the main stream)
		if (readback_main < 0) { const char* e = getenv("NEUS2_READBACK_MAIN"); readback_main = e ? atoi(e) : 0; }
		cudaStream_t rb = readback_main ? stream : ctx->aux_stream;
		if (!readback_main) {
			CUDA_CHECK(cudaEventRecord(ctx->ev_loss_done, stream));
			CUDA_CHECK(cudaStreamWaitEvent(rb, ctx->ev_loss_done, 0));
		}
		if (loss_due) CHECK(read_back_loss_scalars(ctx, rb));
		CUDA_CHECK(cudaMemcpyAsync(ctx->h_counters, ctx->counters, 12 * sizeof(uint32_t), cudaMemcpyDeviceToHost, rb));
		CUDA_CHECK(cudaEventRecord(ctx->ev_counts, rb));
	} else if (ctx->training_step % 16 == 0) {
		CHECK(reduce_loss_scalars(ctx, stream));   // read back by neus2_train_step_counts_ready / neus2_train_step_end, after the caller's all-reduce
	}
	CHECK(backward_on_stream(ctx, stream, ctx->coords_compacted, ctx->dloss, ctx->counters + 2, rows_local, ctx->B, /*grads_zeroed=*/pre_zeroed,
	                         /*fuse_grads_exchange=*/peers && pipelined && weight_gradients_in_kernel()));   // :3941-3946
	if (!pipelined) CUDA_CHECK(cudaMemcpyAsync(ctx->counts_global, ctx->counters, 4 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
	return 0;
}

int neus2_counts_device(neus2_ctx* ctx, uint32_t** counts_dev) {
	if (!counts_dev) FAIL("null argument");
	*counts_dev = ctx->counts_global;
	return 0;
}

// The step's three loss sums {rgb, eikonal, mask, -} (fp32, this rank's rays), valid between neus2_train_step_begin and the read-back on
// steps where training_step % 16 == 0 (*due): a data-parallel caller all-reduces them in place next to the counts.
int neus2_loss_scalars_device(neus2_ctx* ctx, float** sums_dev, int* due) {
	if (!sums_dev) FAIL("null argument");
	*sums_dev = ctx->scalar_tmp;
	if (due) *due = ctx->training_step % 16 == 0 ? 1 : 0;
	return 0;
}

int neus2_local_counts(neus2_ctx* ctx, void* stream_, uint32_t counts[3]) {
	cudaStream_t stream = (cudaStream_t)stream_;
	CUDA_CHECK(cudaMemcpyAsync(ctx->h_counters, ctx->counters, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
	CUDA_CHECK(cudaStreamSynchronize(stream));
	counts[0] = ctx->h_counters[0]; counts[1] = ctx->h_counters[1]; counts[2] = ctx->h_counters[2];
	return 0;
}

// Data-parallel callers: call right after the all-reduce of neus2_counts_device() has been enqueued (it follows the backward pass).
// The read-back of the global counts is taken here, ahead of the optimizer / its collectives, and the step ends like the pipelined
// single-GPU step: no host sync at its end, the next step's samples are generated beside the optimizer.  No-op without
// config.prefetch_next_samples.
int neus2_train_step_counts_ready(neus2_ctx* ctx, void* stream_) {
	cudaStream_t stream = (cudaStream_t)stream_;
	CUDA_CHECK(cudaSetDevice(ctx->device));
	if (!ctx->cfg.prefetch_next_samples || ctx->step_pipelined) return 0;
	if (ctx->training_step % 16 == 0) CHECK(read_back_loss_scalars(ctx, stream));
	CUDA_CHECK(cudaMemcpyAsync(ctx->h_counters, ctx->counters, 12 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
	CUDA_CHECK(cudaEventRecord(ctx->ev_counts, stream));
	ctx->step_pipelined = true;
	return 0;
}

int neus2_train_step_end(neus2_ctx* ctx, void* stream_, const uint32_t* counts_global, neus2_step_stats* stats) {
	cudaStream_t stream = (cudaStream_t)stream_;
	CUDA_CHECK(cudaSetDevice(ctx->device));
	const neus2_config& c = ctx->cfg;
	const bool get_loss_scalar = ctx->training_step % 16 == 0;                           // testbed.cu:2714
	const bool pipelined = ctx->step_pipelined;
	// single-GPU pipelined step: our own Adam kernel runs next and also clears what the next step starts from
	static int no_fused_zero = -1;   // NEUS2_NO_FUSED_ZERO=1: A/B switch (zero-fill launches behind the optimizer instead)
	if (no_fused_zero < 0) { const char* e = getenv("NEUS2_NO_FUSED_ZERO"); no_fused_zero = e ? atoi(e) : 0; }
	// data-parallel: the sharded optimizer + the peer barrier behind it did the same (sharded_zeroed && barrier_zeroed)
	const bool fused_zero = pipelined && !no_fused_zero && !canonical_frozen(ctx) &&
	                        (ctx->optimizer_done ? (ctx->sharded_zeroed && ctx->barrier_zeroed) : c.world_size == 1);
	ctx->sharded_zeroed = ctx->barrier_zeroed = false;
	if (pipelined) CUDA_CHECK(cudaStreamWaitEvent(stream, ctx->ev_counts, 0));   // the optimizer clears the counters the read-back (second stream) copies: long done, but say so
	if (!ctx->optimizer_done) {
		ctx->fuse_zero_into_adam = fused_zero;
		const int rc_opt = neus2_optimizer_step(ctx, stream_, c.loss_scale);   // testbed_nerf.cu:3503-3505
		ctx->fuse_zero_into_adam = false;
		CHECK(rc_opt);
	}
	ctx->optimizer_done = false;
	if (!pipelined) {
		if (get_loss_scalar) CHECK(read_back_loss_scalars(ctx, stream));
		stage_end(ctx, ST_STEP, stream);
		// the one structurally required host sync per step (:3412-3416): local counters + the (possibly all-reduced) global ones
		// counters[0..3] local {kept, N_u, compacted}, [4..7] first-pass work counters, [8..11] the (all-reduced) global counts
		CUDA_CHECK(cudaMemcpyAsync(ctx->h_counters, ctx->counters, 12 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
		if (!ctx->readback_event) CUDA_CHECK(cudaEventCreateWithFlags(&ctx->readback_event, cudaEventDisableTiming));
		CUDA_CHECK(cudaEventRecord(ctx->readback_event, stream));
	}
	// While the host waits for the read-back and prepares the next step, the GPU already zero-fills what the next step starts from
	// (42 MB of gradients + per-ray scratch): stream-ordered behind the optimizer and, data-parallel, behind the peers' barrier.
	if (!fused_zero) {
		CHECK(zero_step_scratch(ctx, stream, ctx->max_rays));
		CUDA_CHECK(cudaMemsetAsync(ctx->grads, 0, ctx->n_params * sizeof(float), stream));
		CUDA_CHECK(cudaMemsetAsync(ctx->var_acc, 0, sizeof(float), stream));
		CUDA_CHECK(cudaMemsetAsync(ctx->w1row0_acc, 0, 64 * sizeof(float), stream));
	}   // else: gradients / accumulators / counters were cleared by the Adam kernel, the per-ray loss arrays are cleared on the aux stream below
	ctx->scratch_zeroed = true;
	if (pipelined) {
		stage_end(ctx, ST_STEP, stream);
		CUDA_CHECK(cudaEventSynchronize(ctx->ev_counts));      // recorded right after the loss kernel: the backward pass is still running
		stage_collect(ctx, ctx->ev_set ^ 1);                   // the previous step's events have all completed by now
	} else {
		CUDA_CHECK(cudaEventSynchronize(ctx->readback_event));
		stage_collect(ctx, 0); stage_collect(ctx, 1);
	}
	const uint32_t n_kept_local = ctx->h_counters[0];
	const int g0 = (pipelined && c.world_size == 1) ? 0 : 8;   // [8..10]: the (all-reduced / exchanged) global counts; one pipelined GPU: the local ones
	uint32_t counts[3] = {ctx->h_counters[g0], ctx->h_counters[g0 + 1], ctx->h_counters[g0 + 2]};
	if (counts_global) { counts[0] = counts_global[0]; counts[1] = counts_global[1]; counts[2] = counts_global[2]; }
	const bool have_global = counts_global != nullptr || ctx->cfg.world_size > 1;
	++ctx->training_step;
	ctx->rng.advance();                                                                  // :3996
	neus2_step_stats s;
	memset(&s, 0, sizeof(s));
	s.training_step = ctx->training_step;
	s.rays_per_batch = ctx->step_rays;
	s.n_rays_kept = have_global ? counts[0] : n_kept_local;
	s.density_grid_updated = ctx->step_density_updated;
	s.samples_prefetched = ctx->step_used_prefetch ? 1 : 0;
	s.first_pass_samples = ctx->cfg.forward_early_termination ? ctx->h_counters[5] : std::min(ctx->h_counters[1], ctx->max_samples);
	// Counters::update_after_training (:3412-3438)
	if (counts[1] == 0 || counts[2] == 0) {
		s.measured_batch_size = 0; s.measured_batch_size_before_compaction = 0;
		ctx->measured_before = 0;
	} else {
		s.measured_batch_size_before_compaction = counts[1];
		s.measured_batch_size = counts[2];
		ctx->measured_before = counts[1];
		if (get_loss_scalar) {
			const float m = (float)counts[2], t = (float)ctx->B;      // sum * measured / target, left to right (:3429-3431)
			s.loss = ctx->h_scalars[0] * m / t; s.ek_loss = ctx->h_scalars[1] * m / t; s.mask_loss = ctx->h_scalars[2] * m / t;
			s.loss_valid = 1;
		}
		ctx->rays_per_batch = neus2_next_rays_per_batch(ctx->rays_per_batch, ctx->B, counts[2], ctx->cfg.world_size, ctx->cfg.max_rays_per_batch);
		if (pipelined && !density_grid_due(canonical_training_step(ctx))) {
			// Generate the NEXT step's samples now, on the aux stream, behind this step's fused forward+backward kernel (which leaves no
			// registers for anyone else) and concurrently with its weight-gradient GEMMs, optimizer and zero-fills.  They land in the
			// other counter buffer; the sample buffers themselves were last read by this step's loss kernel, long done.
			plan_step(ctx, ctx->step_rank, ctx->step_world, ctx->prefetch_plan);
			uint32_t* next_counters = ctx->counters_buf[ctx->counters_cur ^ 1];
			static int gen_after_fwdbwd = -1;   // NEUS2_GEN_AFTER_FWDBWD=1/0 overrides the compile-time choice (A/B switch)
			if (gen_after_fwdbwd < 0) { const char* e = getenv("NEUS2_GEN_AFTER_FWDBWD"); gen_after_fwdbwd = e ? atoi(e) : NEUS2_GEN_AFTER_FWDBWD; }
			CUDA_CHECK(cudaStreamWaitEvent(ctx->aux_stream, gen_after_fwdbwd ? ctx->ev_fwdbwd_done : ctx->ev_counts, 0));
			CUDA_CHECK(cudaMemsetAsync(next_counters, 0, 8 * sizeof(uint32_t), ctx->aux_stream));
			stage_begin_set(ctx, ctx->ev_set ^ 1, ST_GEN, ctx->aux_stream);
			CHECK(launch_gen(ctx, ctx->aux_stream, ctx->prefetch_plan, next_counters));
			stage_end_set(ctx, ctx->ev_set ^ 1, ST_GEN, ctx->aux_stream);
			CUDA_CHECK(cudaEventRecord(ctx->ev_gen_done, ctx->aux_stream));
			ctx->aux_pending = true;
			ctx->prefetched = true; ctx->prefetch_epoch = ctx->state_epoch;
		}
	}
	if (fused_zero) {
		// The per-ray loss arrays were last read by the loss-scalar reduction right behind the loss kernel.  They are cleared on the aux
		// stream, BEHIND the prefetched sample generation: memset kernels in front of it would sit in the queue until the persistent
		// forward+backward kernel frees an SM and hold the generation kernels back with them (measured: +2 % per step).
		CUDA_CHECK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_counts, 0));
		CUDA_CHECK(cudaMemsetAsync(ctx->loss, 0, sizeof(float) * ctx->max_rays, ctx->aux_stream));
		CUDA_CHECK(cudaMemsetAsync(ctx->ek_loss, 0, sizeof(float) * ctx->max_rays, ctx->aux_stream));
		CUDA_CHECK(cudaMemsetAsync(ctx->mask_loss, 0, sizeof(float) * ctx->max_rays, ctx->aux_stream));
		CUDA_CHECK(cudaEventRecord(ctx->ev_gen_done, ctx->aux_stream));
		ctx->aux_pending = true;
	}
	s.rays_per_batch_next = ctx->rays_per_batch;
	if (stats) *stats = s;
	return 0;
}

int neus2_train_step(neus2_ctx* ctx, void* stream, neus2_step_stats* stats) {
	CHECK(neus2_train_step_begin(ctx, stream, 0, 1));
	return neus2_train_step_end(ctx, stream, nullptr, stats);
}

int neus2_debug_buffer(neus2_ctx* ctx, const char* name, void** ptr, uint64_t* n_bytes) {
	struct E { const char* n; void* p; uint64_t b; };
	const uint64_t R = ctx->max_rays, S = ctx->max_samples, B = ctx->B_local;
	const uint64_t G3n = (uint64_t)NERF_GRIDSIZE * NERF_GRIDSIZE * NERF_GRIDSIZE;
	const E table[] = {
		{"numsteps_all", ctx->numsteps_all, R * 4}, {"ray_indices", ctx->ray_indices, R * 4}, {"rays", ctx->rays, R * 24}, {"numsteps", ctx->numsteps, R * 8},
		{"coords", ctx->coords, S * 28}, {"mlp_out", ctx->mlp_out, S * 32}, {"coords_compacted", ctx->coords_compacted, B * 28}, {"dloss", ctx->dloss, B * 32},
		{"loss", ctx->loss, R * 4}, {"ek_loss", ctx->ek_loss, R * 4}, {"mask_loss", ctx->mask_loss, R * 4}, {"counters", ctx->counters, 48},
		{"density_grid", ctx->density_grid, G3n * ctx->n_cascades * 4}, {"bitfield", ctx->bitfield, G3n / 8 * NERF_CASCADES}, {"density_grid_mean", ctx->density_grid_mean, 4},
		{"grid_positions", ctx->grid_positions, G3n * ctx->n_cascades * 12}, {"grid_indices", ctx->grid_indices, G3n * ctx->n_cascades * 4},
		{"grid_density", ctx->grid_density, G3n * ctx->n_cascades * 2},
		{"render_queue", ctx->r_queue, ctx->render_cap * 32}, {"render_counters", ctx->r_counters, 16}, {"render_ray_rgba", ctx->r_ray_rgba, ctx->render_cap * 16},
		{"render_ray_depth", ctx->r_ray_depth, ctx->render_cap * 4}, {"render_ray_samples", ctx->r_ray_samples, ctx->render_cap * 4},
	};
	for (const E& e : table) if (!strcmp(e.n, name)) { if (!e.p) FAIL("buffer not allocated yet"); *ptr = e.p; if (n_bytes) *n_bytes = e.b; return 0; }
	FAIL("unknown buffer name");
}

int neus2_get_state(neus2_ctx* ctx, uint32_t* training_step, uint32_t* rays_per_batch, uint64_t rng_state[2]) {
	if (training_step) *training_step = ctx->training_step;
	if (rays_per_batch) *rays_per_batch = ctx->rays_per_batch;
	if (rng_state) { rng_state[0] = ctx->rng.state; rng_state[1] = ctx->rng.inc; }
	return 0;
}

int neus2_set_state(neus2_ctx* ctx, uint32_t training_step, uint32_t rays_per_batch, uint32_t n_rays_total, uint32_t measured_before) {
	invalidate_prefetch(ctx);
	ctx->training_step = training_step; ctx->rays_per_batch = rays_per_batch; ctx->n_rays_total = n_rays_total; ctx->measured_before = measured_before;
	return 0;
}

int neus2_set_bitfield(neus2_ctx* ctx, const uint8_t* bitfield_host, uint64_t n_bytes) {
	invalidate_prefetch(ctx);
	const uint64_t want = (uint64_t)NERF_GRIDSIZE * NERF_GRIDSIZE * NERF_GRIDSIZE / 8 * NERF_CASCADES;
	if (n_bytes != want) FAIL("bitfield size mismatch");
	CUDA_CHECK(cudaSetDevice(ctx->device));
	CUDA_CHECK(cudaMemcpy(ctx->bitfield, bitfield_host, n_bytes, cudaMemcpyHostToDevice));
	return 0;
}

int neus2_set_profiling(neus2_ctx* ctx, int enable) {
	CUDA_CHECK(cudaSetDevice(ctx->device));
	if (enable && !ctx->ev[0][0]) { for (int k = 0; k < 2; ++k) for (int i = 0; i < 2 * NEUS2_N_STAGES; ++i) CUDA_CHECK(cudaEventCreate(&ctx->ev[k][i])); }
	CUDA_CHECK(cudaDeviceSynchronize());   // events of steps still in flight belong to the old setting
	ctx->profiling = enable;
	for (int i = 0; i < NEUS2_N_STAGES; ++i) { ctx->stage_ms_sum[i] = 0.f; ctx->stage_count[i] = 0; ctx->stage_pending[0][i] = ctx->stage_pending[1][i] = false; }
	return 0;
}

int neus2_get_stage_times(neus2_ctx* ctx, float ms_sum[8], uint32_t count[8]) {
	CUDA_CHECK(cudaSetDevice(ctx->device));
	if (ctx->profiling) { CUDA_CHECK(cudaDeviceSynchronize()); stage_collect(ctx, 0); stage_collect(ctx, 1); }   // the pipelined step reads a set back one step late
	for (int i = 0; i < NEUS2_N_STAGES_BASE; ++i) { ms_sum[i] = ctx->stage_ms_sum[i]; count[i] = ctx->stage_count[i]; }
	return 0;
}

int neus2_get_stage_times_ext(neus2_ctx* ctx, float ms_sum[12], uint32_t count[12]) {
	CUDA_CHECK(cudaSetDevice(ctx->device));
	if (ctx->profiling) { CUDA_CHECK(cudaDeviceSynchronize()); stage_collect(ctx, 0); stage_collect(ctx, 1); }
	for (int i = 0; i < NEUS2_N_STAGES; ++i) { ms_sum[i] = ctx->stage_ms_sum[i]; count[i] = ctx->stage_count[i]; }
	return 0;
}

uint64_t neus2_launch_count(const neus2_ctx* ctx) { (void)ctx; return g_neus_launches; }

int neus2_set_debug_bits(int bits) { g_fwdbwd_debug_bits = bits; return 0; }

// ---- inference render ---------------------------------------------------------------------------------------------------
void neus2_default_render_params(neus2_render_params* p, int32_t width, int32_t height) {
	memset(p, 0, sizeof(*p));
	p->width = width; p->height = height;
	// reset_camera (src/testbed.cu:272-287): fov_axis 1, fov 50.625 deg, camera at 0.5 - 1.5 * view_dir
	const float rel_focal = 1.0f / (2.0f * std::tan(50.625f * 0.5f * 3.14159265358979323846f / 180.0f));   // set_fov / fov_to_focal_length
	p->focal_length[0] = p->focal_length[1] = rel_focal * (float)height;
	p->screen_center[0] = p->screen_center[1] = 0.5f;
	const float cam[12] = {1.f, 0.f, 0.f, 0.5f, 0.f, -1.f, 0.f, 0.5f, 0.f, 0.f, -1.f, 0.5f + 1.5f};
	memcpy(p->camera_matrix, cam, sizeof(cam));
	p->sample_index = 0; p->snap_to_pixel_centers = 1; p->min_transmittance = 0.01f;
	p->background_color[3] = 1.f; p->exposure = 0.f; p->to_srgb = 1; p->use_inference_params = 1;
	p->tile_x0 = 0; p->tile_y0 = 0; p->tile_w = width; p->tile_h = height;
}

static int render_reserve(neus2_ctx* ctx, size_t n) {
	if (n <= ctx->render_cap) return 0;
	void* old[] = {ctx->r_queue, ctx->r_counters, ctx->r_ray_rgba, ctx->r_ray_depth, ctx->r_accumulate, ctx->r_ray_samples};
	for (void* p : old) if (p) cudaFree(p);
	ctx->r_queue = nullptr; ctx->r_counters = nullptr; ctx->r_ray_rgba = ctx->r_ray_depth = ctx->r_accumulate = nullptr; ctx->r_ray_samples = nullptr;
	ctx->render_cap = 0;
	CHECK(dmalloc(&ctx->r_queue, 32 * n));
	DMALLOC(ctx->r_counters, 4); DMALLOC(ctx->r_ray_rgba, 4 * n); DMALLOC(ctx->r_ray_depth, n); DMALLOC(ctx->r_accumulate, 4 * n); DMALLOC(ctx->r_ray_samples, n);
	ctx->render_cap = n;
	return 0;
}

int neus2_render(neus2_ctx* ctx, void* stream_, const neus2_render_params* p, float* rgba_out, float* depth_out) {
	cudaStream_t stream = (cudaStream_t)stream_;
	CUDA_CHECK(cudaSetDevice(ctx->device));
	if (!p || !rgba_out) FAIL("null argument");
	if (p->width <= 0 || p->height <= 0 || p->tile_w < 0 || p->tile_h < 0 || p->tile_x0 < 0 || p->tile_y0 < 0 || p->row_block < 0 ||
	    p->tile_x0 + p->tile_w > p->width) FAIL("render tile outside the image");
	{   // last image row the tile touches
		const int ry = p->tile_h - 1;
		const int last = p->tile_h == 0 ? p->tile_y0 : p->tile_y0 + (p->row_block > 0 ? (ry / p->row_block) * p->row_block * p->row_stride + ry % p->row_block : ry);
		if (p->row_block > 0 && p->row_stride < 1) FAIL("row_stride must be >= 1");
		if (last >= p->height) FAIL("render tile outside the image");
	}
	if (!(p->focal_length[0] > 0.f) || !(p->focal_length[1] > 0.f)) FAIL("focal length must be positive");
	const size_t n = (size_t)p->tile_w * p->tile_h;
	if (n == 0) return 0;
	if (n > 0xffffffffull) FAIL("tile too large");
	CHECK(render_reserve(ctx, n));
	RenderLaunch r;
	r.width = p->width; r.height = p->height;
	r.focal[0] = p->focal_length[0]; r.focal[1] = p->focal_length[1];
	r.screen_center[0] = p->screen_center[0]; r.screen_center[1] = p->screen_center[1];
	memcpy(r.camera, p->camera_matrix, sizeof(r.camera));
	r.sample_index = p->sample_index; r.snap_to_pixel_centers = p->snap_to_pixel_centers; r.min_transmittance = p->min_transmittance;
	memcpy(r.background_color, p->background_color, sizeof(r.background_color));
	r.exposure = p->exposure; r.to_srgb = p->to_srgb;
	r.tile_x0 = p->tile_x0; r.tile_y0 = p->tile_y0; r.tile_w = p->tile_w; r.tile_h = p->tile_h;
	r.row_block = p->row_block; r.row_stride = p->row_stride;
	const neus2_config& c = ctx->cfg;
	const float cos_anneal_ratio = c.anneal_end == 0 ? 1.0f : (float)std::min(1.0, (double)((float)ctx->training_step / c.anneal_end));
	// m_render_aabb == m_aabb unless the user crops (testbed_nerf.cu:3077)
	return launch_render(stream, ctx->gm, ctx->pl, inference_or_training_params(ctx, p->use_inference_params), c.sdf_bias, r, ctx->aabb, ctx->aabb, ctx->gmv, ctx->bitfield,
	                     ctx->cone_angle_constant, cos_anneal_ratio, ctx->r_queue, ctx->r_counters, ctx->r_ray_rgba, ctx->r_ray_depth, ctx->r_ray_samples,
	                     ctx->r_accumulate, rgba_out, depth_out);
}

int neus2_render_stats(neus2_ctx* ctx, void* stream_, uint32_t stats[2]) {
	cudaStream_t stream = (cudaStream_t)stream_;
	if (!ctx->r_counters) FAIL("no render has run on this context");
	CUDA_CHECK(cudaMemcpyAsync(ctx->h_counters + 4, ctx->r_counters, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
	CUDA_CHECK(cudaStreamSynchronize(stream));
	stats[0] = ctx->h_counters[4]; stats[1] = ctx->h_counters[6];
	return 0;
}

// ---- SDF lattice + marching cubes -----------------------------------------------------------------------------------------
int neus2_sdf_lattice(neus2_ctx* ctx, void* stream, const uint32_t res[3], const float aabb[6], uint32_t z0, uint32_t z1, float* out, int use_inference_params) {
	CUDA_CHECK(cudaSetDevice(ctx->device));
	if (!res || !aabb || !out) FAIL("null argument");
	if (!res[0] || !res[1] || !res[2] || z0 > z1 || z1 > res[2]) FAIL("bad lattice extent");
	Aabb bb; for (int k = 0; k < 3; ++k) { bb.mn[k] = aabb[k]; bb.mx[k] = aabb[3 + k]; }
	return launch_sdf_lattice((cudaStream_t)stream, ctx->gm, ctx->pl, inference_or_training_params(ctx, use_inference_params), ctx->cfg.sdf_bias, res, bb, ctx->aabb, z0, z1, out);
}

int neus2_marching_cubes(neus2_ctx* ctx, void* stream_, const uint32_t res[3], const float aabb[6], float thresh, const float* density, uint32_t* n_verts, uint32_t* n_indices) {
	cudaStream_t stream = (cudaStream_t)stream_;
	CUDA_CHECK(cudaSetDevice(ctx->device));
	if (!res || !aabb || !density) FAIL("null argument");
	const uint64_t cells = (uint64_t)res[0] * res[1] * res[2];
	if (cells == 0 || cells > (1ull << 32)) FAIL("lattice must have between 1 and 2^32 points");
	const uint32_t nb = marching_cubes_blocks(res);
	CHECK(regrow(&ctx->mc_block_counts, &ctx->mc_blocks_cap, (size_t)nb * 2));
	if (!ctx->mc_counts) DMALLOC(ctx->mc_counts, 4);
	CHECK(launch_marching_cubes_count(stream, res, thresh, density, ctx->mc_block_counts, ctx->mc_counts));
	CUDA_CHECK(cudaMemcpyAsync(ctx->h_counters + 4, ctx->mc_counts, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
	CUDA_CHECK(cudaStreamSynchronize(stream));   // marching_cubes_gpu sizes its outputs the same way (marching_cubes.cu:812-818)
	ctx->mc_n_verts = ctx->h_counters[4]; ctx->mc_n_indices = ctx->h_counters[5];
	CHECK(regrow(&ctx->mc_vert_prefix, &ctx->mc_cells_cap, (size_t)cells));
	CHECK(regrow(&ctx->mc_verts, &ctx->mc_verts_cap, (size_t)ctx->mc_n_verts * 3));
	CHECK(regrow(&ctx->mc_indices, &ctx->mc_indices_cap, (size_t)ctx->mc_n_indices));
	Aabb bb; for (int k = 0; k < 3; ++k) { bb.mn[k] = aabb[k]; bb.mx[k] = aabb[3 + k]; }
	CHECK(launch_marching_cubes_emit(stream, res, bb, thresh, density, ctx->mc_block_counts, ctx->mc_n_verts, ctx->mc_n_indices, ctx->mc_vert_prefix, ctx->mc_verts, ctx->mc_indices));
	if (n_verts) *n_verts = ctx->mc_n_verts;
	if (n_indices) *n_indices = ctx->mc_n_indices;
	return 0;
}

int neus2_mesh_buffers(neus2_ctx* ctx, float** verts, uint32_t** indices, uint32_t* n_verts, uint32_t* n_indices) {
	if (verts) *verts = ctx->mc_verts;
	if (indices) *indices = ctx->mc_indices;
	if (n_verts) *n_verts = ctx->mc_n_verts;
	if (n_indices) *n_indices = ctx->mc_n_indices;
	return 0;
}

// ---- data-parallel peers: fused gradient reduction + Adam over NVLink peer memory ------------------------------------------
int neus2_ipc_export(neus2_ctx* ctx, unsigned char* handles) {
	CUDA_CHECK(cudaSetDevice(ctx->device));
	if (!handles) FAIL("null argument");
	static_assert(sizeof(cudaIpcMemHandle_t) == NEUS2_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
	void* bufs[NEUS2_IPC_BUFFERS] = {ctx->grads, ctx->params, ctx->ema, ctx->sync_block};
	for (int k = 0; k < NEUS2_IPC_BUFFERS; ++k) {
		cudaIpcMemHandle_t h;
		CUDA_CHECK(cudaIpcGetMemHandle(&h, bufs[k]));
		memcpy(handles + (size_t)k * NEUS2_IPC_HANDLE_BYTES, &h, NEUS2_IPC_HANDLE_BYTES);
	}
	return 0;
}

int neus2_ipc_attach(neus2_ctx* ctx, uint32_t rank, uint32_t world, const unsigned char* all_handles) {
	CUDA_CHECK(cudaSetDevice(ctx->device));
	if (!all_handles) FAIL("null argument");
	if (world != ctx->cfg.world_size || rank >= world) FAIL("rank / world_size do not match the context's config.world_size");
	if (world > NEUS2_MAX_PEERS) FAIL("at most 8 peers (one NVSwitch domain)");
	if (ctx->peer_world) FAIL("peers are already attached");
	for (uint32_t r = 0; r < world; ++r) {
		if (r == rank) { ctx->peer_grads[r] = ctx->grads; ctx->peer_params[r] = ctx->params; ctx->peer_ema[r] = ctx->ema; ctx->peer_sync.block[r] = ctx->sync_block; continue; }
		void* ptrs[NEUS2_IPC_BUFFERS];
		for (int k = 0; k < NEUS2_IPC_BUFFERS; ++k) {
			cudaIpcMemHandle_t h;
			memcpy(&h, all_handles + ((size_t)r * NEUS2_IPC_BUFFERS + k) * NEUS2_IPC_HANDLE_BYTES, NEUS2_IPC_HANDLE_BYTES);
			CUDA_CHECK(cudaIpcOpenMemHandle(&ptrs[k], h, cudaIpcMemLazyEnablePeerAccess));
		}
		ctx->peer_grads[r] = (float*)ptrs[0]; ctx->peer_params[r] = (__half*)ptrs[1]; ctx->peer_ema[r] = (__half*)ptrs[2]; ctx->peer_sync.block[r] = (PeerSyncBlock*)ptrs[3];
	}
	ctx->peer_world = world; ctx->peer_rank = rank;
	ctx->exchange_epoch = ctx->barrier_epoch = 0;
	return 0;
}

int neus2_optimizer_shard(const neus2_ctx* ctx, uint32_t rank, uint32_t world, uint64_t* block_params, uint64_t* n_owned) {
	if (!world || rank >= world || !block_params || !n_owned) FAIL("bad argument");
	const uint64_t bp = 4ull * OPTIMIZER_BLOCK_QUADS, n = ctx->n_params;
	uint64_t owned = 0;
	for (uint64_t b = rank; b * bp < n; b += world) owned += std::min(bp, n - b * bp);
	*block_params = bp; *n_owned = owned;
	return 0;
}

int neus2_optimizer_step_sharded(neus2_ctx* ctx, void* stream, float loss_scale) {
	CUDA_CHECK(cudaSetDevice(ctx->device));
	if (!ctx->peer_world) FAIL("no peers attached: call neus2_ipc_attach first");
	const neus2_config& c = ctx->cfg;
	stage_begin(ctx, ST_OPT, (cudaStream_t)stream);
	if (canonical_frozen(ctx)) {   // as in neus2_optimizer_step: the canonical optimizer is skipped, every rank steps its DeltaNetwork replica
		int rc_f = 0;
		if (ctx->use_delta) rc_f = delta_optimizer_step(ctx, (cudaStream_t)stream, loss_scale);
		stage_end(ctx, ST_OPT, (cudaStream_t)stream);
		ctx->optimizer_done = true;
		return rc_f;
	}
	++ctx->adam_step;
	const float lr = ctx->learning_rate * neus2_learning_rate_factor(&c, ctx->adam_step - 1);
	// inside a pipelined data-parallel step (neus2_train_step_counts_ready has been called) the kernel also clears the gradient quads it
	// consumed, on every rank: with the peer barrier's small zero-fills no memset is left between the optimizer and the next step
	static int no_fused_zero = -1;
	if (no_fused_zero < 0) { const char* e = getenv("NEUS2_NO_FUSED_ZERO"); no_fused_zero = e ? atoi(e) : 0; }
	const int zero_grads = ctx->step_pipelined && !no_fused_zero ? 1 : 0;
	// NEUS2_FUSED_BARRIER=1: the barrier behind the optimizer ("every slice is stored everywhere") rides in the kernel's last block, together
	// with the zero-fills of the step's small arrays.  Off by default -- measured slower (2 GPUs: optimizer + barrier 0.10 -> 0.20 ms): every
	// block then needs a system-scope fence behind its peer stores before it may count itself as arrived, i.e. it sits on its SM until its
	// NVLink writes are acknowledged; a kernel boundary flushes them for free, and a ONE-warp barrier kernel behind it still finds a slot next
	// to the prefetched sample generation (whose blocks fill the register file, not the thread slots)
	static int fuse_barrier = -1;
	if (fuse_barrier < 0) { const char* e = getenv("NEUS2_FUSED_BARRIER"); fuse_barrier = e ? atoi(e) : 0; }
	PeerXfer xb; xb.peers = ctx->peer_sync; xb.epoch = ++ctx->barrier_epoch; xb.rank = ctx->peer_rank; xb.world = ctx->peer_world;
	PeerZero z = {};
	if (zero_grads) { z.p[0] = (uint32_t*)ctx->var_acc; z.n[0] = 1u; z.p[1] = (uint32_t*)ctx->w1row0_acc; z.n[1] = 64u; z.p[2] = ctx->counters; z.n[2] = 8u; }
	const int rc_ = launch_adam_ema_p2p((cudaStream_t)stream, ctx->n_params, ctx->pl.off_grid, ctx->peer_rank, ctx->peer_world, loss_scale, lr, c.beta1, c.beta2,
	                                    c.epsilon, c.l2_reg, c.ema_decay, ctx->adam_step, ctx->peer_grads, ctx->peer_params, ctx->peer_ema, ctx->master, ctx->m1, ctx->m2, ctx->steps, ctx->adam_bias_table, ADAM_BIAS_TABLE, zero_grads,
	                                    fuse_barrier ? &xb : nullptr, fuse_barrier ? &z : nullptr, ctx->sync_done);
	if (!fuse_barrier) --ctx->barrier_epoch;   // neus2_peer_barrier launches the one-warp kernel with this epoch
	ctx->barrier_fused = rc_ == 0 && fuse_barrier;
	ctx->sharded_zeroed = rc_ == 0 && zero_grads;
	ctx->barrier_zeroed = ctx->sharded_zeroed && fuse_barrier;
	if (rc_ == 0 && ctx->use_delta) CHECK(delta_optimizer_step(ctx, (cudaStream_t)stream, loss_scale));   // every rank steps its replica on the all-reduced sums
	stage_end(ctx, ST_OPT, (cudaStream_t)stream);
	ctx->optimizer_done = true;
	return rc_;
}

// The second exchange of a data-parallel step with peer mappings (the first one, counts + loss sums, sits behind the loss kernel inside
// neus2_train_step_begin): call behind neus2_train_step_begin.  Orders every rank's backward pass before anything that follows on the
// stream and leaves, on every rank, the sum over ranks (in rank order: identical bits everywhere) of the DeltaNetwork's gradient sums
// while global movement trains.  All ranks must make the same sequence of calls.  If the step was begun without the pipelined path
// (config.prefetch_next_samples = 0) it also carries the counts and loss sums.
int neus2_peer_exchange(neus2_ctx* ctx, void* stream) {
	CUDA_CHECK(cudaSetDevice(ctx->device));
	if (!ctx->peer_world) FAIL("no peers attached: call neus2_ipc_attach first");
	if (ctx->grads_exchange_fused) { ctx->grads_exchange_fused = false; return 0; }   // the training kernel's last CTA has done it
	const bool carry_counts = !ctx->step_pipelined;
	stage_begin(ctx, ST_X_GRADS, (cudaStream_t)stream);
	PeerXfer x2; x2.peers = ctx->peer_sync; x2.epoch = ++ctx->exchange_epoch; x2.rank = ctx->peer_rank; x2.world = ctx->peer_world;
	const int rc_ = launch_peer_exchange((cudaStream_t)stream, x2, carry_counts ? ctx->counters : nullptr,
	                                     ctx->counts_global, carry_counts && ctx->training_step % 16 == 0 ? ctx->scalar_tmp : nullptr, ctx->use_delta ? ctx->d_acc : nullptr);
	stage_end(ctx, ST_X_GRADS, (cudaStream_t)stream);
	return rc_;
}

// Barrier over the peers, stream-ordered: everything the peers enqueued before THEIR call has completed (and its peer stores have landed)
// before anything enqueued after this call starts.  Behind a sharded optimizer step that cleared its gradients it also clears the step's
// small accumulators and counters.
int neus2_peer_barrier(neus2_ctx* ctx, void* stream) {
	CUDA_CHECK(cudaSetDevice(ctx->device));
	if (!ctx->peer_world) FAIL("no peers attached: call neus2_ipc_attach first");
	if (ctx->barrier_fused) { ctx->barrier_fused = false; return 0; }   // the sharded optimizer's last block has done it (and cleared the small arrays)
	PeerZero z = {};
	if (ctx->sharded_zeroed) {
		z.p[0] = (uint32_t*)ctx->var_acc; z.n[0] = 1u; z.p[1] = (uint32_t*)ctx->w1row0_acc; z.n[1] = 64u; z.p[2] = ctx->counters; z.n[2] = 8u;
		ctx->barrier_zeroed = true;
	}
	stage_begin(ctx, ST_BARRIER, (cudaStream_t)stream);
	PeerXfer xb; xb.peers = ctx->peer_sync; xb.epoch = ++ctx->barrier_epoch; xb.rank = ctx->peer_rank; xb.world = ctx->peer_world;
	const int rc_ = launch_peer_barrier((cudaStream_t)stream, xb, z);
	stage_end(ctx, ST_BARRIER, (cudaStream_t)stream);
	return rc_;
}

// ---- snapshot support (Testbed::save_snapshot / load_snapshot, src/testbed.cu:3144-3266; Trainer / Adam / EMA serialize) ----
int neus2_optimizer_buffers(neus2_ctx* ctx, float** m1_dev, float** m2_dev, uint32_t** steps_dev, uint32_t* current_step) {
	if (m1_dev) *m1_dev = ctx->m1;
	if (m2_dev) *m2_dev = ctx->m2;
	if (steps_dev) *steps_dev = ctx->steps;
	if (current_step) *current_step = ctx->adam_step;
	return 0;
}

int neus2_set_optimizer_step(neus2_ctx* ctx, uint32_t current_step) {
	CUDA_CHECK(cudaSetDevice(ctx->device));
	ctx->adam_step = current_step;
	CHECK(launch_adam_bias_table(nullptr, ctx->cfg.beta1, ctx->cfg.beta2, 1, current_step, ctx->adam_bias_table, ADAM_BIAS_TABLE));
	CUDA_CHECK(cudaDeviceSynchronize());
	return 0;
}

// density_grid: n floats (128^3 per cascade); recomputes mean + bitfield like load_snapshot does (testbed.cu:3232-3236)
int neus2_set_density_grid(neus2_ctx* ctx, void* stream_, const float* density_grid_host, uint64_t n, uint32_t ema_step) {
	invalidate_prefetch(ctx);
	cudaStream_t stream = (cudaStream_t)stream_;
	CUDA_CHECK(cudaSetDevice(ctx->device));
	const uint64_t want = (uint64_t)NERF_GRIDSIZE * NERF_GRIDSIZE * NERF_GRIDSIZE * ctx->n_cascades;
	if (n != want) FAIL("Incompatible number of grid cascades.");
	CUDA_CHECK(cudaMemcpyAsync(ctx->density_grid, density_grid_host, n * sizeof(float), cudaMemcpyHostToDevice, stream));
	ctx->density_grid_ema_step = ema_step;
	CHECK(launch_density_grid_bitfield(stream, ctx->density_grid, ctx->n_cascades, ctx->density_grid_mean, ctx->bitfield));
	CUDA_CHECK(cudaStreamSynchronize(stream));
	return 0;
}

// compute_mesh_1ring's vertex normals + compute_mesh_vertex_colors (src/marching_cubes.cu:333-370,716-722; src/testbed_nerf.cu:4071-4093)
// for the mesh of the last neus2_marching_cubes call.  normals_out / colors_out: [n_verts][3] float (either may be NULL).
int neus2_mesh_attributes(neus2_ctx* ctx, void* stream_, float* normals_out, float* colors_out, int use_inference_params) {
	cudaStream_t stream = (cudaStream_t)stream_;
	CUDA_CHECK(cudaSetDevice(ctx->device));
	const uint32_t nv = ctx->mc_n_verts, nt = ctx->mc_n_indices / 3;
	if (!nv) return 0;
	if (normals_out) CHECK(launch_mesh_normals(stream, nv, nt, ctx->mc_indices, ctx->mc_verts, normals_out));
	if (colors_out) {
		float* coords = nullptr; __half* out16 = nullptr;
		CUDA_CHECK(cudaMallocAsync((void**)&coords, (size_t)nv * 7 * sizeof(float), stream));
		CUDA_CHECK(cudaMallocAsync((void**)&out16, (size_t)nv * 16 * sizeof(__half), stream));
		int rc = launch_mesh_color_inputs(stream, nv, ctx->mc_verts, ctx->aabb, coords);
		if (!rc) rc = launch_network_forward(stream, ctx->gm, ctx->pl, inference_or_training_params(ctx, use_inference_params), coords, nullptr, nv, ctx->cfg.sdf_bias, out16);
		if (!rc) rc = launch_mesh_extract_colors(stream, nv, out16, 0 /* train_in_linear_colors = false */, colors_out);
		cudaFreeAsync(coords, stream); cudaFreeAsync(out16, stream);
		if (rc) return rc;
	}
	return 0;
}

int neus2_get_counters(neus2_ctx* ctx, uint32_t out[6]) {
	out[0] = ctx->n_rays_total; out[1] = ctx->measured_before; out[2] = ctx->density_grid_ema_step; out[3] = ctx->adam_step;
	out[4] = ctx->measured_after; out[5] = canonical_training_step(ctx);
	return 0;
}

// Testbed::training_network_next_frame (src/testbed.cu:2001-2082) without global-movement prediction: the next frame's images have
// been installed with neus2_set_dataset; the model continues from its INFERENCE (EMA) weights (save_snapshot_incremental /
// load_snapshot_incremental round trip, :2059-2060 -> Trainer::deserialize, tcnn trainer.h:295-304), with a fresh optimizer
// (:2043), the frame's step counters at zero, rays_per_batch back at 4096 and m_rng re-seeded (reset_network_incremental,
// :2351-2361).  For the first `frozen_steps` steps of the new frame the canonical parameters do not move (m_train_canonical =
// false until m_training_step == m_predict_global_movement_training_step, :2659-2667) and the progressive-level schedule counts
// from there; `after_learning_rate` replaces the learning rate (:2688-2690).
int neus2_next_frame(neus2_ctx* ctx, void* stream_, float after_learning_rate, uint32_t frozen_steps) {
	invalidate_prefetch(ctx);
	cudaStream_t stream = (cudaStream_t)stream_;
	CUDA_CHECK(cudaSetDevice(ctx->device));
	const uint64_t n = ctx->n_params;
	if (ctx->adam_step > 0) {
		CUDA_CHECK(cudaMemcpyAsync(ctx->params, ctx->ema, n * sizeof(__half), cudaMemcpyDeviceToDevice, stream));
		k_widen_params<<<(uint32_t)((n + 255) / 256), 256, 0, stream>>>(n, ctx->params, ctx->master);
		COUNT_LAUNCH(1);
		CUDA_CHECK(cudaGetLastError());
	}
	CUDA_CHECK(cudaMemsetAsync(ctx->m1, 0, n * sizeof(float), stream));
	CUDA_CHECK(cudaMemsetAsync(ctx->m2, 0, n * sizeof(float), stream));
	CUDA_CHECK(cudaMemsetAsync(ctx->steps, 0, n * sizeof(uint32_t), stream));
	CUDA_CHECK(cudaMemsetAsync(ctx->ema, 0, n * sizeof(__half), stream));
	CUDA_CHECK(cudaStreamSynchronize(stream));
	ctx->adam_step = 0;
	ctx->time_frame += 1;
	ctx->frozen_steps = frozen_steps;
	if (after_learning_rate > 0.f) ctx->learning_rate = after_learning_rate;
	ctx->training_step = 0;
	ctx->n_rays_total = 0;
	ctx->rays_per_batch = std::min(1u << 12, ctx->cfg.max_rays_per_batch);   // testbed.h:606, within what the per-ray arrays hold
	ctx->measured_before = 0;
	ctx->rng = Pcg32(ctx->cfg.seed);
	return 0;
}

// ---- global rigid motion of dynamic scenes ("predict_global_movement", src/testbed.cu:2054, 2076-2078, 2659-2667, 2704-2710) ----
// Start the new frame's motion estimate: fold the delta trained on the previous frame into the accumulated motion
// (NerfNetwork::accumulate_global_movement, nerf_network.h:1163-1178; accumulate_global_movement_rotation_6d_kernel,
// common_operation.cuh:551-580: R_acc <- R_d R_acc, t_acc <- R_d (t_acc + t_d), stored in half), reset the delta to the identity
// with a fresh optimizer, and route every network evaluation through it (m_use_delta = m_train_delta = true).
int neus2_begin_global_movement(neus2_ctx* ctx, void* stream_) {
	invalidate_prefetch(ctx);
	cudaStream_t stream = (cudaStream_t)stream_;
	CUDA_CHECK(cudaSetDevice(ctx->device));
	float x[DELTA_XFORM_FLOATS];
	CUDA_CHECK(cudaMemcpyAsync(x, ctx->d_xform, sizeof(x), cudaMemcpyDeviceToHost, stream));
	CUDA_CHECK(cudaStreamSynchronize(stream));
	const float* R = x + DELTA_R; const float* t = x + DELTA_T;
	float after[9], sum[3];
	for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
		after[i * 3 + j] = fmaf(R[i * 3 + 2], ctx->gmv.R[6 + j], fmaf(R[i * 3 + 1], ctx->gmv.R[3 + j], R[i * 3] * ctx->gmv.R[j]));
	for (int k = 0; k < 3; ++k) sum[k] = ctx->gmv.T[k] + t[k];
	for (int k = 0; k < 9; ++k) ctx->gmv.R[k] = __half2float(__float2half_rn(after[k]));
	for (int k = 0; k < 3; ++k) ctx->gmv.T[k] = __half2float(__float2half_rn(fmaf(R[k * 3 + 2], sum[2], fmaf(R[k * 3 + 1], sum[1], R[k * 3] * sum[0]))));
	CHECK(reset_delta(ctx, stream));
	ctx->use_delta = true;
	ctx->gm.delta_xform = ctx->d_xform;
	return 0;
}

// m_use_delta / m_train_delta (they are equal whenever the frame index is >= 1).  Switching off keeps the delta parameters: the next
// neus2_begin_global_movement still accumulates them ("finetune_global_movement": false drops the delta from the network after
// predict_global_movement_training_step steps, testbed.cu:2659-2663).
int neus2_set_delta_mode(neus2_ctx* ctx, int use_delta) {
	ctx->use_delta = use_delta != 0;
	ctx->gm.delta_xform = ctx->use_delta ? ctx->d_xform : nullptr;
	return 0;
}

// Snapshot / inspection access.  accumulated_*: what ray generation applies (values as stored in half); delta: the 6d rotation and the
// transition of the current frame's DeltaNetwork (fp32 master copy).  Null pointers are skipped.
int neus2_get_global_movement(neus2_ctx* ctx, void* stream_, float accumulated_rotation[9], float accumulated_transition[3], float delta_rotation_6d[6],
                              float delta_transition[3], int* use_delta) {
	cudaStream_t stream = (cudaStream_t)stream_;
	CUDA_CHECK(cudaSetDevice(ctx->device));
	if (accumulated_rotation) memcpy(accumulated_rotation, ctx->gmv.R, 9 * sizeof(float));
	if (accumulated_transition) memcpy(accumulated_transition, ctx->gmv.T, 3 * sizeof(float));
	if (delta_rotation_6d || delta_transition) {
		float m[12];
		CUDA_CHECK(cudaMemcpyAsync(m, ctx->d_master, sizeof(m), cudaMemcpyDeviceToHost, stream));
		CUDA_CHECK(cudaStreamSynchronize(stream));
		if (delta_transition) memcpy(delta_transition, m, 3 * sizeof(float));
		if (delta_rotation_6d) memcpy(delta_rotation_6d, m + 4, 6 * sizeof(float));
	}
	if (use_delta) *use_delta = ctx->use_delta ? 1 : 0;
	return 0;
}

int neus2_set_global_movement(neus2_ctx* ctx, void* stream_, const float accumulated_rotation[9], const float accumulated_transition[3],
                              const float delta_rotation_6d[6], const float delta_transition[3], const float first_frame_offset[3]) {
	invalidate_prefetch(ctx);
	cudaStream_t stream = (cudaStream_t)stream_;
	CUDA_CHECK(cudaSetDevice(ctx->device));
	if (accumulated_rotation) for (int k = 0; k < 9; ++k) ctx->gmv.R[k] = __half2float(__float2half_rn(accumulated_rotation[k]));
	if (accumulated_transition) for (int k = 0; k < 3; ++k) ctx->gmv.T[k] = __half2float(__float2half_rn(accumulated_transition[k]));
	if (first_frame_offset) memcpy(ctx->gmv.first_frame_offset, first_frame_offset, 3 * sizeof(float));
	if (delta_rotation_6d || delta_transition) {
		float m[12]; __half h[12];
		CUDA_CHECK(cudaMemcpyAsync(m, ctx->d_master, sizeof(m), cudaMemcpyDeviceToHost, stream));
		CUDA_CHECK(cudaStreamSynchronize(stream));
		if (delta_transition) memcpy(m, delta_transition, 3 * sizeof(float));
		if (delta_rotation_6d) memcpy(m + 4, delta_rotation_6d, 6 * sizeof(float));
		for (int k = 0; k < 12; ++k) h[k] = __float2half_rn(m[k]);
		CUDA_CHECK(cudaMemcpyAsync(ctx->d_master, m, sizeof(m), cudaMemcpyHostToDevice, stream));
		CUDA_CHECK(cudaMemcpyAsync(ctx->d_params, h, sizeof(h), cudaMemcpyHostToDevice, stream));
		CUDA_CHECK(cudaStreamSynchronize(stream));
	}
	return launch_delta_update(ctx, stream, 0, 1.f, 0.f);
}

// The 9 gradient sums of the DeltaNetwork ([transition 3 | rotation-6d 6], fp32, 12 slots): a data-parallel caller all-reduces them in
// place between the backward pass and the optimizer step.
int neus2_delta_gradients_device(neus2_ctx* ctx, float** acc_dev) {
	if (!acc_dev) FAIL("null argument");
	*acc_dev = ctx->d_acc;
	return 0;
}

int neus2_get_time_frame(const neus2_ctx* ctx, uint32_t* time_frame, float* learning_rate) {
	if (time_frame) *time_frame = ctx->time_frame;
	if (learning_rate) *learning_rate = ctx->learning_rate;
	return 0;
}

}  // extern "C"
