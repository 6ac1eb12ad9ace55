// ref_shim.cu -- thin C ABI over the REFERENCE's own tiny-cuda-nn fork (dependencies/my_tcnn),
// compiled against the reference sources in place by oracle/Makefile.ref into oracle/_ref/libtcnn_ref.so.
//
// TEST INFRASTRUCTURE ONLY.  This is the "real reference" half of the oracle: it runs the reference's
// hash-grid and fused-MLP kernels (forward, backward, second-order backward) on the GPU box so that
// tests/test_ref_kernels.py can pin both the CPU restatement (oracle/neus_oracle.cpp) and the product
// kernels against them, and bench.py can time them as the reference-CUDA baseline.  Nothing in the
// product links or loads this library.  No reference source is copied: only its public C++ API is called
// (encoding.h: create_encoding, network.h: create_network, object.h:96-303 DifferentiableObject).
#include <tiny-cuda-nn/common.h>
#include <tiny-cuda-nn/encoding.h>
#include <tiny-cuda-nn/network.h>
#include <tiny-cuda-nn/gpu_matrix.h>
#include <tiny-cuda-nn/gpu_memory.h>

#include <cstdint>
#include <cstdio>
#include <memory>
#include <string>

#include <tiny-cuda-nn/optimizer.h>
#include <tiny-cuda-nn/random.h>
#include <tiny-cuda-nn/common_device.h>
#include <pcg32/pcg32.h>

using namespace tcnn;
using T = __half;

static thread_local std::string g_err;

struct RefGrid { std::shared_ptr<Encoding<T>> enc; };
struct RefMlp { std::shared_ptr<Network<T>> net; };

#define REF_TRY try {
#define REF_CATCH } catch (const std::exception& e) { g_err = e.what(); fprintf(stderr, "ref_shim: %s\n", e.what()); return -1; }

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void* ref_grid_create(const char* cfg_json) {
	try {
		json cfg = json::parse(cfg_json);
		auto* g = new RefGrid;
		// alignment 2 as in NerfNetwork (nerf_network.h:59)
		g->enc.reset(create_encoding<T>(3, cfg, 2));
		return g;
	} catch (const std::exception& e) { g_err = e.what(); fprintf(stderr, "ref_shim: %s\n", e.what()); return nullptr; }
}
uint64_t ref_grid_n_params(void* h) { return ((RefGrid*)h)->enc->n_params(); }
uint32_t ref_grid_out_width(void* h) { return ((RefGrid*)h)->enc->padded_output_width(); }
void ref_grid_set_training_step(void* h, int step) { ((RefGrid*)h)->enc->set_training_step(step); }
void ref_grid_destroy(void* h) { delete (RefGrid*)h; }

// Runs, on `iters` repetitions (timed as a whole with CUDA events when ms_out != null):
//   forward (with dy_dx)                          -> enc_soa [E][n] half (row-major)
//   backward(dL_dy_soa)  if dL_dy_soa             -> grid_grad (Overwrite, half [n_params]), dL_dx [n][3] float
//   backward_backward_input(v, dL_dy2_soa) if v   -> grid_grad (Accumulate when dL_dy_soa also given, else Overwrite),
//                                                    dL_ddLdy_soa [E][n] half
// coords: device float [n][stride], stride >= 3.  `which` bit0 fwd only timing, bit1 bwd, bit2 bwdbwd (for timing splits).
int ref_grid_eval(void* h, uint32_t n, const float* coords, uint32_t stride, T* params, T* enc_soa,
                  const T* dL_dy_soa, float* dL_dx, T* grid_grad,
                  const float* v, const T* dL_dy2_soa, T* dL_ddLdy_soa, int iters, int which, float* ms_out) {
	REF_TRY
	auto& enc = ((RefGrid*)h)->enc;
	const uint32_t E = enc->padded_output_width();
	enc->set_params(params, params, nullptr, grid_grad);
	cudaStream_t stream = nullptr;
	GPUMatrixDynamic<float> full((float*)coords, stride, n, CM);
	GPUMatrixDynamic<float> pos = full.slice_rows(0, 3);
	GPUMatrixDynamic<T> out(enc_soa, E, n, RM);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	if (iters < 1) iters = 1;
	std::unique_ptr<Context> ctx;
	if (!(which & 1)) ctx = enc->forward(stream, pos, &out, false, true);
	cudaEventRecord(e0, stream);
	for (int it = 0; it < iters; ++it) {
		if (which & 1) ctx = enc->forward(stream, pos, &out, false, true);
		if ((which & 2) && dL_dy_soa) {
			GPUMatrixDynamic<T> dy((T*)dL_dy_soa, E, n, RM);
			GPUMatrixDynamic<float> dx(dL_dx, 3, n, CM);
			enc->backward(stream, *ctx, pos, out, dy, dL_dx ? &dx : nullptr, false, grid_grad ? EGradientMode::Overwrite : EGradientMode::Ignore);
		}
		if ((which & 4) && v) {
			GPUMatrixDynamic<float> vm((float*)v, 3, n, CM);
			GPUMatrixDynamic<T> dy2((T*)dL_dy2_soa, E, n, RM);
			GPUMatrixDynamic<T> ddy(dL_ddLdy_soa, E, n, RM);
			enc->backward_backward_input(stream, *ctx, pos, vm, dy2, dL_ddLdy_soa ? &ddy : nullptr, nullptr, false,
				grid_grad ? (((which & 2) && dL_dy_soa) ? EGradientMode::Accumulate : EGradientMode::Overwrite) : EGradientMode::Ignore);
		}
	}
	cudaEventRecord(e1, stream);
	CUDA_CHECK_THROW(cudaStreamSynchronize(stream));
	float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
	if (ms_out) *ms_out = ms / iters;
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	ctx.reset();
	return 0;
	REF_CATCH
}

void* ref_mlp_create(const char* cfg_json) {
	try {
		json cfg = json::parse(cfg_json);
		auto* m = new RefMlp;
		m->net.reset(create_network<T>(cfg));
		return m;
	} catch (const std::exception& e) { g_err = e.what(); fprintf(stderr, "ref_shim: %s\n", e.what()); return nullptr; }
}
uint64_t ref_mlp_n_params(void* h) { return ((RefMlp*)h)->net->n_params(); }
void ref_mlp_destroy(void* h) { delete (RefMlp*)h; }

// in: half, [in_w][n] row-major when in_soa!=0 (how NerfNetwork feeds the SDF MLP) else [n][in_w].
// out / dL_dout / seed: [n][16] (column-major 16 x n).  dL_din: same layout as `in`.  d_in2: ALWAYS [n][in_w]
// (fully_fused_mlp.cu:1119 wraps dL_ddLdinput in a column-major GPUMatrix whatever its declared layout; NerfNetwork
// allocates d_density_input column-major, nerf_network.h:560).
//   forward                                   -> out
//   backward(dL_dout) if dL_dout              -> dL_din (optional), wgrad (Overwrite, half [n_params]) (optional)
//   backward_backward_input(d_in2, seed)      -> wgrad (Accumulate if the first backward ran with wgrad, else Overwrite)
int ref_mlp_eval(void* h, uint32_t n, uint32_t in_w, int in_soa, const T* in, T* params, T* out,
                 const T* dL_dout, T* dL_din, T* wgrad, const T* d_in2, const T* seed, int iters, int which, float* ms_out) {
	REF_TRY
	auto& net = ((RefMlp*)h)->net;
	net->set_params(params, params, params, wgrad);
	cudaStream_t stream = nullptr;
	const MatrixLayout lay = in_soa ? RM : CM;
	GPUMatrixDynamic<T> inm((T*)in, in_w, n, lay);
	GPUMatrixDynamic<T> outm(out, net->padded_output_width(), n, CM);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	if (iters < 1) iters = 1;
	std::unique_ptr<Context> ctx;
	if (!(which & 1)) ctx = net->forward(stream, inm, &outm, false, true);
	cudaEventRecord(e0, stream);
	for (int it = 0; it < iters; ++it) {
		if (which & 1) ctx = net->forward(stream, inm, &outm, false, true);
		if ((which & 2) && dL_dout) {
			GPUMatrixDynamic<T> dy((T*)dL_dout, net->padded_output_width(), n, CM);
			GPUMatrixDynamic<T> dx(dL_din, in_w, n, lay);
			net->backward(stream, *ctx, inm, outm, dy, dL_din ? &dx : nullptr, false, wgrad ? EGradientMode::Overwrite : EGradientMode::Ignore);
		}
		if ((which & 4) && d_in2) {
			GPUMatrixDynamic<T> d2((T*)d_in2, in_w, n, CM);
			GPUMatrixDynamic<T> sd((T*)seed, net->padded_output_width(), n, CM);
			net->backward_backward_input(stream, *ctx, inm, d2, sd, nullptr, nullptr, false,
				wgrad ? (((which & 2) && dL_dout) ? EGradientMode::Accumulate : EGradientMode::Overwrite) : EGradientMode::Ignore);
		}
	}
	cudaEventRecord(e1, stream);
	CUDA_CHECK_THROW(cudaStreamSynchronize(stream));
	float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
	if (ms_out) *ms_out = ms / iters;
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	ctx.reset();
	return 0;
	REF_CATCH
}

void ref_free_temporary_memory() { free_all_gpu_memory_arenas(); }

// ---- optimizer: the reference's Ema{Adam} exactly as Trainer drives it (trainer.h:170-178; adam.h:52-160,266-324; ema.h:64-147) ----
// A ParametricObject that only describes the parameter layout the optimizer asks for: `n_matrix` leading parameters are covered by
// one "layer" (matrix params: L2 regularisation, always updated), the rest are non-matrix (hash-grid style: skipped when g == 0).
struct RefLayout : public ParametricObject<T> {
	size_t n, n_matrix;
	RefLayout(size_t n_, size_t nm) : n(n_), n_matrix(nm) {}
	void set_params(T*, T*, T*, T*) override {}
	void initialize_params(pcg32&, float*, T*, T*, T*, T*, float) override {}
	size_t n_params() const override { return n; }
	std::vector<std::pair<uint32_t, uint32_t>> layer_sizes() const override {
		if (!n_matrix) return {};
		return {{(uint32_t)n_matrix, 1u}};
	}
	json hyperparams() const override { return {{"otype", "RefLayout"}}; }
};
struct RefOpt { std::shared_ptr<Optimizer<T>> opt; std::shared_ptr<RefLayout> layout; };

void* ref_opt_create(const char* cfg_json, uint64_t n_params, uint64_t n_matrix) {
	try {
		auto* o = new RefOpt;
		o->opt.reset(create_optimizer<T>(json::parse(cfg_json)));
		o->layout = std::make_shared<RefLayout>((size_t)n_params, (size_t)n_matrix);
		o->opt->allocate(o->layout);
		return o;
	} catch (const std::exception& e) { g_err = e.what(); fprintf(stderr, "ref_shim: %s\n", e.what()); return nullptr; }
}
void ref_opt_destroy(void* h) { delete (RefOpt*)h; }
// one Trainer::optimizer_step: master (fp32), params (fp16) and gradients (fp16) are DEVICE pointers; returns the optimizer's
// inference weights (EMA) pointer through ema_out (device, owned by the optimizer)
int ref_opt_step(void* h, float loss_scale, float* master, T* params, const T* grads, T** ema_out) {
	REF_TRY
	auto* o = (RefOpt*)h;
	o->opt->step(nullptr, loss_scale, master, params, grads);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	if (ema_out) *ema_out = o->opt->custom_weights();
	return 0;
	REF_CATCH
}

// ---- integer / RNG primitives of the ngp layer that live in Eigen-free headers ----
uint32_t ref_morton3D(uint32_t x, uint32_t y, uint32_t z) { return morton3D(x, y, z); }                 // common_device.h:347-353
uint32_t ref_morton3D_invert(uint32_t x) { return morton3D_invert(x); }                                  // :355-362
// pcg32 (dependencies/pcg32/pcg32.h): seed, jump ahead, then n draws of each kind
void ref_pcg32_sequence(uint64_t seed, int64_t advance, uint32_t n, uint32_t* uints, float* floats) {
	pcg32 rng{seed};
	rng.advance(advance);
	pcg32 r2 = rng;
	for (uint32_t i = 0; i < n; ++i) uints[i] = rng.next_uint();
	for (uint32_t i = 0; i < n; ++i) floats[i] = r2.next_float();
}
// generate_random_uniform (random.h:67-99): what Trainer::initialize_params fills the hash grid with; out is a DEVICE pointer
int ref_generate_random_uniform(uint64_t seed, uint64_t n, float* out_dev, float lower, float upper, uint64_t* state_after) {
	REF_TRY
	pcg32 rng{seed};
	generate_random_uniform<float>(rng, (size_t)n, out_dev, lower, upper);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	if (state_after) { state_after[0] = rng.state; state_after[1] = rng.inc; }
	return 0;
	REF_CATCH
}

// fill_rollover_and_rescale<half> on the dL/doutput rows and fill_rollover<float> on the coordinates, as train_nerf_step launches them
// (src/testbed_nerf.cu:3922-3930; kernels: tcnn common_device.h:514-534).  All pointers are DEVICE pointers.
int ref_fill_rollover(uint32_t n_elements, const uint32_t* n_input_dev, T* dloss16 /*[n][16]*/, float* coords7 /*[n][7]*/) {
	REF_TRY
	if (dloss16) fill_rollover_and_rescale<T><<<n_blocks_linear(n_elements * 16), n_threads_linear>>>(n_elements, 16, n_input_dev, dloss16);
	if (coords7) fill_rollover<float><<<n_blocks_linear(n_elements * 7), n_threads_linear>>>(n_elements, 7, n_input_dev, coords7);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	return 0;
	REF_CATCH
}

}  // extern "C"
