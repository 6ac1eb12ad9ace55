// TEST INFRASTRUCTURE ONLY.  Instantiates the REFERENCE's own ngp::NerfNetwork<__half>
// (/root/reference/include/neural-graphics-primitives/nerf_network.h: forward_impl / backward_impl with the in-forward normal, the
// second-order terms and the DeltaNetwork branch) on top of the reference's tiny-cuda-nn (oracle/_ref/obj/*.o, built by
// oracle/Makefile.ref) for sm_100.  Only Eigen is shadowed: oracle/ref_eigen_shim/Eigen/Dense forwards to the stand-in of
// oracle/ref_host_shim/ (the snapshot's Eigen/Core is missing); every other header is the reference's own, read where it lies.
// Built into oracle/_ref/libngp_network_ref.so by `make -f oracle/Makefile.ref network`.  C ABI; device pointers in and out; all work on the
// legacy default stream.  Never linked or loaded by the product.
#include <tiny-cuda-nn/common.h>
using namespace tcnn;   // nerf_network.h:1098 calls parallel_for_gpu unqualified (the reference gets the directive through testbed.h)
#include <neural-graphics-primitives/nerf_network.h>

#include <memory>
#include <string>
#include <vector>

using precision_t = network_precision_t;
using json = nlohmann::json;

namespace {
struct RefNet {
	std::shared_ptr<ngp::NerfNetwork<precision_t>> net;
	GPUMemory<precision_t> params, grads;          // the Trainer's m_params / m_param_gradients (trainer.h), reference order
	GPUMemory<precision_t> delta_params, delta_grads;  // m_global_move.trainer's: [transition 4 | rotation-6d 8] (transform_network.h:313-333)
	size_t n = 0;
};
std::string g_err;
}

#define REFNET_TRY try {
#define REFNET_CATCH } catch (const std::exception& e) { g_err = e.what(); return -1; }

extern "C" {

const char* refnet_last_error() { return g_err.c_str(); }

// cfg_json: {"encoding": ..., "dir_encoding": ..., "network": ..., "rgb_network": ...} -- the four blocks of configs/nerf/base.json that
// Testbed::reset_network hands to the constructor (src/testbed.cu:2219-2228: n_pos_dims 3, n_dir_dims 3, n_extra_dims 0, dir_offset 4)
void* refnet_create(const char* cfg_json) {
	try {
		json cfg = json::parse(cfg_json);
		auto* r = new RefNet();
		r->net = std::make_shared<ngp::NerfNetwork<precision_t>>(3u, 3u, 0u, 4u, cfg["encoding"], cfg["dir_encoding"], cfg["network"], cfg["rgb_network"]);
		r->net->set_anneal_end(cfg.value("anneal_end", 0));
		r->n = r->net->n_params();
		r->params.resize(r->n); r->grads.resize(r->n);
		r->params.memset(0); r->grads.memset(0);
		r->net->set_params(r->params.data(), r->params.data(), r->params.data(), r->grads.data());
		r->delta_params.resize(12); r->delta_grads.resize(12);
		r->delta_params.memset(0); r->delta_grads.memset(0);
		r->net->delta_network()->set_params(r->delta_params.data(), r->delta_params.data(), r->delta_params.data(), r->delta_grads.data());
		r->net->m_use_delta = false; r->net->m_train_delta = false;     // frame 0 (src/testbed.cu:2704-2707)
		CUDA_CHECK_THROW(cudaDeviceSynchronize());
		return r;
	} catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void refnet_destroy(void* h) { delete (RefNet*)h; }
uint64_t refnet_n_params(void* h) { return ((RefNet*)h)->n; }

// half parameters in the network's own order: density MLP | rgb MLP | hash grid | (dir encoding: none) | variance 4  (set_params, nerf_network.h:741-790)
int refnet_set_params(void* h, const void* params_half_dev, uint64_t n) {
	REFNET_TRY
	RefNet* r = (RefNet*)h;
	if (n != r->n) throw std::runtime_error{"refnet_set_params: parameter count mismatch"};
	CUDA_CHECK_THROW(cudaMemcpy(r->params.data(), params_half_dev, n * sizeof(precision_t), cudaMemcpyDeviceToDevice));
	return 0;
	REFNET_CATCH
}
// m_training_step of the network and of the hash grid (progressive levels), as Testbed::train sets them before every step (src/testbed.cu:2651-2656)
int refnet_set_training_step(void* h, int32_t step) {
	REFNET_TRY
	RefNet* r = (RefNet*)h;
	r->net->m_training_step = step > 0 ? (uint32_t)step : 0u;
	r->net->encoding()->set_training_step(step);
	return 0;
	REFNET_CATCH
}
// m_use_delta / m_train_delta and the DeltaNetwork's 12 half parameters on the HOST ([transition x y z 0 | rotation-6d a1 a2 0 0]); null keeps them
int refnet_set_delta(void* h, int use_delta, int train_delta, const void* delta_params_half_host) {
	REFNET_TRY
	RefNet* r = (RefNet*)h;
	r->net->m_use_delta = use_delta != 0; r->net->m_train_delta = train_delta != 0;
	if (delta_params_half_host) CUDA_CHECK_THROW(cudaMemcpy(r->delta_params.data(), delta_params_half_host, 12 * sizeof(precision_t), cudaMemcpyHostToDevice));
	return 0;
	REFNET_CATCH
}

// Network::forward as Testbed::train_nerf_step calls it (src/testbed_nerf.cu:3854 / :3941): coords_dev [n][7] float (NerfCoordinate),
// out_dev [n][16] half; n must be a multiple of 128 (FullyFusedMLP batch granularity)
int refnet_forward(void* h, const float* coords_dev, uint32_t n, void* out_half_dev) {
	REFNET_TRY
	RefNet* r = (RefNet*)h;
	GPUMatrix<float> in_m((float*)coords_dev, 7, n);
	GPUMatrix<precision_t> out_m((precision_t*)out_half_dev, 16, n);
	{
		auto ctx = r->net->forward(nullptr, in_m, &out_m, false, false);
		CUDA_CHECK_THROW(cudaDeviceSynchronize());
	}
	return 0;
	REFNET_CATCH
}

// forward + backward with EGradientMode::Overwrite (src/testbed_nerf.cu:3941-3946).  grads_half_dev: [n_params] half (the Trainer's gradient
// buffer, loss-scaled like dL_dout); delta_grads_half_host: 12 halves [transition 4 | rotation 8] when m_use_delta && m_train_delta, may be null
int refnet_backward(void* h, const float* coords_dev, const void* dL_dout_half_dev, uint32_t n, uint32_t indeed_batch_size, void* out_half_dev,
                    void* grads_half_dev, void* delta_grads_half_host) {
	REFNET_TRY
	RefNet* r = (RefNet*)h;
	r->grads.memset(0);
	r->delta_grads.memset(0);
	GPUMatrix<float> in_m((float*)coords_dev, 7, n);
	GPUMatrix<precision_t> out_m((precision_t*)out_half_dev, 16, n), dl_m((precision_t*)dL_dout_half_dev, 16, n);
	r->net->indeed_batch_size = indeed_batch_size;
	{
		auto ctx = r->net->forward(nullptr, in_m, &out_m, false, false);
		r->net->backward(nullptr, *ctx, in_m, out_m, dl_m, nullptr, false, EGradientMode::Overwrite);
		CUDA_CHECK_THROW(cudaDeviceSynchronize());
	}
	CUDA_CHECK_THROW(cudaMemcpy(grads_half_dev, r->grads.data(), r->n * sizeof(precision_t), cudaMemcpyDeviceToDevice));
	if (delta_grads_half_host) CUDA_CHECK_THROW(cudaMemcpy(delta_grads_half_host, r->delta_grads.data(), 12 * sizeof(precision_t), cudaMemcpyDeviceToHost));
	return 0;
	REFNET_CATCH
}

// Timing of the reference's network part of one training step, with CUDA events on the legacy default stream (after `warmup` untimed rounds):
// ms_out[0] = inference_mixed_precision over n_fwd samples (the pass over all generated samples, src/testbed_nerf.cu:3854),
// ms_out[1] = forward + backward over n_bwd compacted samples (:3941-3946).  coords_dev / dL_dout_half_dev hold max(n_fwd, n_bwd) rows.
int refnet_time_step(void* h, const float* coords_dev, const void* dL_dout_half_dev, uint32_t n_fwd, uint32_t n_bwd, uint32_t indeed_batch_size,
                     int warmup, int iters, float* ms_out) {
	REFNET_TRY
	RefNet* r = (RefNet*)h;
	const uint32_t n_max = n_fwd > n_bwd ? n_fwd : n_bwd;
	GPUMemory<precision_t> out(n_max * 16);
	GPUMatrix<float> in_f((float*)coords_dev, 7, n_fwd), in_b((float*)coords_dev, 7, n_bwd);
	GPUMatrix<precision_t> out_f(out.data(), 16, n_fwd), out_b(out.data(), 16, n_bwd), dl_b((precision_t*)dL_dout_half_dev, 16, n_bwd);
	r->net->indeed_batch_size = indeed_batch_size;
	cudaEvent_t e0, e1, e2;
	CUDA_CHECK_THROW(cudaEventCreate(&e0)); CUDA_CHECK_THROW(cudaEventCreate(&e1)); CUDA_CHECK_THROW(cudaEventCreate(&e2));
	float ms_f = 0, ms_b = 0;
	for (int it = -warmup; it < iters; ++it) {
		CUDA_CHECK_THROW(cudaEventRecord(e0, nullptr));
		r->net->inference_mixed_precision(nullptr, in_f, out_f, false);
		CUDA_CHECK_THROW(cudaEventRecord(e1, nullptr));
		{
			auto ctx = r->net->forward(nullptr, in_b, &out_b, false, false);
			r->net->backward(nullptr, *ctx, in_b, out_b, dl_b, nullptr, false, EGradientMode::Overwrite);
			CUDA_CHECK_THROW(cudaEventRecord(e2, nullptr));
			CUDA_CHECK_THROW(cudaEventSynchronize(e2));
		}
		float a = 0, b = 0;
		CUDA_CHECK_THROW(cudaEventElapsedTime(&a, e0, e1)); CUDA_CHECK_THROW(cudaEventElapsedTime(&b, e1, e2));
		if (it >= 0) { ms_f += a; ms_b += b; }
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
	ms_out[0] = ms_f / iters; ms_out[1] = ms_b / iters;
	return 0;
	REFNET_CATCH
}

// NerfNetwork::sdf (nerf_network.h:645-713; marching-cubes lattice, src/testbed_nerf.cu:4127-4131) or ::density (:715-733; occupancy-grid
// update, :3364-3366, training parameters): pos_dev [n][3] float (NerfPosition), out_half_dev [n] half -- row 0 of the reference's
// row-major [padded_output_width][n] output.  n must be a multiple of 128.
int refnet_sdf(void* h, const float* pos_dev, uint32_t n, void* out_half_dev, int as_density) {
	REFNET_TRY
	RefNet* r = (RefNet*)h;
	const uint32_t w = r->net->padded_density_output_width();
	GPUMemory<precision_t> out((size_t)w * n);
	GPUMatrix<float> in_m((float*)pos_dev, 3, n);
	GPUMatrix<precision_t, RM> out_m(out.data(), w, n);
	if (as_density) r->net->density(nullptr, in_m, out_m, false); else r->net->sdf(nullptr, in_m, out_m, false);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	CUDA_CHECK_THROW(cudaMemcpy(out_half_dev, out.data(), (size_t)n * sizeof(precision_t), cudaMemcpyDeviceToDevice));
	return 0;
	REFNET_CATCH
}

}  // extern "C"
