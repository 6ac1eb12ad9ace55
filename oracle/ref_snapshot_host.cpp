// TEST INFRASTRUCTURE ONLY.  The reference's snapshot (de)serialisation run on the CPU: its own optimizer classes -- AdamOptimizer,
// ExponentialDecayOptimizer, EmaOptimizer (dependencies/my_tcnn/include/tiny-cuda-nn/optimizers/*.h, incl. their step kernels) -- its own
// GPUMemory <-> json binary conversion (gpu_memory_json.h) and its own nlohmann json / msgpack library, all compiled AS HOST CODE from where
// they lie against the stand-ins of oracle/ref_opt_shim/ (oracle/Makefile.ref -> oracle/_ref/libref_snapshot_host.so).  What the application
// wraps around them is short and restated here with the lines it follows: Trainer::serialize / deserialize (trainer.h:281-304),
// Testbed::save_snapshot / load_snapshot (src/testbed.cu:3144-3179, 3198-3252), NerfNetwork::save_/load_global_/local_movement
// (nerf_network.h:1179-1265) and the three-name dispatch of create_optimizer (src/optimizer.cu).  Used to check that a file written by
// neus2_b200.snapshot loads through the reference's code and vice versa (tests/test_golden_snapshot.py).
#include <tiny-cuda-nn/optimizers/adam.h>
#include <tiny-cuda-nn/optimizers/ema.h>
#include <tiny-cuda-nn/optimizers/exponential_decay.h>

#include <fstream>

RefHostIdx threadIdx, blockIdx;
RefHostDim blockDim, gridDim;

namespace tcnn {
// create_optimizer (dependencies/my_tcnn/src/optimizer.cu): the three types base.json stacks
template <typename T>
Optimizer<T>* create_optimizer(const json& params) {
	const std::string t = params.value("otype", "Adam");
	if (t == "Adam" || t == "adam") return new AdamOptimizer<T>{params};
	if (t == "Ema" || t == "EMA" || t == "ema") return new EmaOptimizer<T>{params};
	if (t == "ExponentialDecay" || t == "exponentialdecay") return new ExponentialDecayOptimizer<T>{params};
	throw std::runtime_error{"Invalid optimizer type: " + t};
}
}

using namespace tcnn;
using T = __half;

namespace {
// the parameter blocks of NerfNetwork as the optimizer sees them (nerf_network.h:912-932, layer_sizes :934-951)
struct Target : public ParametricObject<T> {
	uint32_t n_density, n_rgb, n_grid, n_variance;
	std::vector<std::pair<uint32_t, uint32_t>> layers;
	json hyperparams() const override { return {{"otype", "RefSnapshotTarget"}}; }
	size_t n_params() const override { return (size_t)n_density + n_rgb + n_grid + n_variance; }
	json n_params_components() const override {
		return {{0, {"density_network", n_density}}, {1, {"rgb_network", n_rgb}}, {2, {"variance_network", n_variance}}, {3, {"pos_encoding", n_grid}}, {4, {"dir_encoding", 0}}};
	}
	std::vector<std::pair<uint32_t, uint32_t>> layer_sizes() const override { return layers; }
};
struct Stack {
	std::shared_ptr<Target> target;
	std::unique_ptr<Optimizer<T>> opt;
	GPUMemory<float> master;       // Trainer::m_params_full_precision
	GPUMemory<T> params, grads;    // m_params, m_param_gradients
	T* params_inference = nullptr; // m_params_inference (trainer.h:90-93)
	float loss = 0; uint32_t training_step = 0, rays_per_batch = 0, measured = 0, measured_before = 0; int aabb_scale = 1;
	GPUMemory<T> density_grid, rotation, transition, local_rotation, local_transition;
};
std::string g_err;
void bits_to_half(const uint16_t* bits, T* out, size_t n) { std::memcpy((void*)out, bits, n * 2); }
}

#define REFS_TRY try {
#define REFS_CATCH } catch (const std::exception& e) { g_err = e.what(); return -1; }

extern "C" {
const char* refs_last_error() { return g_err.c_str(); }

// layers: n_layers (rows, cols) pairs of the matrices at the start of the parameter vector (density MLP, then rgb MLP)
void* refs_create(const char* optimizer_json, uint32_t n_density, uint32_t n_rgb, uint32_t n_grid, uint32_t n_variance, const uint32_t* layers, uint32_t n_layers) {
	try {
		auto* s = new Stack();
		s->target = std::make_shared<Target>();
		s->target->n_density = n_density; s->target->n_rgb = n_rgb; s->target->n_grid = n_grid; s->target->n_variance = n_variance;
		for (uint32_t i = 0; i < n_layers; ++i) s->target->layers.push_back({layers[2 * i], layers[2 * i + 1]});
		s->opt.reset(create_optimizer<T>(json::parse(optimizer_json)));
		s->opt->allocate(s->target);                                   // Trainer::Trainer -> m_optimizer->allocate(m_model) (trainer.h:60-66)
		const size_t n = s->target->n_params();
		s->master.resize(n); s->params.resize(n); s->grads.resize(n);
		s->master.memset(0); s->params.memset(0); s->grads.memset(0);
		s->params_inference = s->opt->custom_weights() ? s->opt->custom_weights() : s->params.data();
		return s;
	} catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void refs_destroy(void* h) { delete (Stack*)h; }

// Trainer::set_params_full_precision (trainer.h:248-260)
int refs_set_params_full_precision(void* h, const float* master, uint64_t n) {
	REFS_TRY
	Stack* s = (Stack*)h;
	if (n != s->target->n_params()) throw std::runtime_error{"Can't set params because CPU buffer has the wrong size."};
	std::memcpy(s->master.data(), master, n * sizeof(float));
	for (size_t i = 0; i < n; ++i) s->params_inference[i] = (T)s->master.data()[i];
	if (s->params_inference != s->params.data()) std::memcpy((void*)s->params.data(), (const void*)s->params_inference, n * sizeof(T));
	return 0;
	REFS_CATCH
}
// Trainer::optimizer_step (trainer.h:170-173): gradients as half bits
int refs_optimizer_step(void* h, float loss_scale, const uint16_t* grad_bits) {
	REFS_TRY
	Stack* s = (Stack*)h;
	bits_to_half(grad_bits, s->grads.data(), s->target->n_params());
	s->opt->step(nullptr, loss_scale, s->master.data(), s->params.data(), s->grads.data());
	return 0;
	REFS_CATCH
}
// state read-back through the optimizers' OWN serialize(): every pointer may be null
int refs_get_state(void* h, float* master, uint16_t* params, uint16_t* params_inference, float* m1, float* m2, uint32_t* steps, uint16_t* ema, uint32_t* current_step,
                   float* learning_rate, float* learning_rate_factor) {
	REFS_TRY
	Stack* s = (Stack*)h;
	const size_t n = s->target->n_params();
	if (master) std::memcpy(master, s->master.data(), n * 4);
	if (params) std::memcpy(params, (const void*)s->params.data(), n * 2);
	if (params_inference) std::memcpy(params_inference, (const void*)s->params_inference, n * 2);
	const json o = s->opt->serialize();                              // Ema{ nested: ExponentialDecay{ nested: Adam } }
	auto blob = [&](const json& j, void* dst, size_t bytes) { const json::binary_t& b = j.get_binary(); if (b.size() != bytes) throw std::runtime_error{"blob size"}; std::memcpy(dst, b.data(), bytes); };
	const json& adam = o["nested"]["nested"];
	if (m1) blob(adam["first_moments_binary"], m1, n * 4);
	if (m2) blob(adam["second_moments_binary"], m2, n * 4);
	if (steps) blob(adam["param_steps_binary"], steps, n * 4);
	if (ema) blob(o["weights_ema_binary"], ema, n * 2);
	if (current_step) *current_step = adam["current_step"];
	if (learning_rate) *learning_rate = o["nested"]["learning_rate"];
	if (learning_rate_factor) *learning_rate_factor = o["nested"]["learning_rate_factor"];
	return 0;
	REFS_CATCH
}
// the application-side numbers a snapshot carries (Testbed members); blobs as half bits, null = leave empty
int refs_set_testbed_state(void* h, uint32_t training_step, float loss, uint32_t rays_per_batch, uint32_t measured, uint32_t measured_before, int aabb_scale,
                           const uint16_t* density_grid_bits, uint64_t n_cells, const uint16_t* rotation12, const uint16_t* transition4, const uint16_t* local_rotation8,
                           const uint16_t* local_transition4) {
	REFS_TRY
	Stack* s = (Stack*)h;
	s->training_step = training_step; s->loss = loss; s->rays_per_batch = rays_per_batch; s->measured = measured; s->measured_before = measured_before; s->aabb_scale = aabb_scale;
	s->density_grid.resize(n_cells); if (n_cells) bits_to_half(density_grid_bits, s->density_grid.data(), n_cells);
	auto set = [](GPUMemory<T>& m, const uint16_t* bits, size_t n) { m.resize(bits ? n : 0); if (bits) bits_to_half(bits, m.data(), n); };
	set(s->rotation, rotation12, 12); set(s->transition, transition4, 4); set(s->local_rotation, local_rotation8, 8); set(s->local_transition, local_transition4, 4);
	return 0;
	REFS_CATCH
}

// Testbed::save_snapshot (src/testbed.cu:3144-3179) without the dataset block: network_config["snapshot"] = Trainer::serialize(...), the global /
// local movement blobs, the Nerf-mode fields, then json::to_msgpack
int refs_save_snapshot(void* h, const char* network_config_json, const char* path, int include_optimizer_state) {
	REFS_TRY
	Stack* s = (Stack*)h;
	json cfg = json::parse(network_config_json);
	{   // Trainer::serialize (trainer.h:281-293)
		const size_t n_params = s->target->n_params();
		json data;
		data["n_params"] = n_params;
		data["params_binary"] = gpu_memory_to_json_binary(s->params_inference, sizeof(T) * n_params);
		if (include_optimizer_state) data["optimizer"] = s->opt->serialize();
		cfg["snapshot"] = data;
	}
	// save_global_movement / save_local_movement (nerf_network.h:1203-1204, 1244-1245)
	cfg["snapshot"]["rotation"] = gpu_memory_to_json_binary(s->rotation.data(), s->rotation.get_bytes());
	cfg["snapshot"]["transition"] = gpu_memory_to_json_binary(s->transition.data(), s->transition.get_bytes());
	cfg["snapshot"]["local_rotation"] = gpu_memory_to_json_binary(s->local_rotation.data(), s->local_rotation.get_bytes());
	cfg["snapshot"]["local_transition"] = gpu_memory_to_json_binary(s->local_transition.data(), s->local_transition.get_bytes());
	auto& snapshot = cfg["snapshot"];
	snapshot["density_grid_size"] = 128;                              // NERF_GRIDSIZE()
	snapshot["density_grid_binary"] = s->density_grid;                // to_json(GPUMemory<__half>) (gpu_memory_json.h:49-52)
	snapshot["nerf"]["aabb_scale"] = s->aabb_scale;
	cfg["snapshot"]["training_step"] = s->training_step;
	cfg["snapshot"]["loss"] = s->loss;
	cfg["snapshot"]["nerf"]["rgb"]["rays_per_batch"] = s->rays_per_batch;
	cfg["snapshot"]["nerf"]["rgb"]["measured_batch_size"] = s->measured;
	cfg["snapshot"]["nerf"]["rgb"]["measured_batch_size_before_compaction"] = s->measured_before;
	std::ofstream f(path, std::ios::out | std::ios::binary);
	json::to_msgpack(cfg, f);
	return 0;
	REFS_CATCH
}

// Testbed::load_snapshot (src/testbed.cu:3198-3252): load_network_config's msgpack branch (:155-158), the Nerf-mode reads, Trainer::deserialize
// (trainer.h:295-304 -> set_params :262-275, m_optimizer->deserialize), load_global_/local_movement (nerf_network.h:1207-1265: half blobs)
int refs_load_snapshot(void* h, const char* path) {
	REFS_TRY
	Stack* s = (Stack*)h;
	std::ifstream f{path, std::ios::in | std::ios::binary};
	json config = json::from_msgpack(f);
	if (!config.contains("snapshot")) throw std::runtime_error{std::string{"File '"} + path + "' does not contain a snapshot."};
	const auto& snapshot = config["snapshot"];
	if (snapshot["density_grid_size"] != 128) throw std::runtime_error{"Incompatible grid size."};
	s->rays_per_batch = snapshot["nerf"]["rgb"]["rays_per_batch"];
	s->measured = snapshot["nerf"]["rgb"]["measured_batch_size"];
	s->measured_before = snapshot["nerf"]["rgb"]["measured_batch_size_before_compaction"];
	if (snapshot["nerf"].contains("aabb_scale")) s->aabb_scale = snapshot["nerf"]["aabb_scale"];
	GPUMemory<T> density_grid_fp16 = snapshot["density_grid_binary"];       // from_json(GPUMemory<__half>) (gpu_memory_json.h:54-59)
	if (density_grid_fp16.size() != 128u * 128u * 128u && density_grid_fp16.size() != 0) throw std::runtime_error{"Incompatible number of grid cascades."};
	s->density_grid = density_grid_fp16;
	s->training_step = config["snapshot"]["training_step"];
	s->loss = config["snapshot"]["loss"];
	{   // Trainer::deserialize
		const json& data = config["snapshot"];
		json::binary_t params_binary = data["params_binary"];
		const size_t n = params_binary.size() / sizeof(T);
		if (n != s->target->n_params()) throw std::runtime_error{"Can't set params because CPU buffer has the wrong size."};
		std::memcpy((void*)s->params_inference, params_binary.data(), n * sizeof(T));
		if (s->params_inference != s->params.data()) std::memcpy((void*)s->params.data(), (const void*)s->params_inference, n * sizeof(T));
		for (size_t i = 0; i < n; ++i) s->master.data()[i] = (float)s->params_inference[i];
		if (data.contains("optimizer")) s->opt->deserialize(data["optimizer"]);
		s->params_inference = s->opt->custom_weights() ? s->opt->custom_weights() : s->params.data();
	}
	{ GPUMemory<T> p = config["snapshot"]["rotation"]; s->rotation = p; }
	{ GPUMemory<T> p = config["snapshot"]["transition"]; s->transition = p; }
	{ GPUMemory<T> p = config["snapshot"]["local_rotation"]; s->local_rotation = p; }
	{ GPUMemory<T> p = config["snapshot"]["local_transition"]; s->local_transition = p; }
	return 0;
	REFS_CATCH
}
// what refs_load_snapshot / refs_set_testbed_state left in the application-side members; blob sizes in elements
int refs_get_testbed_state(void* h, uint32_t* u5, float* loss, int* aabb_scale, uint64_t* sizes5, uint16_t* density_grid_bits, uint16_t* rotation, uint16_t* transition,
                           uint16_t* local_rotation, uint16_t* local_transition) {
	Stack* s = (Stack*)h;
	u5[0] = s->training_step; u5[1] = s->rays_per_batch; u5[2] = s->measured; u5[3] = s->measured_before; u5[4] = 0;
	*loss = s->loss; *aabb_scale = s->aabb_scale;
	GPUMemory<T>* m[5] = {&s->density_grid, &s->rotation, &s->transition, &s->local_rotation, &s->local_transition};
	uint16_t* out[5] = {density_grid_bits, rotation, transition, local_rotation, local_transition};
	for (int i = 0; i < 5; ++i) { sizes5[i] = m[i]->size(); if (out[i] && m[i]->size()) std::memcpy(out[i], (const void*)m[i]->data(), m[i]->get_bytes()); }
	return 0;
}
}  // extern "C"
