// TEST INFRASTRUCTURE ONLY -- the interface classes of tcnn object.h that optimizer.h / adam.h need (Object, ObjectWithMutableHyperparams,
// ParametricObject: pure declarations), without gpu_matrix.h and the DifferentiableObject machinery behind them.
#pragma once
#include <tiny-cuda-nn/common.h>
#include <json/json.hpp>
namespace tcnn {
using json = nlohmann::json;
class Object {
public:
	virtual ~Object() {}
	virtual json hyperparams() const = 0;
};
class ObjectWithMutableHyperparams : public Object {
public:
	virtual void update_hyperparams(const json& params) = 0;
};
template <typename PARAMS_T>
class ParametricObject : public Object {
public:
	virtual size_t n_params() const = 0;
	virtual size_t n_params_canonical() const { return n_params(); }
	virtual size_t n_params_delta() const { return 0; }
	virtual json n_params_components() const { return json({}); }
	virtual std::vector<std::pair<uint32_t, uint32_t>> layer_sizes() const = 0;
	virtual std::vector<std::pair<uint32_t, uint32_t>> layer_sizes_canonical() const { return layer_sizes(); }
	virtual std::vector<std::pair<uint32_t, uint32_t>> layer_sizes_delta() const { return {}; }
};
}
