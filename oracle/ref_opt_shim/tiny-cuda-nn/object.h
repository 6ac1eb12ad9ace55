// TEST INFRASTRUCTURE ONLY -- stand-in for the three interface types of tiny-cuda-nn's object.h that optimizer.h / adam.h name (what an optimizer
// asks of its target: parameter counts, the named parameter blocks, the matrix shapes).  No gpu_matrix.h, no DifferentiableObject.
#pragma once
#include <tiny-cuda-nn/common.h>
#include <json/json.hpp>
namespace tcnn {
using json = nlohmann::json;
typedef std::vector<std::pair<uint32_t, uint32_t>> RefLayerSizes;
struct Object { virtual ~Object() = default; virtual json hyperparams() const = 0; };
struct ObjectWithMutableHyperparams : Object { virtual void update_hyperparams(const json& new_values) = 0; };
template <typename P> struct ParametricObject : Object {
	virtual size_t n_params() const = 0;                                              // all of them ...
	virtual size_t n_params_canonical() const { return this->n_params(); }            // ... of which canonical / delta (the latter none by default)
	virtual size_t n_params_delta() const { return 0u; }
	virtual json n_params_components() const { return json::object(); }               // {index: [name, count]} blocks, in parameter order
	virtual RefLayerSizes layer_sizes() const = 0;                                    // (rows, cols) of the weight matrices at the front
	virtual RefLayerSizes layer_sizes_canonical() const { return this->layer_sizes(); }
	virtual RefLayerSizes layer_sizes_delta() const { return RefLayerSizes{}; }
};
}
