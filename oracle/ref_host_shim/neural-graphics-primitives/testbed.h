// Stand-in (test infrastructure) for neural-graphics-primitives/testbed.h: the kernel part of src/testbed_nerf.cu (lines 1-2396) only
// needs the declaration of Testbed::network_dims_nerf (defined at :141) and the headers testbed.h pulls in for the kernels.
#pragma once
#include <cassert>
#include <sstream>
#include <string>
#include <vector>
#include <neural-graphics-primitives/common.h>
#include <neural-graphics-primitives/common_device.cuh>
#include <neural-graphics-primitives/bounding_box.cuh>
#include <neural-graphics-primitives/nerf.h>
#include <neural-graphics-primitives/nerf_loader.h>
NGP_NAMESPACE_BEGIN
class Testbed {
public:
	struct NetworkDims { uint32_t n_input, n_output, n_pos; };
	NetworkDims network_dims_nerf() const;
};
NGP_NAMESPACE_END
