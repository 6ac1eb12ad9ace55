// Stand-in (test infrastructure) for neural-graphics-primitives/nerf_network.h: the kernels of src/testbed_nerf.cu reach the rotation-6d
// helpers through it; those come from the reference's real common_operation.cuh.
#pragma once
#include <neural-graphics-primitives/common_operation.cuh>
