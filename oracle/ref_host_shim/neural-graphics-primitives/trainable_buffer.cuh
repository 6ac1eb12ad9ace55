// Stand-in (test infrastructure): common_operation.cuh only includes the real header for its transitive Eigen / tcnn includes.
#pragma once
#include <tiny-cuda-nn/common.h>
