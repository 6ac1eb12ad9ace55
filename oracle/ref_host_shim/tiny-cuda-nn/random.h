// Stand-in (test infrastructure) for tiny-cuda-nn/random.h: the generator itself is the REFERENCE's own pcg32.h
// (dependencies/my_tcnn/dependencies/pcg32, on the include path of oracle/Makefile.ref), only the CUDA kernels around it are left out.
#pragma once
#include <tiny-cuda-nn/common.h>
#include <pcg32/pcg32.h>
namespace tcnn { using default_rng_t = pcg32; }
