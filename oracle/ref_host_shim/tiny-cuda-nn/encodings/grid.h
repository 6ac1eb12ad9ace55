// Stand-in (test infrastructure): see tiny-cuda-nn/common.h in this directory.
#pragma once
#include <tiny-cuda-nn/common.h>
