// TEST INFRASTRUCTURE ONLY -- forwards to the stand-in of oracle/ref_host_shim.
#pragma once
#include "../../ref_host_shim/tiny-cuda-nn/random.h"
