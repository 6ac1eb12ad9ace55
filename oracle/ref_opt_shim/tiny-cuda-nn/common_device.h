// TEST INFRASTRUCTURE ONLY -- empty on purpose: oracle/Makefile.ref streams the two helpers the optimizer headers use (weight_decay, cast /
// cast_from) from the reference's own common_device.h into the translation unit ahead of them.
#pragma once
#include <tiny-cuda-nn/common.h>
