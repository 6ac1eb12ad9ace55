// TEST INFRASTRUCTURE ONLY -- stand-ins that let g++ compile the reference's OPTIMIZER headers (dependencies/my_tcnn/include/tiny-cuda-nn/
// optimizers/{adam,ema,exponential_decay}.h, optimizer.h, gpu_memory_json.h) AS HOST CODE from where they lie: their kernels run one simulated
// thread at a time and "device" memory is host memory, so that serialize() / deserialize() -- the "optimizer" subtree of a snapshot file -- are
// the reference's own code (oracle/ref_snapshot_host.cpp).  Builds on the keyword / __half stand-ins of oracle/ref_host_shim.
#pragma once
#define REF_HOST_FUNCTIONAL_GPU_MEMORY 1
#include "../../ref_host_shim/tiny-cuda-nn/common.h"
#include <cctype>
#include <functional>
#include <iostream>
#include <limits>
#include <memory>
#include <stdexcept>
#include <utility>

typedef void* cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
inline int cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind) { if (n) std::memcpy(dst, src, n); return 0; }
inline int cudaDeviceSynchronize() { return 0; }
inline int cudaFree(void* p) { std::free(p); return 0; }
#define CUDA_CHECK_THROW(x) do { if ((x) != 0) throw std::runtime_error(#x " failed"); } while (0)

namespace tcnn {
// the two host helpers of tcnn common.h the dataset loader uses (common.h:118-124 equals_case_insensitive, :180-197 ScopeGuard): stand-ins
inline bool equals_case_insensitive(const std::string& a, const std::string& b) {
	if (a.size() != b.size()) return false;
	for (size_t i = 0; i < a.size(); ++i) if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
	return true;
}
class ScopeGuard {
	std::function<void()> m_callback;
public:
	ScopeGuard(const std::function<void()>& callback) : m_callback{callback} {}
	ScopeGuard(const ScopeGuard&) = delete;
	ScopeGuard& operator=(const ScopeGuard&) = delete;
	~ScopeGuard() { if (m_callback) m_callback(); }
};
// linear_kernel (tcnn common.h:296-307): kernel<<<blocks, 128>>>((uint32_t)n_elements, args...) -- here one simulated thread after the other
template <typename K, typename T, typename... Types>
inline void linear_kernel(K kernel, uint32_t, cudaStream_t, T n_elements, Types... args) {
	blockDim.x = 1; blockIdx.x = 0;
	for (uint32_t i = 0; i < (uint32_t)n_elements; ++i) { threadIdx.x = i; kernel((uint32_t)n_elements, args...); }
	threadIdx.x = 0;
}
}
