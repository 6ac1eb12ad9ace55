// TEST INFRASTRUCTURE ONLY -- tcnn::GPUMemory<T> over host memory (see common.h in this directory): the members the optimizer headers use.
#pragma once
#include <tiny-cuda-nn/common.h>
namespace tcnn {
template <typename T>
class GPUMemory {
	T* m_data = nullptr; size_t m_size = 0;
public:
	GPUMemory() {}
	explicit GPUMemory(size_t n) { resize(n); }
	GPUMemory(const GPUMemory& o) { *this = o; }
	GPUMemory(GPUMemory&& o) noexcept : m_data(o.m_data), m_size(o.m_size) { o.m_data = nullptr; o.m_size = 0; }
	~GPUMemory() { std::free(m_data); }
	GPUMemory& operator=(const GPUMemory& o) { if (this != &o) { resize(o.m_size); if (m_size) std::memcpy(m_data, o.m_data, get_bytes()); } return *this; }
	GPUMemory& operator=(GPUMemory&& o) noexcept { if (this != &o) { std::free(m_data); m_data = o.m_data; m_size = o.m_size; o.m_data = nullptr; o.m_size = 0; } return *this; }
	void resize(size_t n) { if (n != m_size) { std::free(m_data); m_data = n ? (T*)std::malloc(n * sizeof(T)) : nullptr; m_size = n; } }
	void memset(int v) { if (m_size) std::memset(m_data, v, get_bytes()); }
	void enlarge(size_t n) { if (n > m_size) { T* d = (T*)std::malloc(n * sizeof(T)); if (m_size) std::memcpy((void*)d, (const void*)m_data, get_bytes()); std::free(m_data); m_data = d; m_size = n; } }
	void free_memory() { std::free(m_data); m_data = nullptr; m_size = 0; }
	void copy_from_host(const T* host) { if (m_size) std::memcpy((void*)m_data, (const void*)host, get_bytes()); }
	void copy_to_host(std::vector<T>& host) const { if (host.size() < m_size) throw std::runtime_error{"copy_to_host: buffer too small"}; copy_to_host(host.data()); }
	void copy_to_host(T* host) const { if (m_size) std::memcpy((void*)host, (const void*)m_data, get_bytes()); }
	T* data() const { return m_data; }
	size_t size() const { return m_size; }
	size_t get_bytes() const { return m_size * sizeof(T); }
};
}
