// TEST INFRASTRUCTURE ONLY -- host stand-ins that let g++ compile the reference's
// include/neural-graphics-primitives/common_operation.cuh (the DeltaNetwork / rotation-6d kernels) AS HOST CODE, straight from
// /root/reference, so that oracle/neus_oracle.cpp's restatement of them can be checked against the reference's own source
// (oracle/ref_delta_host.cpp, tests/test_ref_delta_host.py).  The reference's ngp layer cannot be built as CUDA here: Eigen/Core
// is missing from the snapshot.  What is provided: the CUDA keywords as no-ops, thread indices as globals (one
// "thread" is run at a time), a half type with CUDA's conversion / operator semantics (round-to-nearest-even on every store,
// arithmetic through float), tcnn::MatrixView, and a 3-vector / 3x3-matrix subset of Eigen with Eigen's semantics
// (column-major default, row-by-row comma initialiser, cofactor inverse).  Nothing here is copied from tiny-cuda-nn or Eigen.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <sstream>
#include <string>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __restrict__

struct RefHostIdx { uint32_t x = 0, y = 0, z = 0; };
struct RefHostDim { uint32_t x = 1, y = 1, z = 1; };
extern RefHostIdx threadIdx, blockIdx;
extern RefHostDim blockDim, gridDim;

// ---- __half with the semantics of cuda_fp16.h (implicit float <-> half, operators computed in float and rounded once) ----
struct __half {
	_Float16 h;
	__half() : h((_Float16)0.0f) {}
	__half(float f) : h((_Float16)f) {}
	__half(double f) : h((_Float16)(float)f) {}
	__half(int i) : h((_Float16)(float)i) {}
	operator float() const { return (float)h; }
};
inline __half operator+(const __half& a, const __half& b) { return __half((float)a + (float)b); }
inline __half operator-(const __half& a, const __half& b) { return __half((float)a - (float)b); }
inline __half operator*(const __half& a, const __half& b) { return __half((float)a * (float)b); }
inline __half operator/(const __half& a, const __half& b) { return __half((float)a / (float)b); }
inline __half& operator+=(__half& a, const __half& b) { a = a + b; return a; }
inline __half& operator-=(__half& a, const __half& b) { a = a - b; return a; }
inline float __expf(float x) { return expf(x); }
inline float __frcp_rn(float x) { return 1.0f / x; }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t atomicMax(uint32_t* p, uint32_t v) { uint32_t old = *p; if (v > old) *p = v; return old; }
template <typename T, typename U> inline T atomicAdd(T* p, U v) { T old = *p; *p = old + (T)v; return old; }
// CUDA's global min / max overloads
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
inline uint32_t min(uint32_t a, int b) { return min(a, (uint32_t)b); }
inline uint32_t min(int a, uint32_t b) { return min((uint32_t)a, b); }
inline uint32_t max(uint32_t a, int b) { return max(a, (uint32_t)b); }
inline uint32_t max(int a, uint32_t b) { return max((uint32_t)a, b); }
inline float min(float a, double b) { return fminf(a, (float)b); }
inline float max(float a, double b) { return fmaxf(a, (float)b); }

struct float2 { float x, y; }; struct float3 { float x, y, z; }; struct float4 { float x, y, z, w; };
struct __half2 { __half x, y; };
// a "surface" is a plain float4 image here: tonemap_kernel's surf2Dwrite lands in host memory
struct RefHostSurface { float4* data; int width; };
typedef RefHostSurface* cudaSurfaceObject_t;
inline void surf2Dwrite(float4 v, cudaSurfaceObject_t s, size_t x_bytes, int y) { s->data[(size_t)y * s->width + x_bytes / sizeof(float4)] = v; }
#define TCNN_HOST_DEVICE
#define TCNN_NAMESPACE_BEGIN namespace tcnn {
#define TCNN_NAMESPACE_END }
inline void sincosf_(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
namespace tcnn {
using network_precision_t = __half;
template <typename T> inline void host_device_swap(T& a, T& b) { T c(a); a = b; b = c; }
template <typename T> inline T clamp(T v, T lo, T hi) { return v < lo ? lo : (hi < v ? hi : v); }   // tcnn common.h:250 (same comparison order as the oracle's clampf)
#ifndef REF_HOST_FUNCTIONAL_GPU_MEMORY   // oracle/ref_opt_shim brings a working one
template <typename T> struct GPUMemory { T* data() const { return nullptr; } size_t size() const { return 0; } };   // never allocated on this path
#endif
template <typename T> inline T next_multiple(T val, T divisor) { return (val + divisor - 1) / divisor * divisor; }
template <typename T> inline T div_round_up(T val, T divisor) { return (val + divisor - 1) / divisor; }
template <typename T>
struct PitchedPtr {   // pointer + byte stride, as tcnn::PitchedPtr (common.h:433-460)
	T* ptr; uint32_t stride_in_bytes;
	PitchedPtr() : ptr(nullptr), stride_in_bytes(sizeof(T)) {}
	PitchedPtr(T* p, size_t stride_in_elements, size_t offset = 0, size_t extra_stride_bytes = 0) : ptr(p + offset), stride_in_bytes((uint32_t)(stride_in_elements * sizeof(T) + extra_stride_bytes)) {}
	template <typename U> explicit PitchedPtr(PitchedPtr<U> o) : ptr((T*)o.ptr), stride_in_bytes(o.stride_in_bytes) {}
	T* operator()(uint32_t y) const { return (T*)((const char*)ptr + y * stride_in_bytes); }
	void operator+=(uint32_t y) { ptr = (T*)((const char*)ptr + y * stride_in_bytes); }
	void operator-=(uint32_t y) { ptr = (T*)((const char*)ptr - y * stride_in_bytes); }
	explicit operator bool() const { return ptr; }
};
// Morton helpers: stand-ins (the published bit tricks); the oracle's own copies are pinned bit for bit against the reference's compiled
// tiny-cuda-nn in tests/test_ref_kernels.py
inline uint32_t expand_bits(uint32_t v) { v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu; v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u; return v; }
inline uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) { return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2); }
inline uint32_t morton3D_invert(uint32_t x) { x = x & 0x49249249; x = (x | (x >> 2)) & 0xc30c30c3; x = (x | (x >> 4)) & 0x0f00f00f; x = (x | (x >> 8)) & 0xff0000ff; x = (x | (x >> 16)) & 0x0000ffff; return x; }
template <typename T, uint32_t N> struct vector_t { T data[N]; T& operator[](uint32_t i) { return data[i]; } const T& operator[](uint32_t i) const { return data[i]; } };
template <typename T>
struct MatrixView {   // same indexing as tcnn::MatrixView: element (i, j) at data[i * stride_i + j * stride_j]
	T* data = nullptr; uint32_t stride_i = 0, stride_j = 0;
	MatrixView() {}
	MatrixView(T* d, uint32_t si, uint32_t sj) : data(d), stride_i(si), stride_j(sj) {}
	T& operator()(uint32_t i, uint32_t j = 0) const { return data[i * stride_i + j * stride_j]; }
	void advance(uint32_t m, uint32_t n) { data = &(*this)(m, n); }
};
template <typename T> inline T logistic(const T x) { return T(1.0f / (1.0f + expf(-(float)x))); }
}

// ---- the subset of Eigen the reference headers use --------------------------------------------------------------------------
#include <Eigen/Dense>
