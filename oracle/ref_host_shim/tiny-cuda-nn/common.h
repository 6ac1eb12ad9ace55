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
template <typename T> inline T atomicAdd(T* p, T v) { T old = *p; *p = old + v; return old; }

namespace tcnn {
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

// ---- the subset of Eigen the header uses ------------------------------------------------------------------------------------
namespace Eigen {
enum { ColMajor = 0, RowMajor = 1 };
struct Vector3f {
	float v[3];
	Vector3f() : v{0.f, 0.f, 0.f} {}
	Vector3f(float a, float b, float c) : v{a, b, c} {}
	float& operator[](int i) { return v[i]; }
	const float& operator[](int i) const { return v[i]; }
	float dot(const Vector3f& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
	float squaredNorm() const { return dot(*this); }
	float norm() const { return std::sqrt(squaredNorm()); }
	Vector3f cross(const Vector3f& o) const { return Vector3f(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]); }
	Vector3f operator+(const Vector3f& o) const { return Vector3f(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
	Vector3f operator-(const Vector3f& o) const { return Vector3f(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
	Vector3f operator*(float s) const { return Vector3f(v[0] * s, v[1] * s, v[2] * s); }
	Vector3f operator/(float s) const { return Vector3f(v[0] / s, v[1] / s, v[2] / s); }
	Vector3f& operator+=(const Vector3f& o) { for (int k = 0; k < 3; ++k) v[k] += o.v[k]; return *this; }
	Vector3f& operator-=(const Vector3f& o) { for (int k = 0; k < 3; ++k) v[k] -= o.v[k]; return *this; }
	Vector3f normalized() const { const float z = squaredNorm(); return z > 0.f ? *this / std::sqrt(z) : *this; }   // Eigen: v / sqrt(squaredNorm) when > 0
	static Vector3f Constant(float c) { return Vector3f(c, c, c); }
	static Vector3f Ones() { return Constant(1.f); }
	static Vector3f Zero() { return Constant(0.f); }
};
inline Vector3f operator*(float s, const Vector3f& a) { return a * s; }

template <int Order>
struct Mat33 {
	float d[9];
	Mat33() { for (float& x : d) x = 0.f; }
	template <int O2> Mat33(const Mat33<O2>& o) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) (*this)(r, c) = o(r, c); }
	float& operator()(int r, int c) { return d[Order == RowMajor ? r * 3 + c : c * 3 + r]; }
	const float& operator()(int r, int c) const { return d[Order == RowMajor ? r * 3 + c : c * 3 + r]; }
	float* data() { return d; }
	const float* data() const { return d; }
	struct Comma {   // Eigen's comma initialiser fills row by row whatever the storage order
		Mat33& m; int k;
		Comma& operator,(float x) { m(k / 3, k % 3) = x; ++k; return *this; }
	};
	Comma operator<<(float x) { (*this)(0, 0) = x; return Comma{*this, 1}; }
	Mat33 inverse() const {   // cofactor formula, as Eigen's fixed-size 3x3 inverse
		const Mat33& a = *this;
		Mat33 c;
		c(0, 0) = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1); c(0, 1) = a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2); c(0, 2) = a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1);
		c(1, 0) = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2); c(1, 1) = a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0); c(1, 2) = a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2);
		c(2, 0) = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0); c(2, 1) = a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1); c(2, 2) = a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0);
		const float det = a(0, 0) * c(0, 0) + a(0, 1) * c(1, 0) + a(0, 2) * c(2, 0);
		const float inv = 1.0f / det;
		for (float& x : c.d) x *= inv;
		return c;
	}
};
template <int O>
inline Vector3f operator*(const Mat33<O>& m, const Vector3f& x) {
	return Vector3f(m(0, 0) * x[0] + m(0, 1) * x[1] + m(0, 2) * x[2], m(1, 0) * x[0] + m(1, 1) * x[1] + m(1, 2) * x[2], m(2, 0) * x[0] + m(2, 1) * x[1] + m(2, 2) * x[2]);
}
template <int A, int B>
inline Mat33<A> operator*(const Mat33<A>& a, const Mat33<B>& b) {
	Mat33<A> r;
	for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = a(i, 0) * b(0, j) + a(i, 1) * b(1, j) + a(i, 2) * b(2, j);
	return r;
}
using Matrix3f = Mat33<ColMajor>;
template <typename Scalar, int Rows, int Cols, int Order = ColMajor> using Matrix = Mat33<Order>;   // only float 3x3 is ever named
}
