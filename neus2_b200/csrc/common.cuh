// common.cuh -- shared device helpers for the neus2_b200 kernels (sm_100a).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define NEUS_MAX_LEVELS 16

// Per-level hash-grid tables (reference: tcnn encodings/grid.h:1472-1508), passed by value to kernels.
struct GridMeta {
	uint32_t n_levels;
	uint32_t valid_level;                 // levels > valid_level are masked (grid.h:198, 2427-2440)
	uint32_t hashed_mask;                 // bit l set <=> level l is hashed (res^3 > hashmap size; grid.h:143-150)
	uint32_t offset[NEUS_MAX_LEVELS + 1]; // in entries (x2 features)
	uint32_t res[NEUS_MAX_LEVELS];
	float scale[NEUS_MAX_LEVELS];         // == res-1 (integer valued, grid.h:1477)
	const float* delta_xform;             // DeltaNetwork in use (m_use_delta): device float[DELTA_XFORM_FLOATS], else nullptr
};

// Device-resident float image of the DeltaNetwork's 12 half parameters (transform_network.h:23-480), rebuilt after every
// update of them: R row-major (columns = Gram-Schmidt of the 6d rotation), t, t - first_frame_offset, first_frame_offset,
// and the Gram-Schmidt inputs the backward needs (a1, a2, u2 = a2 - (b1.a2) b1).
#define DELTA_R 0
#define DELTA_T 9
#define DELTA_TMO 12
#define DELTA_OFF 15
#define DELTA_A1 18
#define DELTA_A2 21
#define DELTA_U2 24
#define DELTA_XFORM_FLOATS 28

// Parameter layout in the reference order (nerf_network.h:741-785): [sdf-MLP | colour-MLP | grid | variance(4)].
struct ParamLayout {
	uint32_t sdf_in_w;     // 32
	uint32_t off_w0, off_w1;            // sdf MLP: W0 [64][sdf_in_w], W1 [16][64]
	uint32_t off_c0, off_c1, off_c2;    // colour MLP: C0 [64][48], C1 [64][64], C2 [16][64]
	uint32_t off_grid;
	uint32_t off_var;
	uint32_t n_params;
};

#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { neus_set_error(cudaGetErrorString(e_), __FILE__, __LINE__); return -1; } } while (0)
void neus_set_error(const char* msg, const char* file, int line);
extern unsigned long long g_neus_launches;   // kernels launched by this library (bench.py's gpu_launches)
#define COUNT_LAUNCH(n) (g_neus_launches += (n))

#ifdef __CUDACC__

// cudaFuncSetAttribute is per DEVICE: a process that drives several GPUs (one context each) must opt every device in.
#define NEUS_MAX_DEVICES 64
template <typename K>
static inline int ensure_dynamic_smem(K kernel, int bytes, bool (&done)[NEUS_MAX_DEVICES]) {
	int dev = 0;
	CUDA_CHECK(cudaGetDevice(&dev));
	if (dev < 0 || dev >= NEUS_MAX_DEVICES) { neus_set_error("device index out of range", __FILE__, __LINE__); return -1; }
	if (!done[dev]) { CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); done[dev] = true; }
	return 0;
}
static inline int device_sm_count() {
	int dev = 0, sms = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	return sms;
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
	__half2 h = __floats2half2_rn(a, b);
	return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float h_lo(uint32_t v) { return __low2float(*reinterpret_cast<__half2*>(&v)); }
__device__ __forceinline__ float h_hi(uint32_t v) { return __high2float(*reinterpret_cast<__half2*>(&v)); }
__device__ __forceinline__ float hbits2f(uint16_t b) { return __half2float(__ushort_as_half(b)); }
__device__ __forceinline__ uint16_t f2hbits(float f) { return __half_as_ushort(__float2half_rn(f)); }
__device__ __forceinline__ float round_h(float f) { return __half2float(__float2half_rn(f)); }

// D(16x8,f32) += A(16x16,f16,row) * B(16x8,f16,col)
__device__ __forceinline__ void mma16816(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
	asm volatile(
		"mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
		: "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
		: "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Reductions without a return value, spelled out: in a kernel that also contains tcgen05 / mbarrier inline assembly the compiler keeps
// atomicAdd() as the RETURNING form (ATOMG instead of REDG in SASS) and every hash-grid update then waits for its round trip to L2
// (measured: 0.057 -> 0.233 ms for the grid atomics of the training kernel).
__device__ __forceinline__ void red_add_f32(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add_v2f32(float* p, float a, float b) { asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(a), "f"(b) : "memory"); }
__device__ __forceinline__ void red_add_v4f32(float* p, float a, float b, float c, float d) {
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ uint32_t ld_nc_u32(const void* p) {
	uint32_t v;
	asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}

// tcnn pcg32 (dependencies/pcg32/pcg32.h) -- device/host restatement of the published generator.
struct Pcg32 {
	uint64_t state, inc;
	__host__ __device__ Pcg32() : state(0x853c49e6748fea9bULL), inc(0xda3e39cb94b95bdbULL) {}
	__host__ __device__ explicit Pcg32(uint64_t initstate, uint64_t initseq = 1u) {
		state = 0u; inc = (initseq << 1u) | 1u;
		next_uint(); state += initstate; next_uint();
	}
	__host__ __device__ uint32_t next_uint() {
		uint64_t old = state;
		state = old * 0x5851f42d4c957f2dULL + inc;
		uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
		uint32_t rot = (uint32_t)(old >> 59u);
		return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
	}
	__host__ __device__ float next_float() {
		union { uint32_t u; float f; } x;
		x.u = (next_uint() >> 9) | 0x3f800000u;
		return x.f - 1.0f;
	}
	__host__ __device__ void advance(int64_t delta_ = (1ll << 32)) {
		uint64_t cur_mult = 0x5851f42d4c957f2dULL, cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
		uint64_t delta = (uint64_t)delta_;
		while (delta > 0) {
			if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
			cur_plus = (cur_mult + 1) * cur_plus;
			cur_mult *= cur_mult;
			delta /= 2;
		}
		state = acc_mult * state + acc_plus;
	}
};

#endif  // __CUDACC__
