// umma_probe.cu -- bring-up probe for the tcgen05 pieces the training kernel's in-kernel weight gradients rely on (sm_100a).
//
// What it pins down on the hardware, with integer-valued fp16 data so that every fp32 sum is exact:
//   1. shared-memory descriptors for MN-major operands without swizzle: element (mn, k) of a [MN x 16] tile lives at
//      (mn / 8) * SBO + (k / 8) * LBO + (k % 8) * 16 + (mn % 8) * 2 bytes (8 x 8 core matrices of 128 contiguous bytes);
//   2. the instruction descriptor of kind::f16 with fp32 accumulation, A and B both MN-major, M = 64, N in {16, 32, 48, 64};
//   3. where a M = 64 accumulator lives in tensor memory: row m -> lane (m % 16) + 32 * (m / 16), column = n;
//   4. accumulation across instructions and the tcgen05.commit -> mbarrier completion path;
//   5. (stress) several warps of one CTA issuing tcgen05.mma into the SAME accumulator without ordering between them: the sum
//      over all of them must still be exact.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu ; run: tools/umma_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
	uint32_t ok;
	asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	uint32_t spins = 0;
	while (!mbar_try_wait(bar, parity)) { if (++spins > (1u << 26)) { printf("mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); } }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
	             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
	               "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// MN-major, no swizzle.  Descriptor fields are in 16-byte units; bit 46 = descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
	return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__host__ __device__ inline uint32_t make_idesc(uint32_t M, uint32_t N) {
	return (1u << 4) /* D = f32 */ | (0u << 7) /* A = f16 */ | (0u << 10) /* B = f16 */ | (1u << 15) /* A MN-major */ | (1u << 16) /* B MN-major */ | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__host__ __device__ inline uint32_t canon_off(uint32_t k, uint32_t mn) { return (mn >> 3) * 256u + (k >> 3) * 128u + (k & 7u) * 16u + (mn & 7u) * 2u; }

#define TMEM_COLS 256

// A: [iters][16 (k)][64 (m)] halves, B: [iters][16 (k)][N (n)] halves, both "sample-major" as the training kernel produces them.
// dump: [128 lanes][TMEM_COLS] floats (raw tensor memory after the last MMA).
__global__ void __launch_bounds__(128) k_probe(const __half* __restrict__ A, const __half* __restrict__ B, int N, int iters, uint32_t lbo, uint32_t sbo, uint32_t col0,
                                               float* __restrict__ dump) {
	__shared__ __align__(128) unsigned char sA[2048];
	__shared__ __align__(128) unsigned char sB[2048];
	__shared__ __align__(8) uint64_t bar;
	__shared__ uint32_t tmem_base_s;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"((uint32_t)TMEM_COLS) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = tmem_base_s;
	const uint32_t idesc = make_idesc(64, (uint32_t)N);
	for (int it = 0; it < iters; ++it) {
		for (int i = tid; i < 16 * 64; i += 128) { const int k = i >> 6, m = i & 63; *reinterpret_cast<__half*>(sA + canon_off(k, m)) = A[(size_t)it * 1024 + i]; }
		for (int i = tid; i < 16 * N; i += 128) { const int k = i / N, n = i - k * N; *reinterpret_cast<__half*>(sB + canon_off(k, n)) = B[(size_t)it * 16 * N + i]; }
		fence_proxy_async();
		tc_fence_before();
		__syncthreads();
		if (tid == 0) {
			tc_fence_after();
			tc_mma_f16(tmem + col0, make_desc(smem_u32(sA), lbo, sbo), make_desc(smem_u32(sB), lbo, sbo), idesc, it > 0 ? 1u : 0u);
			tc_commit(&bar);
		}
		mbar_wait(&bar, (uint32_t)(it & 1));
		tc_fence_after();
		__syncthreads();
	}
	for (int c = 0; c < TMEM_COLS; c += 16) {
		uint32_t r[16];
		tc_ld16(tmem + ((uint32_t)(32 * warp) << 16) + c, r);
		for (int j = 0; j < 16; ++j) dump[(size_t)(32 * warp + lane) * TMEM_COLS + c + j] = __uint_as_float(r[j]);
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)TMEM_COLS) : "memory");
}

// Stress: W warps, each with its own staging and its own mbarrier, issue their MMAs (N = 64) into the SAME accumulator with no
// ordering between warps.  A, B: [warps][iters][16][64].  out: [gridDim][64][64] floats.
#define STRESS_WARPS 12
__global__ void __launch_bounds__(STRESS_WARPS * 32) k_stress(const __half* __restrict__ A, const __half* __restrict__ B, int iters, int n_tiles_total, float* __restrict__ out) {
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ __align__(8) uint64_t bars[STRESS_WARPS + 1];
	__shared__ uint32_t tmem_base_s;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	unsigned char* sA = smem + warp * 4096;
	unsigned char* sB = sA + 2048;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"((uint32_t)64) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	if (tid == 0) { for (int i = 0; i <= STRESS_WARPS; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = tmem_base_s;
	const uint32_t idesc = make_idesc(64, 64);
	// zero the accumulator: one MMA with a zero A tile and accumulate = 0, completed before anybody else starts
	if (warp == 0) {
		for (int i = lane; i < 1024; i += 32) { reinterpret_cast<uint32_t*>(sA)[i & 511] = 0u; reinterpret_cast<uint32_t*>(sB)[i & 511] = 0u; }
		fence_proxy_async();
		__syncwarp();
		if (lane == 0) { tc_fence_after(); tc_mma_f16(tmem, make_desc(smem_u32(sA), 128, 256), make_desc(smem_u32(sB), 128, 256), idesc, 0u); tc_commit(&bars[STRESS_WARPS]); }
		mbar_wait(&bars[STRESS_WARPS], 0);
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	for (int it = 0; it < iters; ++it) {
		const size_t tile = ((size_t)blockIdx.x * STRESS_WARPS + warp) * iters + it;
		const __half* a = A + (tile % n_tiles_total) * 1024;
		const __half* b = B + (tile % n_tiles_total) * 1024;
		if (it > 0) mbar_wait(&bars[warp], (uint32_t)((it - 1) & 1));     // the previous MMA of this warp has read the staging
		for (int i = lane; i < 1024; i += 32) {
			const int k = i >> 6, m = i & 63;
			*reinterpret_cast<__half*>(sA + canon_off(k, m)) = a[i];
			*reinterpret_cast<__half*>(sB + canon_off(k, m)) = b[i];
		}
		fence_proxy_async();
		__syncwarp();
		if (lane == 0) {
			tc_fence_after();
			tc_mma_f16(tmem, make_desc(smem_u32(sA), 128, 256), make_desc(smem_u32(sB), 128, 256), idesc, 1u);
			tc_commit(&bars[warp]);
		}
		__syncwarp();
	}
	if (iters > 0) mbar_wait(&bars[warp], (uint32_t)((iters - 1) & 1));
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	if (warp < 4) {
		for (int c = 0; c < 64; c += 16) {
			uint32_t r[16];
			tc_ld16(tmem + ((uint32_t)(32 * warp) << 16) + c, r);
			if (lane < 16) for (int j = 0; j < 16; ++j) out[((size_t)blockIdx.x * 64 + 16 * warp + lane) * 64 + c + j] = __uint_as_float(r[j]);
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)64) : "memory");
}


// Timing of small MMAs (M = 64, N = n, K = 16) on one CTA per SM.  mode 0: ONE thread issues `iters` MMAs back to back, one commit at the
// end.  mode 1: one thread, a commit behind every MMA, one wait at the end.  mode 2: `warps` warps, each issues MMA + commit and waits
// for its own completion before the next (the training kernel's pattern), all into the same accumulator.  mode 3: like 2, but every warp
// has its own accumulator columns.  mode 4: like 2 with a dedicated pattern of 4 MMAs per commit.
__global__ void __launch_bounds__(STRESS_WARPS * 32) k_timing(int mode, int warps, int n, int iters, float* __restrict__ sink) {
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ __align__(8) uint64_t bars[STRESS_WARPS + 1];
	__shared__ uint32_t tmem_base_s;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	unsigned char* sA = smem + warp * 4096;
	unsigned char* sB = sA + 2048;
	for (int i = tid; i < STRESS_WARPS * 4096 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // halves 1.0
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"((uint32_t)512) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	if (tid == 0) { for (int i = 0; i <= STRESS_WARPS; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
	fence_proxy_async();
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = tmem_base_s;
	const uint32_t idesc = make_idesc(64, (uint32_t)n);
	const uint64_t da = make_desc(smem_u32(sA), 128, 256), db = make_desc(smem_u32(sB), 128, 256);
	if (mode <= 1) {
		if (tid == 0) {
			for (int it = 0; it < iters; ++it) {
				tc_mma_f16(tmem, da, db, idesc, it > 0 ? 1u : 0u);
				if (mode == 1) tc_commit(&bars[1 + (it & 7)]);
			}
			tc_commit(&bars[0]);
		}
		if (warp == 0) mbar_wait(&bars[0], 0);
	} else if (warp < warps) {
		const uint32_t col = mode == 3 ? (uint32_t)(warp * 32) : 0u;
		const int per = mode == 4 ? 4 : 1;
		for (int it = 0; it < iters; ++it) {
			if (lane == 0) {
				for (int j = 0; j < per; ++j) tc_mma_f16(tmem + col, da, db, idesc, 1u);
				tc_commit(&bars[warp]);
			}
			__syncwarp();
			mbar_wait(&bars[warp], (uint32_t)(it & 1));
		}
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	if (warp == 0) {
		uint32_t r[16];
		tc_ld16(tmem, r);
		if (lane == 0) sink[blockIdx.x] = __uint_as_float(r[0]);
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)512) : "memory");
}

static void run_timing(int mode, int warps, int n, int iters) {
	float* sink; CK(cudaMalloc(&sink, 148 * 4));
	CK(cudaFuncSetAttribute(k_timing, cudaFuncAttributeMaxDynamicSharedMemorySize, STRESS_WARPS * 4096));
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int rep = 0; rep < 2; ++rep) {
		cudaEventRecord(e0);
		k_timing<<<148, STRESS_WARPS * 32, STRESS_WARPS * 4096>>>(mode, warps, n, iters, sink);
		cudaEventRecord(e1);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) { printf("timing: kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
	}
	float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
	const double mmas = (mode <= 1 ? 1.0 : (double)warps * (mode == 4 ? 4 : 1)) * iters;
	printf("timing mode %d warps %2d N=%2d: %.3f ms for %.0f MMAs per CTA -> %.1f ns per MMA (%.0f cycles at 1.965 GHz)\n", mode, warps, n, ms, mmas, ms * 1e6 / mmas, ms * 1e6 / mmas * 1.965);
	cudaFree(sink);
}

// Do legacy mma.sync (HMMA) and tcgen05.mma share the tensor cores gracefully?  Warps 1..11 run a fixed amount of mma.sync work
// (independent accumulators); warp 0 meanwhile issues `umma_groups` groups of (MMA + commit + wait), spaced by `gap` clock cycles.
__global__ void __launch_bounds__(STRESS_WARPS * 32) k_interfere(int hmma_iters, int umma_groups, int gap, float* __restrict__ sink, long long* __restrict__ cycles) {
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ __align__(8) uint64_t bars[2];
	__shared__ uint32_t tmem_base_s;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	for (int i = tid; i < 4096 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"((uint32_t)64) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	if (tid == 0) { mbar_init(&bars[0], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
	fence_proxy_async();
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = tmem_base_s;
	const long long t0 = clock64();
	if (warp == 0) {
		const uint64_t da = make_desc(smem_u32(smem), 128, 256), db = make_desc(smem_u32(smem + 2048), 128, 256);
		for (int it = 0; it < umma_groups; ++it) {
			if (lane == 0) { tc_mma_f16(tmem, da, db, make_idesc(64, 64), it > 0 ? 1u : 0u); tc_commit(&bars[0]); }
			__syncwarp();
			mbar_wait(&bars[0], (uint32_t)(it & 1));
			const long long w0 = clock64();
			while (clock64() - w0 < gap) { }
		}
	} else {
		float c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0}, c3[4] = {0, 0, 0, 0};
		uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u};
		const uint32_t b0 = 0x3c003c00u + lane, b1 = 0x3c003c00u;
		for (int it = 0; it < hmma_iters; ++it) {
			asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+f"(c0[0]), "+f"(c0[1]), "+f"(c0[2]), "+f"(c0[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
			asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+f"(c1[0]), "+f"(c1[1]), "+f"(c1[2]), "+f"(c1[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
			asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+f"(c2[0]), "+f"(c2[1]), "+f"(c2[2]), "+f"(c2[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
			asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+f"(c3[0]), "+f"(c3[1]), "+f"(c3[2]), "+f"(c3[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
		}
		if (c0[0] + c1[0] + c2[0] + c3[0] == 123.456f) sink[0] = 1.f;
		if (lane == 0 && warp == 1 && blockIdx.x == 0) cycles[0] = clock64() - t0;
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)64) : "memory");
}

static void run_interfere(int hmma_iters, int umma_groups, int gap) {
	float* sink; long long* cyc; CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&cyc, 8));
	CK(cudaFuncSetAttribute(k_interfere, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096));
	for (int rep = 0; rep < 2; ++rep) {
		k_interfere<<<148, STRESS_WARPS * 32, 4096>>>(hmma_iters, umma_groups, gap, sink, cyc);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) { printf("interfere: kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
	}
	long long h = 0; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
	printf("interfere: 11 warps x %d x 4 mma.sync beside %d tcgen05 groups (gap %d cycles): HMMA warp finished after %lld cycles (%.1f cycles per mma.sync per warp)\n",
	       hmma_iters, umma_groups, gap, h, (double)h / (4.0 * hmma_iters));
	cudaFree(sink); cudaFree(cyc);
}

static uint32_t rng_state = 12345u;
static int rnd_small() { rng_state = rng_state * 1664525u + 1013904223u; return (int)((rng_state >> 24) % 7) - 3; }

static bool run_probe(int N, int iters, uint32_t lbo, uint32_t sbo, uint32_t col0, bool verbose) {
	std::vector<__half> hA((size_t)iters * 1024), hB((size_t)iters * 16 * N);
	std::vector<int> iA(hA.size()), iB(hB.size());
	for (size_t i = 0; i < hA.size(); ++i) { iA[i] = rnd_small(); hA[i] = __float2half((float)iA[i]); }
	for (size_t i = 0; i < hB.size(); ++i) { iB[i] = rnd_small(); hB[i] = __float2half((float)iB[i]); }
	__half *dA, *dB; float* dDump;
	CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dDump, 128 * TMEM_COLS * 4));
	CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
	CK(cudaMemset(dDump, 0xff, 128 * TMEM_COLS * 4));
	k_probe<<<1, 128>>>(dA, dB, N, iters, lbo, sbo, col0, dDump);
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) { printf("probe N=%d lbo=%u sbo=%u: kernel failed: %s\n", N, lbo, sbo, cudaGetErrorString(e)); exit(3); }
	std::vector<float> dump(128 * TMEM_COLS);
	CK(cudaMemcpy(dump.data(), dDump, dump.size() * 4, cudaMemcpyDeviceToHost));
	std::vector<long> ref((size_t)64 * N, 0);
	for (int it = 0; it < iters; ++it) for (int k = 0; k < 16; ++k) for (int m = 0; m < 64; ++m) for (int n = 0; n < N; ++n)
		ref[(size_t)m * N + n] += (long)iA[(size_t)it * 1024 + k * 64 + m] * iB[((size_t)it * 16 + k) * N + n];
	long bad = 0;
	for (int m = 0; m < 64; ++m) for (int n = 0; n < N; ++n) {
		const int lane = (m % 16) + 32 * (m / 16);
		if (dump[(size_t)lane * TMEM_COLS + col0 + n] != (float)ref[(size_t)m * N + n]) ++bad;
	}
	printf("probe M=64 N=%d iters=%d lbo=%u sbo=%u col0=%u: %s (%ld of %d mismatches)\n", N, iters, lbo, sbo, col0, bad ? "FAIL" : "PASS", bad, 64 * N);
	if (bad && verbose) {
		// where did row 0 / row 17 go?  print the first columns of a few lanes next to the expectation
		for (int m : {0, 1, 17, 40}) { printf("  ref row %2d:", m); for (int n = 0; n < 8; ++n) printf(" %5ld", ref[(size_t)m * N + n]); printf("\n"); }
		for (int lane : {0, 1, 16, 17, 32, 33, 49, 64, 72}) { printf("  lane %3d:", lane); for (int n = 0; n < 8; ++n) printf(" %7.0f", dump[(size_t)lane * TMEM_COLS + col0 + n]); printf("\n"); }
	}
	cudaFree(dA); cudaFree(dB); cudaFree(dDump);
	return bad == 0;
}

static bool run_stress(int blocks, int iters) {
	const int n_tiles = 4096;
	std::vector<__half> hA((size_t)n_tiles * 1024), hB((size_t)n_tiles * 1024);
	std::vector<signed char> iA(hA.size()), iB(hB.size());
	for (size_t i = 0; i < hA.size(); ++i) { iA[i] = (signed char)rnd_small(); hA[i] = __float2half((float)iA[i]); iB[i] = (signed char)rnd_small(); hB[i] = __float2half((float)iB[i]); }
	__half *dA, *dB; float* dOut;
	CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dOut, (size_t)blocks * 4096 * 4));
	CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
	CK(cudaFuncSetAttribute(k_stress, cudaFuncAttributeMaxDynamicSharedMemorySize, STRESS_WARPS * 4096));
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	k_stress<<<blocks, STRESS_WARPS * 32, STRESS_WARPS * 4096>>>(dA, dB, iters, n_tiles, dOut);
	cudaEventRecord(e1);
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) { printf("stress: kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
	float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
	std::vector<float> out((size_t)blocks * 4096);
	CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
	// per-tile products once, then sum per block
	std::vector<int> prod((size_t)n_tiles * 4096);
	for (int t = 0; t < n_tiles; ++t) for (int m = 0; m < 64; ++m) for (int n = 0; n < 64; ++n) {
		int s = 0;
		for (int k = 0; k < 16; ++k) s += (int)iA[(size_t)t * 1024 + k * 64 + m] * iB[(size_t)t * 1024 + k * 64 + n];
		prod[(size_t)t * 4096 + m * 64 + n] = s;
	}
	long bad = 0;
	std::vector<int> count(n_tiles);
	for (int b = 0; b < blocks; ++b) {
		std::fill(count.begin(), count.end(), 0);
		for (int w = 0; w < STRESS_WARPS; ++w) for (int it = 0; it < iters; ++it) ++count[(((size_t)b * STRESS_WARPS + w) * iters + it) % n_tiles];
		std::vector<long> ref(4096, 0);
		for (int t = 0; t < n_tiles; ++t) if (count[t]) { const int c = count[t]; const int* p = &prod[(size_t)t * 4096]; for (int i = 0; i < 4096; ++i) ref[i] += (long)c * p[i]; }
		for (int i = 0; i < 4096; ++i) if (out[(size_t)b * 4096 + i] != (float)ref[i]) ++bad;
	}
	printf("stress: %d CTAs x %d warps x %d unordered MMAs into one accumulator: %s (%ld mismatching elements), %.3f ms, %.1f ns per MMA per CTA\n",
	       blocks, STRESS_WARPS, iters, bad ? "FAIL" : "PASS", bad, ms, ms * 1e6 / ((double)STRESS_WARPS * iters));
	cudaFree(dA); cudaFree(dB); cudaFree(dOut);
	return bad == 0;
}

int main(int argc, char** argv) {
	bool ok = true;
	// hypothesis 1: LBO = stride between the two 8-deep K blocks (128 B), SBO = stride between 8-wide MN blocks (256 B)
	bool h1 = run_probe(64, 1, 128, 256, 0, true);
	if (!h1) {
		printf("hypothesis 1 failed; trying the swapped meaning of LBO / SBO\n");
		bool h2 = run_probe(64, 1, 256, 128, 0, true);
		printf("swapped meaning: %s\n", h2 ? "PASS" : "FAIL");
		return 1;
	}
	for (int N : {16, 32, 48, 64}) ok &= run_probe(N, 3, 128, 256, 0, true);
	ok &= run_probe(16, 5, 128, 256, 64, true);      // accumulator at a column offset
	ok &= run_probe(48, 4, 128, 256, 80, true);
	ok &= run_probe(32, 7, 128, 256, 208, true);
	for (int n : {16, 32, 64}) { run_timing(0, 1, n, 20000); run_timing(1, 1, n, 20000); }
	for (int w : {1, 4, 12}) { run_timing(2, w, 64, 4000); run_timing(3, w, 32, 4000); }
	run_timing(4, 12, 64, 2000);
	run_interfere(20000, 0, 0);
	run_interfere(20000, 100000, 0);        // tcgen05 groups back to back for the whole duration
	run_interfere(20000, 100000, 2000);     // one group per ~1 us
	run_interfere(20000, 100000, 20000);    // one group per ~10 us
	ok &= run_stress(148, 400);
	ok &= run_stress(148, 2000);
	printf(ok ? "ALL PASS\n" : "SOME FAILED\n");
	return ok ? 0 : 1;
}
