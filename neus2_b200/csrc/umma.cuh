// umma.cuh -- the tcgen05 / tensor-memory pieces used by the training kernel's in-kernel weight gradients (sm_100a).
//
// dW = sum over samples of dY^T X is a GEMM whose reduction dimension is the batch: every warp tile of 16 samples is one K = 16
// step.  Both operands come out of the warp's registers "sample-major" (row = sample, contiguous = feature), which is exactly the
// MN-major operand form of tcgen05.mma: A = dY^T is [M = 64 features][K = 16 samples] with M contiguous, B = X is [K][N] with N
// contiguous.  Operands are written to shared memory in the canonical no-swizzle layout (8 x 8 core matrices of 128 contiguous
// bytes), the accumulators live in tensor memory for the whole kernel and are read back once.
//
// Every constant below (descriptor bit fields, the canonical layout, lane mapping of an M = 64 accumulator, several warps of one
// CTA issuing into one accumulator) is pinned on the hardware by tools/umma_probe.cu with exact integer sums.
#pragma once
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier (completion of asynchronous tensor-core work) ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
	uint32_t ok;
	asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	return ok != 0;
}
// A wait that can never complete (a malformed instruction stream during bring-up) must not hang the GPU: trap after ~1 s of polling.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	uint32_t spins = 0;
	while (!mbar_try_wait(bar, parity)) { if (++spins > (1u << 27)) __trap(); }
}

// ---- proxies and ordering ----
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }   // generic-proxy smem writes -> visible to the tensor core
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- tensor memory ----
// One full warp allocates `cols` columns (power of two >= 32) and writes the base address to *slot (shared memory).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(cols) : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(cols) : "memory"); }
// 32 lanes x 16 columns of 32-bit: thread i of the warp gets lane (lane field of taddr) + i, columns [col, col + 16).
// A warp can only read the 32-lane quadrant (warp index % 4) of tensor memory.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
	uint32_t r[16];
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
	             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
	               "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
	#pragma unroll
	for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

// ---- operands ----
// Canonical MN-major layout without swizzle of a [16 samples (K)][F features (MN)] half tile: feature blocks of 8 are 256 bytes
// apart (SBO), the two sample blocks of 8 are 128 bytes apart (LBO), a sample's 8 features are 16 contiguous bytes.
#define UMMA_LBO 128u
#define UMMA_SBO 256u
__device__ __forceinline__ uint32_t canon_off(uint32_t sample, uint32_t feature) { return (feature >> 3) * UMMA_SBO + (sample >> 3) * UMMA_LBO + (sample & 7u) * 16u + (feature & 7u) * 2u; }
// shared-memory matrix descriptor: 16-byte units; bit 46 = descriptor version of this architecture; layout type 0 = no swizzle
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
	return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(UMMA_LBO >> 4) << 16) | ((uint64_t)(UMMA_SBO >> 4) << 32) | (1ull << 46);
}
// instruction descriptor of kind::f16: D fp32, A and B fp16, both MN-major, M x N (K = 16)
__host__ __device__ constexpr uint32_t make_idesc(uint32_t M, uint32_t N) {
	return (1u << 4) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem], issued by ONE thread
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// the mbarrier gets one arrival when every tcgen05.mma this thread issued before has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory"); }

// mma.sync A fragments (16 samples x 16*KT features, thread (g, t): rows g / g+8, feature pairs 2t / 2t+8) -> canonical tile.
// Each register is one conflict-free 128-byte-wide store of the warp.
template <int KT>
__device__ __forceinline__ void stage_frags(const uint32_t (&a)[KT][4], unsigned char* S, int g, int t) {
	#pragma unroll
	for (int kt = 0; kt < KT; ++kt) {
		unsigned char* p = S + (2 * kt) * UMMA_SBO + g * 16 + 4 * t;
		*reinterpret_cast<uint32_t*>(p) = a[kt][0];
		*reinterpret_cast<uint32_t*>(p + UMMA_LBO) = a[kt][1];
		*reinterpret_cast<uint32_t*>(p + UMMA_SBO) = a[kt][2];
		*reinterpret_cast<uint32_t*>(p + UMMA_SBO + UMMA_LBO) = a[kt][3];
	}
}
// 32 features per sample of a row-major shared tile (row stride ld_bytes, first feature at src) -> canonical feature blocks
// blk0 .. blk0 + 3.  lane -> (sample = lane >> 1, 16 features = lane & 1): two 16-byte chunks per lane.
__device__ __forceinline__ void stage_rows32(const unsigned char* src, int ld_bytes, unsigned char* S, int blk0, int lane) {
	const int s = lane >> 1, h = lane & 1;
	const uint4 v0 = *reinterpret_cast<const uint4*>(src + s * ld_bytes + h * 32);
	const uint4 v1 = *reinterpret_cast<const uint4*>(src + s * ld_bytes + h * 32 + 16);
	unsigned char* p = S + (blk0 + 2 * h) * UMMA_SBO + (s >> 3) * UMMA_LBO + (s & 7) * 16;
	*reinterpret_cast<uint4*>(p) = v0;
	*reinterpret_cast<uint4*>(p + UMMA_SBO) = v1;
}

}  // namespace umma
