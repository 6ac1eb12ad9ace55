// optimizer.cu -- fused Adam (+ per-parameter step counters, sparse skip, L2 on the MLP weights) and fp16 EMA.
// Restates adam_step (tcnn optimizers/adam.h:52-160) and ema_step_half_precision (optimizers/ema.h:64-77,112-147)
// as ONE pass over the parameters (the reference launches two kernels and re-reads the fp16 params).
// Compiled with -fmad=false; the fused multiply-adds written out below are exactly the ones nvcc's default -fmad=true puts into
// the reference's kernels (read off their SASS), so the update is bit-identical to the reference's Ema{Adam} and to
// oracle/neus_oracle.cpp:orc_adam_ema_step.  The bias-correction factors (powf) come from a table filled by a kernel that lives in an
// FMA-enabled translation unit (api.cu), so that libdevice's powf is compiled the way it is inside the reference build.
#include <cmath>
#include "ngp_common.cuh"
#include "ngp_kernels.h"
#include "peer_sync.cuh"

// Optimizer state (fp32 master copy, both moments, step counters) and the EMA copy are touched once per step and by nobody else.  With
// ADAM_STREAM_HINTS=1 their loads and stores carry the streaming hint (ld / st.global.cs, evict-first in L2), so that the ~250 MB they amount to
// do not push the fp16 parameters and the gradient buffer -- which the next step's network kernels gather from and reduce into -- out of the L2.
#ifndef ADAM_STREAM_HINTS
#define ADAM_STREAM_HINTS 0   // measured (same box, profiles/r02_bench_n1_adam_hints.json / _no_hints.json): first pass 2.7 us faster, optimizer 3.6 us slower, step equal
#endif
#if ADAM_STREAM_HINTS
#define LD_STATE(p) __ldcs(p)
#define ST_STATE(p, v) __stcs(p, v)
#else
#define LD_STATE(p) (*(p))
#define ST_STATE(p, v) (*(p) = (v))
#endif

struct AdamHyper { float loss_scale, lr, beta1, beta2, eps, l2_reg, ema_decay, ema_debias_old, ema_debias_new;
                   uint32_t cur_step; const float* bias_table; uint32_t bias_table_size;
                   int zero_grads;   // also clear the gradients that were consumed (the next step starts from zeros; only touched quads are rewritten)
                   float inv_loss_scale; };   // 1 / loss_scale when loss_scale is a power of two (the division is then an exact scaling: same bits, no IEEE divide), else 0
// Small accumulators of the step that the next step expects cleared (variance / W1-row-0 sums, the step's counters): block 0 clears them.
struct AdamZero { uint32_t* p[3]; uint32_t n[3]; };

// Bias correction of adam_step (adam.h:149-150) for a parameter whose private step counter reads `cs`:
// sqrt(1 - beta2^cs) / (1 - beta1^cs).  Two powf per parameter made the kernel issue-bound (83 % issue-active, 40 % of HBM peak
// in the first ncu capture), so the factor is looked up in a table that k_adam_bias_table (api.cu) fills once when the context is created;
// the inline evaluation below is only the fallback for counters beyond the table (> 2^20 steps).
__device__ __forceinline__ float adam_bias_factor(const AdamHyper& h, uint32_t cs) {
	return sqrtf(1 - powf(h.beta2, (float)cs)) / (1 - powf(h.beta1, (float)cs));
}

__device__ __forceinline__ __half adam_one(bool is_matrix, float g_raw, const AdamHyper& h, float& w, float& m1, float& m2, uint32_t& step, __half p, bool& touched) {
	float gradient = h.inv_loss_scale != 0.f ? g_raw * h.inv_loss_scale : g_raw / h.loss_scale;
	touched = false;
	if (is_matrix || gradient != 0) {            // adam.h:109-112: non-matrix params with zero gradient are skipped entirely
		touched = true;
		if (is_matrix) gradient = fmaf(h.l2_reg, w, gradient);   // adam.h:132-135
		const float gsq = gradient * gradient;
		const float fm = fmaf(h.beta1, m1, (1 - h.beta1) * gradient);
		const float sm = fmaf(h.beta2, m2, (1 - h.beta2) * gsq);
		m1 = fm; m2 = sm;
		float learning_rate = h.lr;
		const uint32_t cs = ++step;
		learning_rate *= cs < h.bias_table_size ? h.bias_table[cs] : adam_bias_factor(h, cs);
		const float eff = fminf(fmaxf(learning_rate / (sqrtf(sm) + h.eps), 0.f), 3.402823466e+38f);
		const float decayed = (1 - 0.f * learning_rate) * w - copysignf(0.f * learning_rate, w);
		const float nw = fmaf(-eff, fm, decayed);
		w = nw;
		p = __float2half_rn(nw);
	}
	return p;
}

// 1 / x if x is a (normal) power of two with a normal reciprocal, else 0
static float exact_reciprocal(float x) {
	int e = 0;
	if (!(x > 0.f) || std::frexp(x, &e) != 0.5f || e < -120 || e > 120) return 0.f;
	return 1.0f / x;
}

// 4 consecutive parameters per thread (n and n_matrix are multiples of 4): 16-byte gradient loads; the optimizer state of an
// untouched quad (the common case for the sparse hash-grid levels) is never read.
struct AdamBuffers { float* grads; float* master; __half* params; float* m1; float* m2; uint32_t* steps; __half* ema; };

__device__ __forceinline__ void adam_ema_quad(uint64_t q, uint64_t n_matrix4, const AdamHyper& h, const AdamBuffers& B) {
	const bool is_matrix = q < n_matrix4;
	const float4 g4 = reinterpret_cast<const float4*>(B.grads)[q];
	if (h.zero_grads && (g4.x != 0.f || g4.y != 0.f || g4.z != 0.f || g4.w != 0.f)) reinterpret_cast<float4*>(B.grads)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
	uint2 pw = reinterpret_cast<const uint2*>(B.params)[q];
	__half* ph = reinterpret_cast<__half*>(&pw);
	const float g[4] = {g4.x, g4.y, g4.z, g4.w};
	if ((is_matrix || g4.x != 0.f || g4.y != 0.f || g4.z != 0.f || g4.w != 0.f)) {
		float4 w4 = LD_STATE(reinterpret_cast<const float4*>(B.master) + q);
		float4 a4 = LD_STATE(reinterpret_cast<const float4*>(B.m1) + q);
		float4 b4 = LD_STATE(reinterpret_cast<const float4*>(B.m2) + q);
		uint4 s4 = LD_STATE(reinterpret_cast<const uint4*>(B.steps) + q);
		float* w = reinterpret_cast<float*>(&w4); float* a = reinterpret_cast<float*>(&a4); float* b = reinterpret_cast<float*>(&b4);
		uint32_t* st = reinterpret_cast<uint32_t*>(&s4);
		bool any = false;
		#pragma unroll
		for (int k = 0; k < 4; ++k) {
			bool touched;
			ph[k] = adam_one(is_matrix, g[k], h, w[k], a[k], b[k], st[k], ph[k], touched);
			any |= touched;
		}
		if (any) {
			ST_STATE(reinterpret_cast<float4*>(B.master) + q, w4);
			ST_STATE(reinterpret_cast<float4*>(B.m1) + q, a4);
			ST_STATE(reinterpret_cast<float4*>(B.m2) + q, b4);
			ST_STATE(reinterpret_cast<uint4*>(B.steps) + q, s4);
			reinterpret_cast<uint2*>(B.params)[q] = pw;
		}
	}
	uint2 ew = LD_STATE(reinterpret_cast<const uint2*>(B.ema) + q);
	__half* eh = reinterpret_cast<__half*>(&ew);
	#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const float filtered = fmaf(__half2float(eh[k]) * h.ema_decay, h.ema_debias_old, __half2float(ph[k]) * (1 - h.ema_decay)) * h.ema_debias_new;
		eh[k] = __float2half_rn(filtered);
	}
	ST_STATE(reinterpret_cast<uint2*>(B.ema) + q, ew);
}

__global__ void __launch_bounds__(256)
k_adam_ema(uint64_t n4, uint64_t n_matrix4, AdamHyper h, AdamBuffers B, AdamZero z) {
	if (h.zero_grads && blockIdx.x == 0) {
		#pragma unroll
		for (int a = 0; a < 3; ++a) for (uint32_t i = threadIdx.x; i < z.n[a]; i += blockDim.x) z.p[a][i] = 0u;
	}
	const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q < n4) adam_ema_quad(q, n_matrix4, h, B);
}

int launch_adam_ema(cudaStream_t stream, uint64_t n, uint64_t n_matrix, float loss_scale, float lr, float beta1, float beta2, float eps, float l2_reg,
                    float ema_decay, uint32_t ema_step, float* grads, float* master, __half* params, float* m1, float* m2, uint32_t* steps, __half* ema,
                    float* bias_table, uint32_t bias_table_size, int zero_grads, uint32_t* const zero_ptrs[3], const uint32_t zero_counts[3]) {
	const float ema_debias_old = 1 - (float)pow((double)ema_decay, (double)(ema_step - 1));
	const float ema_debias_new = 1.0f / (1 - (float)pow((double)ema_decay, (double)ema_step));
	if (n % 4 != 0 || n_matrix % 4 != 0) { neus_set_error("parameter counts must be multiples of 4", __FILE__, __LINE__); return -1; }
	AdamHyper h = {loss_scale, lr, beta1, beta2, eps, l2_reg, ema_decay, ema_debias_old, ema_debias_new, ema_step, bias_table, bias_table_size, zero_grads, exact_reciprocal(loss_scale)};
	AdamZero z = {};
	if (zero_grads) for (int a = 0; a < 3; ++a) { z.p[a] = zero_ptrs ? zero_ptrs[a] : nullptr; z.n[a] = (zero_ptrs && zero_ptrs[a]) ? zero_counts[a] : 0u; }
	const uint64_t n4 = n / 4;
	const uint32_t blocks = (uint32_t)((n4 + 255) / 256);
	const AdamBuffers bufs = {grads, master, params, m1, m2, steps, ema};
	k_adam_ema<<<blocks, 256, 0, stream>>>(n4, n_matrix / 4, h, bufs, z);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}

// ------------------------------------------------------------------------------------------------------------------------
// Data-parallel optimizer step fused with its collective: instead of all-reduce(grads) followed by a replicated
// Adam, rank r owns every W-th block of 1024 parameters, SUMS the gradients of all ranks for its blocks
// straight out of the peers' gradient buffers (NVLink/NVSwitch P2P loads, fixed rank order), runs Adam + EMA on it, and
// stores the refreshed fp16 parameters and EMA values into every rank's buffers (P2P stores).  Optimizer state (fp32 master,
// moments, step counters) is touched for 1/W of the parameters per GPU; no gradient ever round-trips through a reduced copy.
// Ordering between ranks (backward finished before peers read; stores landed before peers' next forward) is the caller's:
// a stream-ordered tiny collective on either side (neus2_b200/parallel.py).
// ------------------------------------------------------------------------------------------------------------------------
struct PeerPtrs { float* grads[NEUS2_MAX_PEERS]; __half* params[NEUS2_MAX_PEERS]; __half* ema[NEUS2_MAX_PEERS]; };

__device__ __forceinline__ void adam_ema_p2p_quad(uint64_t q, uint64_t n_matrix4, const AdamHyper& h, int world, const PeerPtrs& peers,
                                                  float* __restrict__ master, float* __restrict__ m1, float* __restrict__ m2, uint32_t* __restrict__ steps);

__global__ void __launch_bounds__(256)
k_adam_ema_p2p(uint64_t n4, uint32_t rank, uint64_t n_matrix4, AdamHyper h, int world, PeerPtrs peers,
               float* __restrict__ master, float* __restrict__ m1, float* __restrict__ m2, uint32_t* __restrict__ steps,
               PeerXfer barrier, PeerZero zero, uint32_t* __restrict__ done_counter) {
	// rank r owns the parameter blocks r, r + W, r + 2W, ... (OPTIMIZER_BLOCK_QUADS quads each): this CTA takes this rank's blockIdx-th block
	const uint64_t q = ((uint64_t)blockIdx.x * (uint32_t)world + rank) * OPTIMIZER_BLOCK_QUADS + threadIdx.x;
	if (q < n4) adam_ema_p2p_quad(q, n_matrix4, h, world, peers, master, m1, m2, steps);
	if (barrier.world == 0) return;
	// The barrier "every rank's slice is stored everywhere" rides in this kernel's LAST block (it owns its SM already: a kernel of its own
	// would queue behind the next step's sample generation, which fills every SM the training kernel frees).  Peer stores of all blocks are
	// ordered before the flags: fence.sys + arrival counter on the writers' side, counter + fence.sys + release on the last block's side.
	__shared__ int s_last;
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence_system();
		s_last = atomicInc(done_counter, gridDim.x - 1) == gridDim.x - 1;   // wraps to 0 with the last arrival
	}
	__syncthreads();
	if (!s_last) return;
	if (threadIdx.x < 32) { __threadfence_system(); peer_barrier_warp(barrier, (int)threadIdx.x); }
	__syncthreads();
	#pragma unroll
	for (int a = 0; a < 3; ++a) for (uint32_t i = threadIdx.x; i < zero.n[a]; i += blockDim.x) zero.p[a][i] = 0u;
}

__device__ __forceinline__ void adam_ema_p2p_quad(uint64_t q, uint64_t n_matrix4, const AdamHyper& h, int world, const PeerPtrs& peers,
                                                  float* __restrict__ master, float* __restrict__ m1, float* __restrict__ m2, uint32_t* __restrict__ steps) {
	const bool is_matrix = q < n_matrix4;
	float4 gr[NEUS2_MAX_PEERS];
	#pragma unroll
	for (int r = 0; r < NEUS2_MAX_PEERS; ++r) if (r < world) gr[r] = reinterpret_cast<const float4*>(peers.grads[r])[q];   // all loads in flight together
	float4 g4 = gr[0];
	#pragma unroll
	for (int r = 1; r < NEUS2_MAX_PEERS; ++r) if (r < world) { g4.x += gr[r].x; g4.y += gr[r].y; g4.z += gr[r].z; g4.w += gr[r].w; }
	if (h.zero_grads) {   // nobody else reads this quad of any rank's buffer: clear the consumed ones in place (the next backward pass starts from zeros)
		#pragma unroll
		for (int r = 0; r < NEUS2_MAX_PEERS; ++r)
			if (r < world && (gr[r].x != 0.f || gr[r].y != 0.f || gr[r].z != 0.f || gr[r].w != 0.f)) reinterpret_cast<float4*>(peers.grads[r])[q] = make_float4(0.f, 0.f, 0.f, 0.f);
	}
	uint2 pw = reinterpret_cast<const uint2*>(peers.params[0])[q];     // slot 0 is always this rank's own buffer
	__half* ph = reinterpret_cast<__half*>(&pw);
	const float g[4] = {g4.x, g4.y, g4.z, g4.w};
	if ((is_matrix || g4.x != 0.f || g4.y != 0.f || g4.z != 0.f || g4.w != 0.f)) {
		float4 w4 = LD_STATE(reinterpret_cast<const float4*>(master) + q);
		float4 a4 = LD_STATE(reinterpret_cast<const float4*>(m1) + q);
		float4 b4 = LD_STATE(reinterpret_cast<const float4*>(m2) + q);
		uint4 s4 = LD_STATE(reinterpret_cast<const uint4*>(steps) + q);
		float* w = reinterpret_cast<float*>(&w4); float* a = reinterpret_cast<float*>(&a4); float* b = reinterpret_cast<float*>(&b4);
		uint32_t* st = reinterpret_cast<uint32_t*>(&s4);
		bool any = false;
		#pragma unroll
		for (int k = 0; k < 4; ++k) {
			bool touched;
			ph[k] = adam_one(is_matrix, g[k], h, w[k], a[k], b[k], st[k], ph[k], touched);
			any |= touched;
		}
		if (any) {
			ST_STATE(reinterpret_cast<float4*>(master) + q, w4);
			ST_STATE(reinterpret_cast<float4*>(m1) + q, a4);
			ST_STATE(reinterpret_cast<float4*>(m2) + q, b4);
			ST_STATE(reinterpret_cast<uint4*>(steps) + q, s4);
			#pragma unroll
			for (int r = 0; r < NEUS2_MAX_PEERS; ++r) if (r < world) reinterpret_cast<uint2*>(peers.params[r])[q] = pw;
		}
	}
	uint2 ew = reinterpret_cast<const uint2*>(peers.ema[0])[q];
	__half* eh = reinterpret_cast<__half*>(&ew);
	#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const float filtered = fmaf(__half2float(eh[k]) * h.ema_decay, h.ema_debias_old, __half2float(ph[k]) * (1 - h.ema_decay)) * h.ema_debias_new;
		eh[k] = __float2half_rn(filtered);
	}
	#pragma unroll
	for (int r = 0; r < NEUS2_MAX_PEERS; ++r) if (r < world) reinterpret_cast<uint2*>(peers.ema[r])[q] = ew;
}

int launch_adam_ema_p2p(cudaStream_t stream, uint64_t n, uint64_t n_matrix, uint32_t rank, uint32_t world, float loss_scale, float lr, float beta1, float beta2,
                        float eps, float l2_reg, float ema_decay, uint32_t ema_step, float* const* grads_peers, __half* const* params_peers,
                        __half* const* ema_peers, float* master, float* m1, float* m2, uint32_t* steps, float* bias_table, uint32_t bias_table_size, int zero_grads,
                        const PeerXfer* barrier, const PeerZero* zero, uint32_t* done_counter) {
	const float ema_debias_old = 1 - (float)pow((double)ema_decay, (double)(ema_step - 1));
	const float ema_debias_new = 1.0f / (1 - (float)pow((double)ema_decay, (double)ema_step));
	if (n % 4 != 0 || n_matrix % 4 != 0) { neus_set_error("parameter counts must be multiples of 4", __FILE__, __LINE__); return -1; }
	if (world < 1 || world > NEUS2_MAX_PEERS || rank >= world) { neus_set_error("bad rank / world size for the sharded optimizer", __FILE__, __LINE__); return -1; }
	AdamHyper h = {loss_scale, lr, beta1, beta2, eps, l2_reg, ema_decay, ema_debias_old, ema_debias_new, ema_step, bias_table, bias_table_size, zero_grads, exact_reciprocal(loss_scale)};
	PeerPtrs pp;
	for (uint32_t r = 0; r < NEUS2_MAX_PEERS; ++r) { pp.grads[r] = nullptr; pp.params[r] = nullptr; pp.ema[r] = nullptr; }
	// slot 0 = own buffers, then the peers in ascending rank order: the sum order is the same on every step
	uint32_t slot = 0;
	pp.grads[slot] = grads_peers[rank]; pp.params[slot] = params_peers[rank]; pp.ema[slot] = ema_peers[rank]; ++slot;
	for (uint32_t r = 0; r < world; ++r) if (r != rank) { pp.grads[slot] = grads_peers[r]; pp.params[slot] = params_peers[r]; pp.ema[slot] = ema_peers[r]; ++slot; }
	PeerXfer xb = {}; PeerZero z = {};
	if (barrier) { if (!done_counter) { neus_set_error("the fused barrier needs an arrival counter", __FILE__, __LINE__); return -1; } xb = *barrier; if (zero) z = *zero; }
	const uint64_t n_blocks = (n / 4 + OPTIMIZER_BLOCK_QUADS - 1) / OPTIMIZER_BLOCK_QUADS;
	const uint32_t blocks = (uint32_t)((n_blocks + world - 1) / world);   // the last one may fall past the end on some ranks: its threads do nothing
	static_assert(OPTIMIZER_BLOCK_QUADS == 256, "one CTA of 256 threads per parameter block");
	k_adam_ema_p2p<<<blocks, 256, 0, stream>>>(n / 4, rank, n_matrix / 4, h, (int)world, pp, master, m1, m2, steps, xb, z, done_counter);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}

// Ownership of the sharded optimizer: parameter blocks of OPTIMIZER_BLOCK_QUADS quads (1024 parameters) are dealt round-robin, block b
// belongs to rank b % world.  (Contiguous slices gave the rank holding the finest hash-grid levels -- touched by every sample -- up to 25 %
// more Adam work than the rank holding the sparsely touched coarse hashed levels, and the barrier behind the optimizer waits for the slowest.)
uint32_t optimizer_block_owner(uint64_t param_index, uint32_t world) { return (uint32_t)((param_index / (4ull * OPTIMIZER_BLOCK_QUADS)) % world); }
