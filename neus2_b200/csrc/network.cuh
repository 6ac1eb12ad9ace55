// network.cuh -- warp-level building blocks of the fused NeuS network kernels.
//
// One warp owns a tile of 16 samples end to end: hash-grid gather -> SDF MLP -> analytic normal ->
// SH -> colour MLP (-> full backward incl. the second-order terms in the training kernel).  Activations
// travel between layers in mma.sync fragments (registers); only the per-sample scalar stages go through
// the warp's private shared-memory slab.  Warps never synchronise with each other after the weights are
// staged, so gather latency of one warp overlaps tensor work of the others.
//
// Restates (does not copy) NerfNetwork::forward_impl / backward_impl (include/neural-graphics-primitives/
// nerf_network.h:145-654), kernel_grid* (tcnn encodings/grid.h:174-1207) and the FullyFusedMLP arithmetic
// (tcnn src/fully_fused_mlp.cu).  Rounding points are those of oracle/neus_oracle.cpp.
#pragma once
#include "common.cuh"

#define TILE 16                // samples per warp tile (one mma m16 row block)
#define LD_X0 40               // 32 + 8 pad (halves)   -> conflict-free fragment loads
#define LD_H 72                // 64 + 8 pad
#define LD_C0 56               // 48 + 8 pad
#define LD_DY 85               // 84 floats + 1 pad per sample
#define SDF_IN 32
#define RGB_IN 48

// ---- shared-memory weight block (halves) -------------------------------------------------------
// One copy per matrix, row = output unit, contiguous = input unit.  B fragments come from ldmatrix: plain for the forward
// orientation (B(k,n) = W[n][k]), .trans for the transposed products of the backward pass (B(k,n) = W[k][n]) -- no transposed copies.
// Row strides are odd multiples of 16 bytes: the 8 row addresses of an 8x8 ldmatrix tile fall into 8 different bank groups.
struct WeightSmem {
	static constexpr int W0 = 0;                       // [64][LD_X0]
	static constexpr int W1 = W0 + 64 * LD_X0;         // [16][LD_H]
	static constexpr int C0 = W1 + 16 * LD_H;          // [64][LD_C0]
	static constexpr int C1 = C0 + 64 * LD_C0;         // [64][LD_H]
	static constexpr int C2 = C1 + 64 * LD_H;          // [16][LD_H]
	static constexpr int FWD_END = C2 + 16 * LD_H;
	static constexpr int BWD_END = FWD_END;
};

// per-warp slab (bytes); every offset is a multiple of 16
struct WarpSlab {
	static constexpr int COORD = 0;                          // float [16][7]
	static constexpr int X0 = COORD + 16 * 7 * 4;            // half  [16][LD_X0]   sdf-MLP input
	static constexpr int DIN = X0 + 16 * LD_X0 * 2;          // half  [16][LD_X0]   d sdf / d sdf-input  (gE in cols 3..30)
	static constexpr int CIN = DIN + 16 * LD_X0 * 2;         // half  [16][LD_X0]   colour input cols 16..47
	static constexpr int SDF = CIN + 16 * LD_X0 * 2;         // half  [16]          raw sdf-MLP output 0
	static constexpr int NRM = SDF + 16 * 2;                 // float [16][3]       normal
	static constexpr int DY = NRM + 16 * 3 * 4;              // float [16][LD_DY]
	static constexpr int OUT = DY + 16 * LD_DY * 4;          // half  [16][16]      output staging (inference kernels only)
	static constexpr int FWD_END = OUT + 16 * 16 * 2;
	// training extras
	static constexpr int DL = OUT;                           // half  [16][16]      dL/dout (after rollover scaling); the training kernel produces no outputs
	static constexpr int DX = FWD_END;                       // half  [16][LD_X0]   dL/d sdf-input
	static constexpr int DI = DX + 16 * LD_X0 * 2;           // half  [16][LD_X0]   d_density_input (2nd order)
	static constexpr int V = DI + 16 * LD_X0 * 2;            // float [16][3]       v = dL/d normal
	static constexpr int VN = V + 16 * 3 * 4;                // half  [16][8]       colour-input grad cols 32..39
	static constexpr int BWD_END = VN + 16 * 8 * 2;
	// DeltaNetwork training extras (k_network_fwdbwd<true> only)
	static constexpr int DORIG = BWD_END;                    // float [16][3]       un-deformed positions of the tile
	static constexpr int DACC = DORIG + 16 * 3 * 4;          // float [16][9]       per-lane sums of the 9 delta gradients
	static constexpr int DELTA_END = DACC + 16 * 9 * 4;
	// operand staging of the in-kernel weight gradients (umma.cuh): A tile [16][64] + B tile up to [16][48], canonical layout
	static constexpr int STAGE_A = 0;                        // relative to the warp's staging block
	static constexpr int STAGE_B = 16 * 64 * 2;
	static constexpr int STAGE_BYTES = STAGE_B + 16 * 48 * 2;
};
static_assert(WarpSlab::OUT % 16 == 0 && WarpSlab::DX % 16 == 0 && WarpSlab::DI % 16 == 0 && WarpSlab::BWD_END % 16 == 0 && WarpSlab::DELTA_END % 16 == 0, "slab alignment");

#ifndef NEUS2_STAGE_WEIGHTS_TMA
#define NEUS2_STAGE_WEIGHTS_TMA 0   // measured (same box, profiles/r02_bench_n1_tma_ab.md): the bulk-copy variant costs 8 us in the first pass (224 small copies per CTA)
#endif
__device__ __forceinline__ void stage_matrix(__half* dst, int ld_dst, const __half* __restrict__ src, int rows, int cols, int tid, int nthreads) {
	for (int i = tid; i < rows * cols; i += nthreads) {
		int r = i / cols, c = i - r * cols;
		dst[r * ld_dst + c] = src[i];
	}
}

// All five weight matrices: global -> shared memory with the bulk-copy engine (TMA, cp.async.bulk), one copy per (padded) row, issued by
// the first warp and completed on an mbarrier that every thread of the CTA then waits for.  224 row copies of 64..128 bytes replace 11,264
// scalar load/store pairs per CTA; rows and their destinations are 16-byte aligned by construction (row strides 80 / 112 / 144 bytes).
// Must be called by ALL threads of the CTA, once, before the first use of `sw`; contains the block-wide synchronisation.
__device__ __forceinline__ void stage_weights(__half* sw, const __half* __restrict__ params, const ParamLayout& pl, bool train, int tid, int nthreads) {
	(void)train;
#if !NEUS2_STAGE_WEIGHTS_TMA   // A/B switch: the thread-copy variant this replaced
	stage_matrix(sw + WeightSmem::W0, LD_X0, params + pl.off_w0, 64, SDF_IN, tid, nthreads);
	stage_matrix(sw + WeightSmem::W1, LD_H, params + pl.off_w1, 16, 64, tid, nthreads);
	stage_matrix(sw + WeightSmem::C0, LD_C0, params + pl.off_c0, 64, RGB_IN, tid, nthreads);
	stage_matrix(sw + WeightSmem::C1, LD_H, params + pl.off_c1, 64, 64, tid, nthreads);
	stage_matrix(sw + WeightSmem::C2, LD_H, params + pl.off_c2, 16, 64, tid, nthreads);
	return;   // every caller synchronises the block behind this call
#endif
	(void)nthreads;
	__shared__ __align__(8) uint64_t bar;
	const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_addr));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (tid < 32) {
		constexpr uint32_t total_bytes = (64 * SDF_IN + 16 * 64 + 64 * RGB_IN + 64 * 64 + 16 * 64) * 2;
		if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_addr), "r"(total_bytes) : "memory");
		__syncwarp();
		// (matrix, row) pairs in one list: W0 64 rows, W1 16, C0 64, C1 64, C2 16 = 224 rows, 7 per lane
		for (int i = tid; i < 224; i += 32) {
			int r = i, dst_off, ld, cols; uint32_t src_off;
			if (r < 64) { dst_off = WeightSmem::W0; ld = LD_X0; cols = SDF_IN; src_off = pl.off_w0; }
			else if ((r -= 64) < 16) { dst_off = WeightSmem::W1; ld = LD_H; cols = 64; src_off = pl.off_w1; }
			else if ((r -= 16) < 64) { dst_off = WeightSmem::C0; ld = LD_C0; cols = RGB_IN; src_off = pl.off_c0; }
			else if ((r -= 64) < 64) { dst_off = WeightSmem::C1; ld = LD_H; cols = 64; src_off = pl.off_c1; }
			else { r -= 64; dst_off = WeightSmem::C2; ld = LD_H; cols = 64; src_off = pl.off_c2; }
			const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sw + dst_off + r * ld);
			const __half* src = params + src_off + (size_t)r * cols;
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			             :: "r"(dst), "l"(src), "r"((uint32_t)(cols * 2)), "r"(bar_addr) : "memory");
		}
	}
	uint32_t done = 0;
	while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar_addr) : "memory");
}

// ---- fragment helpers ------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const __half* p) {
	const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
	asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const __half* p) {
	const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
	asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], const __half* p) {
	const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
	asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], const __half* p) {
	const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
	asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}

// c[NT][4] = A(regs a[KT][4]) x B.  W is a row-major [out][in] weight matrix, row stride ld (halves).
//   TRANS = false: B(k,n) = W[n][k]   (forward: k = input unit, n = output unit)
//   TRANS = true:  B(k,n) = W[k][n]   (backward: k = output unit, n = input unit)
// One ldmatrix.x4 feeds two n-tiles of one k-tile (4 x fewer shared-memory instructions than 32-bit fragment loads).  Each c[nt]
// accumulates its k-tiles in ascending order.
template <int KT, int NT, bool TRANS = false>
__device__ __forceinline__ void mma_layer(const uint32_t (&a)[KT][4], const __half* W, int ld, float (&c)[NT][4], int g, int t) {
	const int lane = g * 4 + t, q = lane >> 3, r = lane & 7;
	#pragma unroll
	for (int nt = 0; nt < NT; ++nt) c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
	#pragma unroll
	for (int p = 0; p < NT / 2; ++p) {
		// tiles of the x4 load: (n-tile 2p, k lo) (n-tile 2p, k hi) (n-tile 2p+1, k lo) (n-tile 2p+1, k hi)
		const __half* base = TRANS ? W + ((q & 1) * 8 + r) * ld + 8 * (2 * p + (q >> 1))
		                           : W + (8 * (2 * p + (q >> 1)) + r) * ld + (q & 1) * 8;
		#pragma unroll
		for (int kt = 0; kt < KT; ++kt) {
			uint32_t b[4];
			if (TRANS) ldsm_x4_t(b, base + 16 * kt * ld); else ldsm_x4(b, base + 16 * kt);
			mma16816(c[2 * p], a[kt], b[0], b[1]);
			mma16816(c[2 * p + 1], a[kt], b[2], b[3]);
		}
	}
	if (NT & 1) {   // last n-tile alone: x2 (lanes 0..15 supply the row addresses)
		constexpr int nt = NT - 1;
		const int q2 = q & 1;
		const __half* base = TRANS ? W + (q2 * 8 + r) * ld + 8 * nt : W + (8 * nt + r) * ld + q2 * 8;
		#pragma unroll
		for (int kt = 0; kt < KT; ++kt) {
			uint32_t b[2];
			if (TRANS) ldsm_x2_t(b, base + 16 * kt * ld); else ldsm_x2(b, base + 16 * kt);
			mma16816(c[nt], a[kt], b[0], b[1]);
		}
	}
}

// load an A fragment (16 rows x 16 cols starting at col0) from a row-major half tile in shared memory
__device__ __forceinline__ void load_a_frag(uint32_t (&a)[4], const __half* S, int ld, int col0, int g, int t) {
	a[0] = *reinterpret_cast<const uint32_t*>(S + g * ld + col0 + 2 * t);
	a[1] = *reinterpret_cast<const uint32_t*>(S + (g + 8) * ld + col0 + 2 * t);
	a[2] = *reinterpret_cast<const uint32_t*>(S + g * ld + col0 + 8 + 2 * t);
	a[3] = *reinterpret_cast<const uint32_t*>(S + (g + 8) * ld + col0 + 8 + 2 * t);
}

template <int NT, bool RELU>
__device__ __forceinline__ void pack_frags(const float (&c)[NT][4], uint32_t (&a)[NT / 2][4]) {
	#pragma unroll
	for (int kt = 0; kt < NT / 2; ++kt) {
		#pragma unroll
		for (int r = 0; r < 4; ++r) {
			const int nt = 2 * kt + (r >> 1);
			float v0 = c[nt][(r & 1) * 2], v1 = c[nt][(r & 1) * 2 + 1];
			if (RELU) { v0 = v0 > 0.f ? v0 : 0.f; v1 = v1 > 0.f ? v1 : 0.f; }
			a[kt][r] = pack_h2(v0, v1);
		}
	}
}

// per-half mask: 0xffff where the half in `ref` is > 0 (hidden activation passed the ReLU)
__device__ __forceinline__ uint32_t relu_mask2(uint32_t ref) {
	return __hgt2_mask(*reinterpret_cast<const __half2*>(&ref), __half2half2(__ushort_as_half((unsigned short)0)));   // one HSET2
}

template <int NT>
__device__ __forceinline__ void pack_frags_masked(const float (&c)[NT][4], const uint32_t (&ref)[NT / 2][4], uint32_t (&a)[NT / 2][4]) {
	#pragma unroll
	for (int kt = 0; kt < NT / 2; ++kt) {
		#pragma unroll
		for (int r = 0; r < 4; ++r) {
			const int nt = 2 * kt + (r >> 1);
			a[kt][r] = pack_h2(c[nt][(r & 1) * 2], c[nt][(r & 1) * 2 + 1]) & relu_mask2(ref[kt][r]);
		}
	}
}

// write C tiles (fp32 -> half) to a row-major shared tile: cols 8*nt + 2t
template <int NT>
__device__ __forceinline__ void store_c_smem(const float (&c)[NT][4], __half* S, int ld, int g, int t) {
	#pragma unroll
	for (int nt = 0; nt < NT; ++nt) {
		*reinterpret_cast<uint32_t*>(S + g * ld + 8 * nt + 2 * t) = pack_h2(c[nt][0], c[nt][1]);
		*reinterpret_cast<uint32_t*>(S + (g + 8) * ld + 8 * nt + 2 * t) = pack_h2(c[nt][2], c[nt][3]);
	}
}

// store A-layout fragments (KT k-tiles = 16*KT columns) to a row-major global matrix [rows][ld] of halves
template <int KT>
__device__ __forceinline__ void store_a_global(const uint32_t (&a)[KT][4], __half* G, int ld, int row0, int n_valid, int g, int t) {
	#pragma unroll
	for (int kt = 0; kt < KT; ++kt) {
		if (g < n_valid) {
			*reinterpret_cast<uint32_t*>(G + (size_t)(row0 + g) * ld + 16 * kt + 2 * t) = a[kt][0];
			*reinterpret_cast<uint32_t*>(G + (size_t)(row0 + g) * ld + 16 * kt + 8 + 2 * t) = a[kt][2];
		}
		if (g + 8 < n_valid) {
			*reinterpret_cast<uint32_t*>(G + (size_t)(row0 + g + 8) * ld + 16 * kt + 2 * t) = a[kt][1];
			*reinterpret_cast<uint32_t*>(G + (size_t)(row0 + g + 8) * ld + 16 * kt + 8 + 2 * t) = a[kt][3];
		}
	}
}

// ---- hash grid ----------------------------------------------------------------------------------
// grid.h:119-153 with the per-level "hashed" decision hoisted to the host (GridMeta.hashed semantics:
// hashed <=> res^3 > hashmap_size).
// general wrap of a dense index (grid.h:146: index % hashmap_size) -- out of line: for positions inside [0,1] the dense index is
// below 2 * hashmap_size and the caller's single conditional subtract is the whole story
static __device__ __noinline__ uint32_t grid_index_wrap(uint32_t index, uint32_t hashmap_size) { return index % hashmap_size; }

__device__ __forceinline__ uint32_t grid_index(bool hashed, uint32_t hashmap_size, uint32_t res, uint32_t x, uint32_t y, uint32_t z) {
	uint32_t index;
	if (hashed) {
		index = x ^ (y * 2654435761u) ^ (z * 805459861u);
		return (index & (hashmap_size - 1u)) * 2u;   // hashed levels have power-of-two size (2^log2_hashmap_size)
	}
	index = x + y * res + z * res * res;
	if (index >= hashmap_size) {
		index -= hashmap_size;
		if (index >= hashmap_size) index = grid_index_wrap(index, hashmap_size);
	}
	return index * 2u;
}

struct LevelGeom {
	float w[3];          // fractional position
	uint32_t idx[8];     // feature-pair index (in halves) of the 8 corners, corner bit d set => +1 in dim d
	uint32_t cell[3];    // integer cell coordinates (pos_grid)
};

__device__ __forceinline__ void level_geom(const GridMeta& gm, int l, float x, float y, float z, LevelGeom& lg) {
	const float scale = gm.scale[l];
	const uint32_t res = gm.res[l];
	const uint32_t hs = gm.offset[l + 1] - gm.offset[l];
	const bool hashed = (gm.hashed_mask >> l) & 1u;
	uint32_t p[3];
	const float in[3] = {x, y, z};
	#pragma unroll
	for (int d = 0; d < 3; ++d) {
		float pf = fmaf(in[d], scale, 0.5f);      // pos_fract, tcnn common_device.h:403-434
		float fl = floorf(pf);
		p[d] = (uint32_t)(int)fl;
		lg.w[d] = pf - fl;
		lg.cell[d] = p[d];
	}
	#pragma unroll
	for (int c = 0; c < 8; ++c) {
		lg.idx[c] = grid_index(hashed, hs, res, p[0] + (c & 1), p[1] + ((c >> 1) & 1), p[2] + ((c >> 2) & 1));
	}
}

// Gather for one 16-sample tile.  lane -> (sample lane&15, level parity lane>>4).
// Writes: X0[s][3+2l..] = encoding (half, accumulated in half in corner order, grid.h:316-318),
//         DY[s][ (2l+f)*3 + d ] = d enc / d x (fp32, grid.h:329-367)  when WANT_DY.
// The lane's levels are processed GATHER_BATCH at a time: all 8*GATHER_BATCH corner loads are issued before the first one
// is consumed, so a warp keeps 4x more L2 requests in flight than a level-by-level loop (the gather is latency bound).
#ifndef GATHER_BATCH_FWD
#define GATHER_BATCH_FWD 2      // inference kernels: registers are free during the gather
#endif
#ifndef GATHER_BATCH_SDF
#define GATHER_BATCH_SDF 4      // SDF-only kernels (occupancy refresh, lattice): nothing else is live, 4 levels in flight measured fastest
#endif
#ifndef GATHER_BATCH_TRAIN
#define GATHER_BATCH_TRAIN 2    // training kernel: per-warp accumulators are live across the gather
#endif
template <bool WANT_DY, int GATHER_BATCH>
__device__ __forceinline__ void gather_tile(const GridMeta& gm, const __half* __restrict__ grid, const float* sCoord, __half* sX0, float* sDY, int lane) {
	const int s = lane & 15;
	const float x = sCoord[s * 7 + 0], y = sCoord[s * 7 + 1], z = sCoord[s * 7 + 2];
	const int n_levels = (int)gm.n_levels;
	for (int l0 = lane >> 4; l0 < n_levels; l0 += 2 * GATHER_BATCH) {
		uint32_t v[GATHER_BATCH][8];
		float w[GATHER_BATCH][3];
		#pragma unroll
		for (int u = 0; u < GATHER_BATCH; ++u) {
			const int l = l0 + 2 * u;
			if (l < n_levels && (uint32_t)l <= gm.valid_level) {
				LevelGeom lg;
				level_geom(gm, l, x, y, z, lg);
				w[u][0] = lg.w[0]; w[u][1] = lg.w[1]; w[u][2] = lg.w[2];
				const __half* g = grid + (size_t)gm.offset[l] * 2;
				#pragma unroll
				for (int c = 0; c < 8; ++c) v[u][c] = ld_nc_u32(g + lg.idx[c]);
			}
		}
		#pragma unroll
		for (int u = 0; u < GATHER_BATCH; ++u) {
			const int l = l0 + 2 * u;
			if (l >= n_levels) break;
			__half* xo = sX0 + s * LD_X0 + 3 + 2 * l;
			float* dyo = sDY + s * LD_DY + 6 * l;
			if ((uint32_t)l > gm.valid_level) {
				xo[0] = __float2half(0.f); xo[1] = __float2half(0.f);
				if (WANT_DY) { for (int k = 0; k < 6; ++k) dyo[k] = 0.f; }
				continue;
			}
			const float wx[2] = {1.f - w[u][0], w[u][0]}, wy[2] = {1.f - w[u][1], w[u][1]}, wz[2] = {1.f - w[u][2], w[u][2]};
			// both features at once: one packed convert and one HADD2 per corner (each half of the pair rounds exactly as the scalar form)
			__half2 r = __float2half2_rn(0.f);
			#pragma unroll
			for (int c = 0; c < 8; ++c) {
				const float weight = (wx[c & 1] * wy[(c >> 1) & 1]) * wz[(c >> 2) & 1];
				r = __hadd2(r, __floats2half2_rn(weight * h_lo(v[u][c]), weight * h_hi(v[u][c])));
			}
			xo[0] = __low2half(r); xo[1] = __high2half(r);
			if (WANT_DY) {
				const float scale = gm.scale[l];
				float g0[3] = {0.f, 0.f, 0.f}, g1[3] = {0.f, 0.f, 0.f};
				#pragma unroll
				for (int i = 0; i < 4; ++i) {
					const int b0 = i & 1, b1 = (i >> 1) & 1;
					{   // d/dx : other dims (y, z)
						const float weight = (scale * wy[b0]) * wz[b1];
						const int cl = (b0 << 1) | (b1 << 2), cr = cl | 1;
						g0[0] = fmaf(weight, h_lo(v[u][cr]) - h_lo(v[u][cl]), g0[0]);
						g1[0] = fmaf(weight, h_hi(v[u][cr]) - h_hi(v[u][cl]), g1[0]);
					}
					{   // d/dy : other dims (x, z)
						const float weight = (scale * wx[b0]) * wz[b1];
						const int cl = b0 | (b1 << 2), cr = cl | 2;
						g0[1] = fmaf(weight, h_lo(v[u][cr]) - h_lo(v[u][cl]), g0[1]);
						g1[1] = fmaf(weight, h_hi(v[u][cr]) - h_hi(v[u][cl]), g1[1]);
					}
					{   // d/dz : other dims (x, y)
						const float weight = (scale * wx[b0]) * wy[b1];
						const int cl = b0 | (b1 << 1), cr = cl | 4;
						g0[2] = fmaf(weight, h_lo(v[u][cr]) - h_lo(v[u][cl]), g0[2]);
						g1[2] = fmaf(weight, h_hi(v[u][cr]) - h_hi(v[u][cl]), g1[2]);
					}
				}
				dyo[0] = g0[0]; dyo[1] = g0[1]; dyo[2] = g0[2];
				dyo[3] = g1[0]; dyo[4] = g1[1]; dyo[5] = g1[2];
			}
		}
	}
}

// DeltaNetwork forward (add_global_movement_with_rotation_6d, common_operation.cuh:416-470) in place on the warp's 16
// coordinates: pos' = R (pos + t - off) + off on lanes 0..15; the warped direction is un-warped, rotated and warped again on
// lanes 16..31.  Same expression order as oracle/neus_oracle.cpp:delta_apply.
__device__ __forceinline__ void delta_tile(const float* __restrict__ dx, float* sCoord, int lane, bool with_dir) {
	const int s = lane & 15;
	float R[9];
	#pragma unroll
	for (int k = 0; k < 9; ++k) R[k] = __ldg(dx + DELTA_R + k);
	if (lane < 16) {
		const float q0 = __fadd_rn(sCoord[s * 7 + 0], __ldg(dx + DELTA_TMO + 0)), q1 = __fadd_rn(sCoord[s * 7 + 1], __ldg(dx + DELTA_TMO + 1)),
		            q2 = __fadd_rn(sCoord[s * 7 + 2], __ldg(dx + DELTA_TMO + 2));
		#pragma unroll
		for (int k = 0; k < 3; ++k) {
			const float r = fmaf(R[k * 3 + 2], q2, fmaf(R[k * 3 + 1], q1, __fmul_rn(R[k * 3], q0)));
			sCoord[s * 7 + k] = __fadd_rn(r, __ldg(dx + DELTA_OFF + k));
		}
	} else if (with_dir) {
		const float u0 = fmaf(sCoord[s * 7 + 4], 2.0f, -1.0f), u1 = fmaf(sCoord[s * 7 + 5], 2.0f, -1.0f), u2 = fmaf(sCoord[s * 7 + 6], 2.0f, -1.0f);
		#pragma unroll
		for (int k = 0; k < 3; ++k) {
			const float r = fmaf(R[k * 3 + 2], u2, fmaf(R[k * 3 + 1], u1, __fmul_rn(R[k * 3], u0)));
			sCoord[s * 7 + 4 + k] = __fmul_rn(__fadd_rn(r, 1.0f), 0.5f);
		}
	}
	__syncwarp();
}

// gradients_for_normalize (common_operation.cuh:62-90)
__device__ __forceinline__ float dot3f(const float a[3], const float b[3]) { return fmaf(a[2], b[2], fmaf(a[1], b[1], __fmul_rn(a[0], b[0]))); }
__device__ __forceinline__ void grad_normalize(const float in[3], const float go[3], float gi[3]) {
	const float nrm = sqrtf(dot3f(in, in));
	const float n3 = __fmul_rn(__fmul_rn(nrm, nrm), nrm);
	const float xx = __fdiv_rn(fmaf(in[2], in[2], __fmul_rn(in[1], in[1])), n3), yy = __fdiv_rn(fmaf(in[2], in[2], __fmul_rn(in[0], in[0])), n3),
	            zz = __fdiv_rn(fmaf(in[1], in[1], __fmul_rn(in[0], in[0])), n3);
	const float xy = __fdiv_rn(__fmul_rn(-in[0], in[1]), n3), xz = __fdiv_rn(__fmul_rn(-in[0], in[2]), n3), yz = __fdiv_rn(__fmul_rn(-in[1], in[2]), n3);
	gi[0] = fmaf(go[2], xz, fmaf(go[1], xy, __fmul_rn(go[0], xx)));
	gi[1] = fmaf(go[2], yz, fmaf(go[1], yy, __fmul_rn(go[0], xy)));
	gi[2] = fmaf(go[2], zz, fmaf(go[1], yz, __fmul_rn(go[0], xz)));
}
// One sample of DeltaNetwork::backward_impl (add_loss_to_rotation_6d_each<half, float> + gradient_rotation_matrix_to_6d<half>,
// common_operation.cuh:92-150, 788-840): out[0..2] = d/d transition, out[3..8] = d/d rotation-6d, each rounded to half as the
// reference's per-sample matrices are.  L = dL/d(deformed position), xin = un-deformed position.  Mirrors
// oracle/neus_oracle.cpp:delta_backward_sample expression by expression.
__device__ __forceinline__ void delta_backward_sample(const float* __restrict__ dx, const float L[3], const float xin[3], float out[9]) {
	float R[9], a1[3], a2[3], u2[3], b1[3], b2[3], iv[3];
	#pragma unroll
	for (int k = 0; k < 9; ++k) R[k] = __ldg(dx + DELTA_R + k);
	#pragma unroll
	for (int k = 0; k < 3; ++k) {
		a1[k] = __ldg(dx + DELTA_A1 + k); a2[k] = __ldg(dx + DELTA_A2 + k); u2[k] = __ldg(dx + DELTA_U2 + k);
		b1[k] = R[k * 3 + 0]; b2[k] = R[k * 3 + 1];
		iv[k] = __fadd_rn(xin[k], __ldg(dx + DELTA_T + k));
	}
	#pragma unroll
	for (int j = 0; j < 3; ++j) out[j] = round_h(fmaf(R[6 + j], L[2], fmaf(R[3 + j], L[1], __fmul_rn(R[j], L[0]))));
	float G[3][3];
	#pragma unroll
	for (int a = 0; a < 3; ++a) {
		#pragma unroll
		for (int b = 0; b < 3; ++b) G[a][b] = __fmul_rn(L[a], iv[b]);
	}
	float db1[3], db2[3];
	#pragma unroll
	for (int k = 0; k < 3; ++k) { db1[k] = round_h(G[k][0]); db2[k] = round_h(G[k][1]); }
	db1[0] = round_h(__fadd_rn(db1[0], round_h(fmaf(b2[1], G[2][2], -__fmul_rn(b2[2], G[1][2])))));
	db1[1] = round_h(__fadd_rn(db1[1], round_h(fmaf(b2[2], G[0][2], -__fmul_rn(b2[0], G[2][2])))));
	db1[2] = round_h(__fadd_rn(db1[2], round_h(fmaf(b2[0], G[1][2], -__fmul_rn(b2[1], G[0][2])))));
	db2[0] = round_h(__fadd_rn(db2[0], round_h(fmaf(b1[2], G[1][2], -__fmul_rn(b1[1], G[2][2])))));
	db2[1] = round_h(__fadd_rn(db2[1], round_h(fmaf(b1[0], G[2][2], -__fmul_rn(b1[2], G[0][2])))));
	db2[2] = round_h(__fadd_rn(db2[2], round_h(fmaf(b1[1], G[0][2], -__fmul_rn(b1[0], G[1][2])))));
	float r[3];
	grad_normalize(u2, db2, r);
	const float rb = dot3f(r, b1), ba = dot3f(b1, a2);
	float da2[3], db1f[3], da1[3];
	#pragma unroll
	for (int k = 0; k < 3; ++k) {
		da2[k] = __fsub_rn(r[k], __fmul_rn(b1[k], rb));
		db1f[k] = __fsub_rn(__fsub_rn(db1[k], __fmul_rn(r[k], ba)), __fmul_rn(a2[k], rb));
	}
	grad_normalize(a1, db1f, da1);
	#pragma unroll
	for (int k = 0; k < 3; ++k) { out[3 + k] = round_h(da1[k]); out[6 + k] = round_h(da2[k]); }
}

// SH degree 4 on the warped direction (tcnn encodings/spherical_harmonics.h:47-100); 16 halves.
__device__ __forceinline__ void sh4(float dx, float dy, float dz, __half* out) {
	const float x = fmaf(dx, 2.f, -1.f), y = fmaf(dy, 2.f, -1.f), z = fmaf(dz, 2.f, -1.f);
	const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
	out[0] = __float2half_rn(0.28209479177387814f);
	out[1] = __float2half_rn(-0.48860251190291987f * y);
	out[2] = __float2half_rn(0.48860251190291987f * z);
	out[3] = __float2half_rn(-0.48860251190291987f * x);
	out[4] = __float2half_rn(1.0925484305920792f * xy);
	out[5] = __float2half_rn(-1.0925484305920792f * yz);
	out[6] = __float2half_rn(fmaf(0.94617469575755997f, z2, -0.31539156525251999f));
	out[7] = __float2half_rn(-1.0925484305920792f * xz);
	out[8] = __float2half_rn(fmaf(0.54627421529603959f, x2, -(0.54627421529603959f * y2)));
	out[9] = __float2half_rn(__fmul_rn(__fmul_rn(0.59004358992664352f, y), fmaf(-3.0f, x2, y2)));
	out[10] = __float2half_rn(__fmul_rn(__fmul_rn(2.8906114426405538f, xy), z));
	out[11] = __float2half_rn(__fmul_rn(__fmul_rn(0.45704579946446572f, y), fmaf(-5.0f, z2, 1.0f)));
	out[12] = __float2half_rn(__fmul_rn(__fmul_rn(0.3731763325901154f, z), fmaf(5.0f, z2, -3.0f)));
	out[13] = __float2half_rn(__fmul_rn(__fmul_rn(0.45704579946446572f, x), fmaf(-5.0f, z2, 1.0f)));
	out[14] = __float2half_rn(__fmul_rn(__fmul_rn(1.4453057213202769f, z), __fsub_rn(x2, y2)));
	out[15] = __float2half_rn(__fmul_rn(__fmul_rn(0.59004358992664352f, x), fmaf(3.0f, y2, -x2)));
}

// Everything the forward leaves in registers for the backward.
struct FwdState {
	uint32_t h0a[4][4];    // SDF hidden (post-ReLU, half) as A fragments
	uint32_t soa[4];       // SDF-MLP output (16 wide, half) as one A fragment
	uint32_t h1a[4][4];    // colour hidden 1
	uint32_t h2a[4][4];    // colour hidden 2
};

// Forward for one tile.  On return: sOUT holds the assembled 16x16 half output (nerf_network.h:287-325),
// sDIN the d sdf/d sdf-input rows, sNRM the normals, sCIN colour-input cols 16..47.
template <bool KEEP>
__device__ __forceinline__ void forward_tile(const GridMeta& gm, const __half* __restrict__ params, const ParamLayout& pl,
                                             const __half* sw, unsigned char* slab, float sdf_bias, int lane, FwdState& st) {
	const int g = lane >> 2, t = lane & 3;
	float* sCoord = reinterpret_cast<float*>(slab + WarpSlab::COORD);
	__half* sX0 = reinterpret_cast<__half*>(slab + WarpSlab::X0);
	__half* sDIN = reinterpret_cast<__half*>(slab + WarpSlab::DIN);
	__half* sCIN = reinterpret_cast<__half*>(slab + WarpSlab::CIN);
	__half* sOUT = reinterpret_cast<__half*>(slab + WarpSlab::OUT);
	__half* sSDF = reinterpret_cast<__half*>(slab + WarpSlab::SDF);
	float* sNRM = reinterpret_cast<float*>(slab + WarpSlab::NRM);
	float* sDY = reinterpret_cast<float*>(slab + WarpSlab::DY);

	if (gm.delta_xform) delta_tile(gm.delta_xform, sCoord, lane, true);   // m_use_delta (nerf_network.h:162-181)
	// --- SDF-MLP input: [half(x)-half(0.5) | enc | 0]  (nerf_network.h:206-212)
	if (lane < 16) {
		#pragma unroll
		for (int d = 0; d < 3; ++d) sX0[lane * LD_X0 + d] = __hsub(__float2half_rn(sCoord[lane * 7 + d]), __float2half_rn(0.5f));
		for (int k = 3 + 2 * (int)gm.n_levels; k < SDF_IN; ++k) sX0[lane * LD_X0 + k] = __float2half(0.f);
	}
	gather_tile<true, KEEP ? GATHER_BATCH_TRAIN : GATHER_BATCH_FWD>(gm, params + pl.off_grid, sCoord, sX0, sDY, lane);
	__syncwarp();

	// --- SDF MLP (32 -> 64 -> 16), fp32 accumulate, half activations
	uint32_t h0a[4][4];
	{
		uint32_t xa[2][4];
		load_a_frag(xa[0], sX0, LD_X0, 0, g, t);
		load_a_frag(xa[1], sX0, LD_X0, 16, g, t);
		float c[8][4];
		mma_layer<2, 8>(xa, sw + WeightSmem::W0, LD_X0, c, g, t);
		pack_frags<8, true>(c, h0a);
	}
	uint32_t soa[4];
	{
		float c[2][4];
		mma_layer<4, 2>(h0a, sw + WeightSmem::W1, LD_H, c, g, t);
		uint32_t tmp[1][4];
		pack_frags<2, false>(c, tmp);
		soa[0] = tmp[0][0]; soa[1] = tmp[0][1]; soa[2] = tmp[0][2]; soa[3] = tmp[0][3];
		if (t == 0) {
			sSDF[g] = __ushort_as_half((unsigned short)(soa[0] & 0xffffu));
			sSDF[g + 8] = __ushort_as_half((unsigned short)(soa[1] & 0xffffu));
		}
	}
	// --- d sdf / d input = W0^T (relu'(h0) * W1[0,:])   (nerf_network.h:228-245; seed = 1 on output 0)
	{
		uint32_t gh[4][4];
		#pragma unroll
		for (int kt = 0; kt < 4; ++kt) {
			const uint32_t wlo = *reinterpret_cast<const uint32_t*>(sw + WeightSmem::W1 + 16 * kt + 2 * t);
			const uint32_t whi = *reinterpret_cast<const uint32_t*>(sw + WeightSmem::W1 + 16 * kt + 8 + 2 * t);
			gh[kt][0] = wlo & relu_mask2(h0a[kt][0]);
			gh[kt][1] = wlo & relu_mask2(h0a[kt][1]);
			gh[kt][2] = whi & relu_mask2(h0a[kt][2]);
			gh[kt][3] = whi & relu_mask2(h0a[kt][3]);
		}
		float c[4][4];
		mma_layer<4, 4, true>(gh, sw + WeightSmem::W0, LD_X0, c, g, t);
		store_c_smem<4>(c, sDIN, LD_X0, g, t);
	}
	__syncwarp();

	// --- per-sample stage: normal (lanes 0..15), SH of the view direction (lanes 16..31)
	if (lane < 16) {
		const int s = lane;
		float n0 = 0.f, n1 = 0.f, n2 = 0.f;
		const int E = 2 * (int)gm.n_levels;
		for (int k = 0; k < E; ++k) {   // kernel_grid_backward_input, grid.h:803-830
			const float dl = __half2float(sDIN[s * LD_X0 + 3 + k]);
			n0 = fmaf(dl, sDY[s * LD_DY + 3 * k + 0], n0);
			n1 = fmaf(dl, sDY[s * LD_DY + 3 * k + 1], n1);
			n2 = fmaf(dl, sDY[s * LD_DY + 3 * k + 2], n2);
		}
		n0 = __fadd_rn(n0, __half2float(sDIN[s * LD_X0 + 0]));   // identity path, nerf_network.h:249
		n1 = __fadd_rn(n1, __half2float(sDIN[s * LD_X0 + 1]));
		n2 = __fadd_rn(n2, __half2float(sDIN[s * LD_X0 + 2]));
		sNRM[s * 3 + 0] = n0; sNRM[s * 3 + 1] = n1; sNRM[s * 3 + 2] = n2;
		__half* ci = sCIN + s * LD_X0;   // colour-input cols 16..47 live at [0..31]
		ci[16] = __float2half_rn(sCoord[s * 7 + 0]); ci[17] = __float2half_rn(sCoord[s * 7 + 1]); ci[18] = __float2half_rn(sCoord[s * 7 + 2]);
		ci[19] = __float2half_rn(n0); ci[20] = __float2half_rn(n1); ci[21] = __float2half_rn(n2);
		#pragma unroll
		for (int k = 22; k < 32; ++k) ci[k] = __float2half(0.f);
	} else {
		const int s = lane - 16;
		sh4(sCoord[s * 7 + 4], sCoord[s * 7 + 5], sCoord[s * 7 + 6], sCIN + s * LD_X0);
	}
	__syncwarp();

	// --- colour MLP (48 -> 64 -> 64 -> 16)
	uint32_t h1a[4][4], h2a[4][4];
	{
		uint32_t ca[3][4];
		ca[0][0] = soa[0]; ca[0][1] = soa[1]; ca[0][2] = soa[2]; ca[0][3] = soa[3];
		load_a_frag(ca[1], sCIN, LD_X0, 0, g, t);
		load_a_frag(ca[2], sCIN, LD_X0, 16, g, t);
		float c[8][4];
		mma_layer<3, 8>(ca, sw + WeightSmem::C0, LD_C0, c, g, t);
		pack_frags<8, true>(c, h1a);
	}
	{
		float c[8][4];
		mma_layer<4, 8>(h1a, sw + WeightSmem::C1, LD_H, c, g, t);
		pack_frags<8, true>(c, h2a);
	}
	if (!KEEP) {   // the training kernel needs the activations, not the outputs: the loss gradient arrives from outside
		{
			float c[2][4];
			mma_layer<4, 2>(h2a, sw + WeightSmem::C2, LD_H, c, g, t);
			store_c_smem<2>(c, sOUT, 16, g, t);
		}
		__syncwarp();
		// --- output assembly (nerf_network.h:287-325)
		if (lane < 16) {
			const int s = lane;
			__half* o = sOUT + s * 16;
			o[3] = __hadd(sSDF[s], __float2half_rn(sdf_bias));
			o[4] = __float2half_rn(sNRM[s * 3 + 0]); o[5] = __float2half_rn(sNRM[s * 3 + 1]); o[6] = __float2half_rn(sNRM[s * 3 + 2]);
			o[7] = params[pl.off_var];
			o[8] = __float2half_rn(sCoord[s * 7 + 4]); o[9] = __float2half_rn(sCoord[s * 7 + 5]); o[10] = __float2half_rn(sCoord[s * 7 + 6]);
		}
		__syncwarp();
	}
	if (KEEP) {
		#pragma unroll
		for (int kt = 0; kt < 4; ++kt) {
			#pragma unroll
			for (int r = 0; r < 4; ++r) { st.h0a[kt][r] = h0a[kt][r]; st.h1a[kt][r] = h1a[kt][r]; st.h2a[kt][r] = h2a[kt][r]; }
		}
		st.soa[0] = soa[0]; st.soa[1] = soa[1]; st.soa[2] = soa[2]; st.soa[3] = soa[3];
	}
}

// SDF-MLP only (NerfNetwork::sdf / ::density, nerf_network.h:656-739): positions in the slab's COORD rows -> raw sdf-MLP
// output 0 of the 16 samples in the slab's SDF row (half).  Needs W0 and W1 staged in `sw`.
__device__ __forceinline__ void sdf_tile(const GridMeta& gm, const __half* __restrict__ grid, const __half* sw, unsigned char* slab, int lane) {
	const int g = lane >> 2, t = lane & 3;
	float* sCoord = reinterpret_cast<float*>(slab + WarpSlab::COORD);
	__half* sX0 = reinterpret_cast<__half*>(slab + WarpSlab::X0);
	__half* sSDF = reinterpret_cast<__half*>(slab + WarpSlab::SDF);
	if (gm.delta_xform) delta_tile(gm.delta_xform, sCoord, lane, false);   // nerf_network.h:663-674
	if (lane < 16) {
		#pragma unroll
		for (int d = 0; d < 3; ++d) sX0[lane * LD_X0 + d] = __hsub(__float2half_rn(sCoord[lane * 7 + d]), __float2half_rn(0.5f));
		for (int k = 3 + 2 * (int)gm.n_levels; k < SDF_IN; ++k) sX0[lane * LD_X0 + k] = __float2half(0.f);
	}
	gather_tile<false, GATHER_BATCH_SDF>(gm, grid, sCoord, sX0, nullptr, lane);
	__syncwarp();
	uint32_t xa[2][4], h0a[4][4];
	load_a_frag(xa[0], sX0, LD_X0, 0, g, t);
	load_a_frag(xa[1], sX0, LD_X0, 16, g, t);
	{
		float c[8][4];
		mma_layer<2, 8>(xa, sw + WeightSmem::W0, LD_X0, c, g, t);
		pack_frags<8, true>(c, h0a);
	}
	{
		float c[1][4];   // only output n-tile 0 is needed (sdf = output 0)
		mma_layer<4, 1>(h0a, sw + WeightSmem::W1, LD_H, c, g, t);
		if (t == 0) { sSDF[g] = __float2half_rn(c[0][0]); sSDF[g + 8] = __float2half_rn(c[0][2]); }
	}
	__syncwarp();
}
