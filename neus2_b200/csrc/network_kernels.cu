// network_kernels.cu -- fused NeuS network kernels: inference forward, SDF/density only, and the fused
// forward+backward training kernel.  See network.cuh for the per-warp pipeline.
#include "network.cuh"
#include "umma.cuh"
#include "peer_sync.cuh"
#include "kernels.h"
#include <cstdlib>
#include <cstring>

#define WARPS_PER_CTA 8
#define THREADS_PER_CTA (WARPS_PER_CTA * 32)

static constexpr int FWD_SMEM = WeightSmem::FWD_END * 2 + WARPS_PER_CTA * WarpSlab::FWD_END;
#ifndef WARPS_BWD
#define WARPS_BWD 12
#endif
#define WARPS_BWD_LAUNCH_BOUND 12   // registers are budgeted for 12 warps (168 each) whatever WARPS_BWD is
#ifndef MERGE_MAX_HEADS
#define MERGE_MAX_HEADS 10     // pre-reduce a level's atomics in the warp when its 16 samples fall into <= this many runs of equal cells
#endif
// Training kernel, shared memory: weights | per-warp slabs | (TC) per-warp operand staging | (TC) one mbarrier per warp + the
// tensor-memory base address.  The DeltaNetwork variant has a larger slab and runs one warp less when it stages in shared memory.
#define WARPS_BWD_DELTA_TC (WARPS_BWD < 11 ? WARPS_BWD : 11)
static constexpr int BWD_TAIL = 16 * 8 + 16;   // mbarriers (8 B each, up to 16 warps) + tmem slot
template <bool DELTA, bool TC> struct BwdCfg {
	static constexpr int WARPS = (DELTA && TC) ? WARPS_BWD_DELTA_TC : WARPS_BWD;
	static constexpr int SLAB = DELTA ? WarpSlab::DELTA_END : WarpSlab::BWD_END;
	static constexpr int STAGE0 = WeightSmem::BWD_END * 2 + WARPS * SLAB;                 // first warp's staging block
	static constexpr int BARS = STAGE0 + (TC ? WARPS * WarpSlab::STAGE_BYTES : 0);
	static constexpr int SMEM = BARS + (TC ? BWD_TAIL : 0);
	static_assert(SMEM <= 227 * 1024, "training kernel exceeds the 227 KB shared-memory limit");
	static_assert(STAGE0 % 16 == 0 && BARS % 8 == 0, "staging / barrier alignment");
};
// tensor-memory columns of the five weight-gradient accumulators (fp32, M = 64 rows on lanes (m % 16) + 32 (m / 16))
#define TM_C2T 0      // dC2^T [64 hidden2][16 out]
#define TM_C1 16      // dC1   [64 out][64 in]
#define TM_C0 80      // dC0   [64 out][48 in]
#define TM_W1T 128    // dW1^T [64 hidden0][16 out]
#define TM_W0 144     // dW0   [64 out][32 in]   (first-order + second-order reverse sweep)
#define TM_COLS_USED 176
#define TM_COLS_ALLOC 256

// A tile's coords (16 x 7 floats) travel global -> registers -> warp slab: the loads for the NEXT tile are issued before the
// current tile is processed, so their HBM latency hides behind a whole tile of work.  Rows >= n are zero-filled; `wrap` maps
// tile row -> source row (0 = identity, else r % wrap: the fill_rollover wrap).
__device__ __forceinline__ void prefetch_coords(float (&cr)[4], const float* __restrict__ coords, uint32_t row0, uint32_t n, uint32_t wrap, int lane) {
	#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const int i = lane + 32 * k;
		const uint32_t r = row0 + i / 7;
		float v = 0.f;
		if (i < 16 * 7 && r < n) {
			const uint32_t src = wrap ? (r % wrap) : r;
			v = coords[(size_t)src * 7 + (i % 7)];
		}
		cr[k] = v;
	}
}
__device__ __forceinline__ void store_coords(float* sCoord, const float (&cr)[4], int lane) {
	#pragma unroll
	for (int k = 0; k < 4; ++k) { const int i = lane + 32 * k; if (i < 16 * 7) sCoord[i] = cr[k]; }
}

// ------------------------------------------------------------------------------------------------
// Inference forward: coords [n][7] -> out [n][16] half   (NerfNetwork::inference_mixed_precision_impl ==
// forward_impl, nerf_network.h:100-103,145-328)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS_PER_CTA, 2)
k_network_forward(GridMeta gm, ParamLayout pl, const __half* __restrict__ params, const float* __restrict__ coords,
                  const uint32_t* __restrict__ n_ptr, uint32_t n_max, float sdf_bias, __half* __restrict__ out) {
	extern __shared__ __align__(16) unsigned char smem[];
	__half* sw = reinterpret_cast<__half*>(smem);
	stage_weights(sw, params, pl, false, threadIdx.x, blockDim.x);
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	unsigned char* slab = smem + WeightSmem::FWD_END * 2 + warp * WarpSlab::FWD_END;
	uint32_t n = n_ptr ? *n_ptr : n_max;
	if (n > n_max) n = n_max;
	const uint32_t n_tiles = (n + TILE - 1) / TILE;
	FwdState st;
	float cr[4];
	const uint32_t tile0 = blockIdx.x * WARPS_PER_CTA + warp, tile_stride = gridDim.x * WARPS_PER_CTA;
	if (tile0 < n_tiles) prefetch_coords(cr, coords, tile0 * TILE, n, 0, lane);
	for (uint32_t tile = tile0; tile < n_tiles; tile += tile_stride) {
		const uint32_t row0 = tile * TILE;
		store_coords(reinterpret_cast<float*>(slab + WarpSlab::COORD), cr, lane);
		__syncwarp();
		if (tile + tile_stride < n_tiles) prefetch_coords(cr, coords, (tile + tile_stride) * TILE, n, 0, lane);
		forward_tile<false>(gm, params, pl, sw, slab, sdf_bias, lane, st);
		// 16 rows x 32 B, coalesced 16-byte stores
		const uint32_t r = row0 + (lane >> 1);
		if (r < n) {
			const uint4 v = reinterpret_cast<const uint4*>(slab + WarpSlab::OUT)[lane];
			reinterpret_cast<uint4*>(out)[(size_t)r * 2 + (lane & 1)] = v;
		}
		__syncwarp();
	}
}

// ------------------------------------------------------------------------------------------------
// SDF / density only: positions [n][stride] -> half [n]   (NerfNetwork::sdf / density, nerf_network.h:656-739;
// sdf_to_density_variance_buffer, common_operation.cuh:306-324)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS_PER_CTA, 2)
k_network_sdf(GridMeta gm, ParamLayout pl, const __half* __restrict__ params, const float* __restrict__ pos, uint32_t stride,
              uint32_t n, float sdf_bias, int as_density, __half* __restrict__ out_h, float* __restrict__ out_f) {
	extern __shared__ __align__(16) unsigned char smem[];
	__half* sw = reinterpret_cast<__half*>(smem);
	stage_matrix(sw + WeightSmem::W0, LD_X0, params + pl.off_w0, 64, SDF_IN, threadIdx.x, blockDim.x);
	stage_matrix(sw + WeightSmem::W1, LD_H, params + pl.off_w1, 16, 64, threadIdx.x, blockDim.x);
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	unsigned char* slab = smem + WeightSmem::FWD_END * 2 + warp * WarpSlab::FWD_END;
	float* sCoord = reinterpret_cast<float*>(slab + WarpSlab::COORD);
	__half* sSDF = reinterpret_cast<__half*>(slab + WarpSlab::SDF);
	const uint32_t n_tiles = (n + TILE - 1) / TILE;
	const __half variance = params[pl.off_var];
	for (uint32_t tile = blockIdx.x * WARPS_PER_CTA + warp; tile < n_tiles; tile += gridDim.x * WARPS_PER_CTA) {
		const uint32_t row0 = tile * TILE;
		for (int i = lane; i < 16 * 3; i += 32) {
			const uint32_t r = row0 + i / 3;
			sCoord[(i / 3) * 7 + (i % 3)] = r < n ? pos[(size_t)r * stride + (i % 3)] : 0.f;
		}
		__syncwarp();
		sdf_tile(gm, params + pl.off_grid, sw, slab, lane);
		if (lane < 16 && row0 + lane < n) {
			__half sdf = __hadd(sSDF[lane], __float2half_rn(sdf_bias));
			__half r = sdf;
			if (as_density) {
				// all in half arithmetic as the reference's T = __half instantiation
				const __half s = __float2half_rn(__expf(__half2float(__hmul(variance, __float2half_rn(10.0f)))));
				const float xs = __half2float(__hmul(sdf, s));
				const __half sig = __float2half_rn(1.0f / (1.0f + expf(-xs)));
				const __half density = __hmul(__hmul(s, sig), __hsub(__float2half_rn(1.0f), sig));
				r = density;
			}
			if (out_h) out_h[row0 + lane] = r;
			if (out_f) out_f[row0 + lane] = __half2float(r);
		}
		__syncwarp();
	}
}

// ------------------------------------------------------------------------------------------------
// Training: fused second forward + full backward (nerf_network.h:330-601) for the compacted batch.
//   rows i in [0, n_rows):  source row = i % n_in (fill_rollover, tcnn common_device.h:515-534); dL rows of
//   wrapped copies are scaled by n_in / n_rows and re-rounded to half.
// Outputs:
//   grads_f32[grid part]  += first + second order grid gradients (fp32 vector atomics)
//   var_acc               += sum of dL/dout row 7
//   w1row0_acc[64]        += column sums of the forward-mode sweep (second-order term of W1 row 0)
//   staged operand matrices (half, row-major [n_rows][w]) for the weight-gradient GEMMs.
// ------------------------------------------------------------------------------------------------
// DELTA (m_train_delta && m_use_delta, nerf_network.h:604-651 + DeltaNetwork::backward_impl, transform_network.h:110-250):
// the tile's coordinates pass through the DeltaNetwork inside forward_tile; the gradient that reaches its output
// (grid backward w.r.t. the position + the xyz rows of both MLP input gradients) is turned into the 9 parameter gradients
// per sample and summed: per lane in shared memory, per warp at the end, one atomicAdd per warp into delta_acc[9].
// TC (default): the five MLP weight gradients are accumulated inside this kernel by the 5th-generation tensor cores: each warp
// writes its tile's (dY, X) operand pairs to shared memory in the canonical MN-major layout and issues tcgen05.mma (M = 64,
// K = 16 samples) into accumulators that live in tensor memory for the whole kernel; they are read back once per CTA and added to
// the gradient buffer.  !TC: the operand pairs are staged in global memory for the split-K GEMM kernel of wgrad.cu (243 MB written,
// 290 MB read per step: the step's largest HBM stream, which the tensor-memory accumulators remove).
template <bool DELTA, bool TC>
__global__ void __launch_bounds__(WARPS_BWD_LAUNCH_BOUND * 32, 1)
k_network_fwdbwd(GridMeta gm, ParamLayout pl, const __half* __restrict__ params, const float* __restrict__ coords,
                 const __half* __restrict__ dL_dout, const uint32_t* __restrict__ n_in_ptr, uint32_t n_rows, float sdf_bias,
                 float batch_f, float* __restrict__ grads, float* __restrict__ var_acc, float* __restrict__ w1row0_acc, StagedOperands so, int dbg,
                 float* __restrict__ delta_acc, uint32_t* __restrict__ done_counter, PeerXfer grads_ready) {
	using Cfg = BwdCfg<DELTA, TC>;
	constexpr int NWARPS = Cfg::WARPS;
	extern __shared__ __align__(16) unsigned char smem[];
	__half* sw = reinterpret_cast<__half*>(smem);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int g = lane >> 2, t = lane & 3;
	unsigned char* stage = smem + Cfg::STAGE0 + warp * WarpSlab::STAGE_BYTES;
	unsigned char* sA = stage + WarpSlab::STAGE_A;
	unsigned char* sB = stage + WarpSlab::STAGE_B;
	uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BARS);
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::BARS + 16 * 8);
	uint32_t tmem = 0;
	uint32_t n_mma_groups = 0;      // commits this warp has issued: its mbarrier completes phase (n - 1) & 1 when the n-th group is done
	if (TC) {
		if (warp == 0) umma::tmem_alloc(tmem_slot, TM_COLS_ALLOC);
		if (threadIdx.x == 32) { for (int i = 0; i < NWARPS; ++i) umma::mbar_init(&bars[i], 1); umma::mbar_init_fence(); }
	}
	stage_weights(sw, params, pl, true, threadIdx.x, blockDim.x);
	if (TC) umma::fence_before_sync();
	__syncthreads();
	if (TC) {
		umma::fence_after_sync();
		tmem = *tmem_slot;
		// clear the accumulators: 0 = 0 * 0 with accumulate off, over all used columns, before any warp adds to them
		if (warp == 0) {
			for (int i = lane; i < 16 * 64 * 2 / 16; i += 32) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0u, 0u, 0u, 0u);
			umma::fence_proxy_async();
			__syncwarp();
			if (lane == 0) {
				umma::fence_after_sync();
				const uint64_t z = umma::make_desc(umma::smem_u32(sA));
				umma::mma_f16(tmem + 0, z, z, umma::make_idesc(64, 64), 0u);
				umma::mma_f16(tmem + 64, z, z, umma::make_idesc(64, 64), 0u);
				umma::mma_f16(tmem + 128, z, z, umma::make_idesc(64, 48), 0u);
				umma::commit(&bars[0]);
			}
			__syncwarp();
			umma::mbar_wait(&bars[0], 0u);
			n_mma_groups = 1;
		}
		umma::fence_before_sync();
		__syncthreads();
		umma::fence_after_sync();
	}
	// the previous tcgen05.mma group of this warp has read its staging block: the block may be rewritten
	auto stage_free = [&]() { if (TC && n_mma_groups && !(dbg & 24)) umma::mbar_wait(&bars[warp], (n_mma_groups - 1) & 1u); };
	// operands written: make them visible to the tensor core, then ONE lane issues the group and its commit
	auto issue = [&](uint32_t col, uint32_t a_off, uint32_t b_off, uint32_t idesc) {
		if (dbg & 8) return;      // NEUS2_DBG profiling bits (results are wrong): 8 = operands staged but never multiplied, 16 = no waits, 4 = no read-back
		umma::fence_proxy_async();
		__syncwarp();
		if (lane == 0) {
			umma::fence_after_sync();
			umma::mma_f16(tmem + col, umma::make_desc(umma::smem_u32(sA) + a_off), umma::make_desc(umma::smem_u32(sB) + b_off), idesc, 1u);
			umma::commit(&bars[warp]);
		}
		__syncwarp();
		++n_mma_groups;
	};
	unsigned char* slab = smem + WeightSmem::BWD_END * 2 + warp * Cfg::SLAB;
	float* sCoord = reinterpret_cast<float*>(slab + WarpSlab::COORD);
	__half* sX0 = reinterpret_cast<__half*>(slab + WarpSlab::X0);
	__half* sDIN = reinterpret_cast<__half*>(slab + WarpSlab::DIN);
	__half* sCIN = reinterpret_cast<__half*>(slab + WarpSlab::CIN);
	float* sDY = reinterpret_cast<float*>(slab + WarpSlab::DY);
	float* sORIG = reinterpret_cast<float*>(slab + WarpSlab::DORIG);
	float* sDACC = reinterpret_cast<float*>(slab + WarpSlab::DACC);
	if (DELTA && lane < 16) {
		#pragma unroll
		for (int k = 0; k < 9; ++k) sDACC[lane * 9 + k] = 0.f;
	}
	__half* sDL = reinterpret_cast<__half*>(slab + WarpSlab::DL);
	__half* sDX = reinterpret_cast<__half*>(slab + WarpSlab::DX);
	__half* sDI = reinterpret_cast<__half*>(slab + WarpSlab::DI);
	float* sV = reinterpret_cast<float*>(slab + WarpSlab::V);
	__half* sVN = reinterpret_cast<__half*>(slab + WarpSlab::VN);

	uint32_t n_in = *n_in_ptr;
	if (n_in > n_rows) n_in = n_rows;
	const uint32_t n_tiles = n_in == 0 ? 0 : (n_rows + TILE - 1) / TILE;
	const float wrap_scale_num = (float)n_in, wrap_scale_den = (float)n_rows;
	float var_local = 0.f;
	float w1acc[4][4];
	#pragma unroll
	for (int a = 0; a < 4; ++a) { w1acc[a][0] = w1acc[a][1] = w1acc[a][2] = w1acc[a][3] = 0.f; }
	float* ggrid = grads + pl.off_grid;
	const int E = 2 * (int)gm.n_levels;

	FwdState st;
	float cr[4];
	uint32_t dlr[4];   // the tile's dL/dout rows (16 x 16 halves = 128 words), 4 words per lane, prefetched like the coords
	auto prefetch_dl = [&](uint32_t row0) {
		#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const int w = lane + 32 * k;                       // word w: row w >> 3, halves 2 * (w & 7), +1
			const uint32_t r = row0 + (w >> 3);
			dlr[k] = r < n_rows ? reinterpret_cast<const uint32_t*>(dL_dout)[(size_t)(r < n_in ? r : r % n_in) * 8 + (w & 7)] : 0u;
		}
	};
	const uint32_t tile0 = blockIdx.x * NWARPS + warp, tile_stride = gridDim.x * NWARPS;
	if (tile0 < n_tiles) { prefetch_coords(cr, coords, tile0 * TILE, n_rows, n_in, lane); prefetch_dl(tile0 * TILE); }
	for (uint32_t tile = tile0; tile < n_tiles; tile += tile_stride) {
		const uint32_t row0 = tile * TILE;
		const int n_valid = (int)min((uint32_t)TILE, n_rows - row0);
		store_coords(sCoord, cr, lane);
		// dL/dout rows with the rollover wrap + rescale (fill_rollover_and_rescale)
		#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const int w = lane + 32 * k;
			const uint32_t r = row0 + (w >> 3);
			uint32_t v = dlr[k];
			if (r < n_rows && r >= n_in) {
				const float lo = __fdiv_rn(__fmul_rn(h_lo(v), wrap_scale_num), wrap_scale_den), hi = __fdiv_rn(__fmul_rn(h_hi(v), wrap_scale_num), wrap_scale_den);
				v = pack_h2(lo, hi);
			}
			reinterpret_cast<uint32_t*>(sDL)[w] = v;
		}
		__syncwarp();
		if (tile + tile_stride < n_tiles) { prefetch_coords(cr, coords, (tile + tile_stride) * TILE, n_rows, n_in, lane); prefetch_dl((tile + tile_stride) * TILE); }
		if (DELTA && lane < 16) {   // forward_tile deforms sCoord in place; the DeltaNetwork backward needs its input
			#pragma unroll
			for (int d = 0; d < 3; ++d) sORIG[lane * 3 + d] = sCoord[lane * 7 + d];
		}
		forward_tile<true>(gm, params, pl, sw, slab, sdf_bias, lane, st);

		// ---- colour MLP backward (fully_fused_mlp.cu:968-1085) -----------------------------------
		uint32_t dout[1][4];   // dL/drgb: only output columns 0..2 (extract_rgb, nerf_network.h:345-349)
		{
			uint32_t r0 = *reinterpret_cast<const uint32_t*>(sDL + g * 16 + 2 * t);
			uint32_t r1 = *reinterpret_cast<const uint32_t*>(sDL + (g + 8) * 16 + 2 * t);
			if (t == 1) { r0 &= 0xffffu; r1 &= 0xffffu; }
			if (t >= 2) { r0 = 0u; r1 = 0u; }
			dout[0][0] = r0; dout[0][1] = r1; dout[0][2] = 0u; dout[0][3] = 0u;
		}
		{   // dL/drgb rows [16][16] (columns 0..2, rest zero): one 16-byte store per lane
			const int r = lane >> 1;
			uint4 v = make_uint4(0u, 0u, 0u, 0u);
			if ((lane & 1) == 0) {
				v.x = *reinterpret_cast<const uint32_t*>(sDL + r * 16);
				v.y = *reinterpret_cast<const uint32_t*>(sDL + r * 16 + 2) & 0xffffu;
			}
			if (TC) {
				// dC2^T += H2^T dOUT: A = hidden 2 (64 wide), B = dL/drgb (16 wide).  Rows past the batch carry zero dL/dout, so every
				// operand pair below contributes nothing for them without masking.
				stage_free();
				umma::stage_frags<4>(st.h2a, sA, g, t);
				*reinterpret_cast<uint4*>(sB + (lane & 1) * UMMA_SBO + (r >> 3) * UMMA_LBO + (r & 7) * 16) = v;
				issue(TM_C2T, 0, 0, umma::make_idesc(64, 16));
			} else if (r < n_valid) reinterpret_cast<uint4*>(so.dOUT + (size_t)(row0 + r) * 16)[lane & 1] = v;
		}
		uint32_t dh2a[4][4], dh1a[4][4];
		{
			float c[8][4];
			mma_layer<1, 8, true>(dout, sw + WeightSmem::C2, LD_H, c, g, t);
			pack_frags_masked<8>(c, st.h2a, dh2a);
		}
		if (TC) {
			// dC1 += dH2^T H1, as two N = 32 halves: the second half of H1 is staged behind the next layer, when the first half has been read
			stage_free();
			umma::stage_frags<4>(dh2a, sA, g, t);
			{ const uint32_t lo[2][4] = {{st.h1a[0][0], st.h1a[0][1], st.h1a[0][2], st.h1a[0][3]}, {st.h1a[1][0], st.h1a[1][1], st.h1a[1][2], st.h1a[1][3]}}; umma::stage_frags<2>(lo, sB, g, t); }
			issue(TM_C1, 0, 0, umma::make_idesc(64, 32));
		} else {
			store_a_global<4>(st.h2a, so.H2, 64, row0, n_valid, g, t);
			store_a_global<4>(dh2a, so.dH2, 64, row0, n_valid, g, t);
		}
		{
			float c[8][4];
			mma_layer<4, 8, true>(dh2a, sw + WeightSmem::C1, LD_H, c, g, t);
			pack_frags_masked<8>(c, st.h1a, dh1a);
		}
		if (TC) {
			stage_free();
			{ const uint32_t hi[2][4] = {{st.h1a[2][0], st.h1a[2][1], st.h1a[2][2], st.h1a[2][3]}, {st.h1a[3][0], st.h1a[3][1], st.h1a[3][2], st.h1a[3][3]}}; umma::stage_frags<2>(hi, sB, g, t); }
			issue(TM_C1 + 32, 0, 0, umma::make_idesc(64, 32));     // A = dH2 is still in place
		} else {
			store_a_global<4>(st.h1a, so.H1, 64, row0, n_valid, g, t);
			store_a_global<4>(dh1a, so.dH1, 64, row0, n_valid, g, t);
		}
		uint32_t dso[1][4];
		{
			float c[6][4];
			mma_layer<4, 6, true>(dh1a, sw + WeightSmem::C0, LD_C0, c, g, t);
			// cols 0..15 -> gradient of the sdf-MLP output; + dL/dsdf on column 0 (add_density_gradient, nerf_network.h:401-408)
			float c00 = round_h(c[0][0]), c02 = round_h(c[0][2]);
			if (t == 0) {
				c00 = __half2float(__hadd(__float2half_rn(c[0][0]), sDL[g * 16 + 3]));
				c02 = __half2float(__hadd(__float2half_rn(c[0][2]), sDL[(g + 8) * 16 + 3]));
			}
			dso[0][0] = pack_h2(c00, c[0][1]);
			dso[0][1] = pack_h2(c02, c[0][3]);
			dso[0][2] = pack_h2(c[1][0], c[1][1]);
			dso[0][3] = pack_h2(c[1][2], c[1][3]);
			// cols 32..39 (n-tile 4) -> shared, the normal slots are 35..37
			*reinterpret_cast<uint32_t*>(sVN + g * 8 + 2 * t) = pack_h2(c[4][0], c[4][1]);
			*reinterpret_cast<uint32_t*>(sVN + (g + 8) * 8 + 2 * t) = pack_h2(c[4][2], c[4][3]);
		}
		// colour-MLP input [n][48]: cols 0..15 from registers, 16..47 from the slab
		if (TC) {
			// dC0 += dH1^T CIN
			stage_free();
			umma::stage_frags<4>(dh1a, sA, g, t);
			{ const uint32_t tmp[1][4] = {{st.soa[0], st.soa[1], st.soa[2], st.soa[3]}}; umma::stage_frags<1>(tmp, sB, g, t); }
			umma::stage_rows32(reinterpret_cast<const unsigned char*>(sCIN), LD_X0 * 2, sB, 2, lane);
			issue(TM_C0, 0, 0, umma::make_idesc(64, 48));
		} else {
			uint32_t tmp[1][4] = {{st.soa[0], st.soa[1], st.soa[2], st.soa[3]}};
			store_a_global<1>(tmp, so.CIN, 48, row0, n_valid, g, t);
			for (int i = lane; i < 16 * 16; i += 32) {   // 16 rows x 16 words (32 halves)
				const int r = i >> 4, w = i & 15;
				if (r < n_valid) reinterpret_cast<uint32_t*>(so.CIN + (size_t)(row0 + r) * 48 + 16)[w] = reinterpret_cast<const uint32_t*>(sCIN + r * LD_X0)[w];
			}
		}
		// ---- sdf MLP backward --------------------------------------------------------------------
		uint32_t dh0a[4][4];
		{
			float c[8][4];
			mma_layer<1, 8, true>(dso, sw + WeightSmem::W1, LD_H, c, g, t);
			pack_frags_masked<8>(c, st.h0a, dh0a);
		}
		if (TC) {
			// dW1^T += H0^T dSO
			stage_free();
			umma::stage_frags<4>(st.h0a, sA, g, t);
			umma::stage_frags<1>(dso, sB, g, t);
			issue(TM_W1T, 0, 0, umma::make_idesc(64, 16));
		} else {
			store_a_global<1>(dso, so.dSO, 16, row0, n_valid, g, t);
			store_a_global<4>(st.h0a, so.H0, 64, row0, n_valid, g, t);
			store_a_global<4>(dh0a, so.dH0, 64, row0, n_valid, g, t);
		}
		{
			float c[4][4];
			mma_layer<4, 4, true>(dh0a, sw + WeightSmem::W0, LD_X0, c, g, t);
			store_c_smem<4>(c, sDX, LD_X0, g, t);
		}
		// sdf-MLP input [n][32]
		if (TC) {
			// dW0 += dH0^T X0
			stage_free();
			umma::stage_frags<4>(dh0a, sA, g, t);
			umma::stage_rows32(reinterpret_cast<const unsigned char*>(sX0), LD_X0 * 2, sB, 0, lane);
			issue(TM_W0, 0, 0, umma::make_idesc(64, 32));
		} else {
			for (int i = lane; i < 16 * 16; i += 32) {
				const int r = i >> 4, w = i & 15;
				if (r < n_valid) reinterpret_cast<uint32_t*>(so.X0 + (size_t)(row0 + r) * 32)[w] = reinterpret_cast<const uint32_t*>(sX0 + r * LD_X0)[w];
			}
		}
		__syncwarp();
		// ---- per-sample: v = dL/d normal, pos_encoding_dy, second-order sdf-MLP input (nerf_network.h:476-570)
		if (lane < 16) {
			const int s = lane;
			float v[3];
			#pragma unroll
			for (int d = 0; d < 3; ++d) {
				float a = __half2float(sVN[s * 8 + 3 + d]);
				a = __fadd_rn(a, __fdiv_rn(__half2float(sDL[s * 16 + 4 + d]), batch_f));
				a = __fadd_rn(a, __half2float(sDL[s * 16 + 8 + d]));
				v[d] = a;
				sV[s * 3 + d] = a;
				sDI[s * LD_X0 + d] = __float2half_rn(a);
			}
			for (int k = 0; k < E; ++k) {   // kernel_grid_backward_input_backward_dLdoutput, grid.h:1183-1207
				float r = 0.f;
				r = fmaf(sDY[s * LD_DY + 3 * k + 0], v[0], r);
				r = fmaf(sDY[s * LD_DY + 3 * k + 1], v[1], r);
				r = fmaf(sDY[s * LD_DY + 3 * k + 2], v[2], r);
				sDI[s * LD_X0 + 3 + k] = __float2half_rn(r);
			}
			for (int k = 3 + E; k < SDF_IN; ++k) sDI[s * LD_X0 + k] = __float2half(0.f);
			if (s < n_valid) var_local += __half2float(sDL[s * 16 + 7]);
			if (DELTA && s < n_valid) {
				// dL/d(deformed position): kernel_grid_backward_input on the first-order dL/d enc (nerf_network.h:423-442), then
				// += colour-input xyz rows, += sdf-MLP-input xyz rows (add_positions_view<float, T>, :607-626)
				float L[3] = {0.f, 0.f, 0.f};
				for (int k = 0; k < E; ++k) {
					const float dl = __half2float(sDX[s * LD_X0 + 3 + k]);
					L[0] = fmaf(dl, sDY[s * LD_DY + 3 * k + 0], L[0]);
					L[1] = fmaf(dl, sDY[s * LD_DY + 3 * k + 1], L[1]);
					L[2] = fmaf(dl, sDY[s * LD_DY + 3 * k + 2], L[2]);
				}
				float xin[3], o9[9];
				#pragma unroll
				for (int d = 0; d < 3; ++d) {
					L[d] = __fadd_rn(L[d], __half2float(sVN[s * 8 + d]));
					L[d] = __fadd_rn(L[d], __half2float(sDX[s * LD_X0 + d]));
					xin[d] = sORIG[s * 3 + d];
				}
				delta_backward_sample(gm.delta_xform, L, xin, o9);
				#pragma unroll
				for (int k = 0; k < 9; ++k) sDACC[s * 9 + k] += o9[k];
			}
		}
		__syncwarp();
		// ---- second-order sdf-MLP sweeps (fully_fused_mlp.cu:1089-1198)
		{
			// (the reverse-sweep operand relu'(h0) * W1[0,:] is rebuilt from the staged H0 inside the weight-gradient kernel)
			// forward-mode sweep: relu'(h0) * (W0 d_in)  -> contributes to dW1 row 0 (column sums over samples)
			uint32_t da[2][4];
			load_a_frag(da[0], sDI, LD_X0, 0, g, t);
			load_a_frag(da[1], sDI, LD_X0, 16, g, t);
			float c[8][4];
			mma_layer<2, 8>(da, sw + WeightSmem::W0, LD_X0, c, g, t);
			uint32_t tf[4][4];
			pack_frags_masked<8>(c, st.h0a, tf);
			#pragma unroll
			for (int kt = 0; kt < 4; ++kt) {
				// rows g (regs 0,2) and g+8 (regs 1,3); invalid rows were zero-input so they add 0 -- but mask anyway
				const float m0 = g < n_valid ? 1.f : 0.f, m1 = g + 8 < n_valid ? 1.f : 0.f;
				w1acc[kt][0] += m0 * h_lo(tf[kt][0]) + m1 * h_lo(tf[kt][1]);
				w1acc[kt][1] += m0 * h_hi(tf[kt][0]) + m1 * h_hi(tf[kt][1]);
				w1acc[kt][2] += m0 * h_lo(tf[kt][2]) + m1 * h_lo(tf[kt][3]);
				w1acc[kt][3] += m0 * h_hi(tf[kt][2]) + m1 * h_hi(tf[kt][3]);
			}
			if (TC) {
				// dW0 += gH^T dI with gH = relu'(h0) * W1[0,:]: the reverse sweep of the second-order term (seed e0, fully_fused_mlp.cu:1163-1174)
				uint32_t gh[4][4];
				#pragma unroll
				for (int kt = 0; kt < 4; ++kt) {
					const uint32_t wlo = *reinterpret_cast<const uint32_t*>(sw + WeightSmem::W1 + 16 * kt + 2 * t);
					const uint32_t whi = *reinterpret_cast<const uint32_t*>(sw + WeightSmem::W1 + 16 * kt + 8 + 2 * t);
					gh[kt][0] = wlo & relu_mask2(st.h0a[kt][0]);
					gh[kt][1] = wlo & relu_mask2(st.h0a[kt][1]);
					gh[kt][2] = whi & relu_mask2(st.h0a[kt][2]);
					gh[kt][3] = whi & relu_mask2(st.h0a[kt][3]);
				}
				stage_free();
				umma::stage_frags<4>(gh, sA, g, t);
				umma::stage_rows32(reinterpret_cast<const unsigned char*>(sDI), LD_X0 * 2, sB, 0, lane);
				issue(TM_W0, 0, 0, umma::make_idesc(64, 32));
			} else {
				for (int i = lane; i < 16 * 16; i += 32) {
					const int r = i >> 4, w = i & 15;
					if (r < n_valid) reinterpret_cast<uint32_t*>(so.dI + (size_t)(row0 + r) * 32)[w] = reinterpret_cast<const uint32_t*>(sDI + r * LD_X0)[w];
				}
			}
		}
		// ---- hash-grid gradient: first order (grid.h:371-500) + second order (grid.h:880-1007), one vector
		//      atomic per corner instead of 8 + 24 half2 atomics.  Samples of a tile are consecutive along a ray, so at the
		//      coarse levels neighbouring lanes fall into the SAME cell: contributions of such a run of lanes are summed
		//      with a segmented shuffle reduction and only the run head issues the atomics (warp-level pre-reduction).
		{
			const int s = lane & 15;
			const bool active = s < n_valid;
			const float x = sCoord[s * 7 + 0], y = sCoord[s * 7 + 1], z = sCoord[s * 7 + 2];
			const float v0 = sV[s * 3 + 0], v1 = sV[s * 3 + 1], v2 = sV[s * 3 + 2];
			const uint32_t half_mask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;
			for (int l = lane >> 4; l < (int)gm.n_levels; l += 2) {   // both half-warps run the same number of iterations +-1
				if ((uint32_t)l > gm.valid_level) continue;
				LevelGeom lg;
				level_geom(gm, l, x, y, z, lg);
				const float scale = gm.scale[l];
				const float d0 = __half2float(sDX[s * LD_X0 + 3 + 2 * l]), d1 = __half2float(sDX[s * LD_X0 + 4 + 2 * l]);   // dL/d enc
				const float e0 = __half2float(sDIN[s * LD_X0 + 3 + 2 * l]), e1 = __half2float(sDIN[s * LD_X0 + 4 + 2 * l]); // d sdf/d enc
				const float wx[2] = {1.f - lg.w[0], lg.w[0]}, wy[2] = {1.f - lg.w[1], lg.w[1]}, wz[2] = {1.f - lg.w[2], lg.w[2]};
				const float gi0 = scale * v0, gi1 = scale * v1, gi2 = scale * v2;   // grad_in (pos_derivative = 1)
				float a0[8], a1[8];
				#pragma unroll
				for (int c = 0; c < 8; ++c) {
					const int bx = c & 1, by = (c >> 1) & 1, bz = (c >> 2) & 1;
					const float w1 = (wx[bx] * wy[by]) * wz[bz];
					// second order: sum over grad_dim of +-(grad_in * w_other0 * w_other1)
					const float sx = (gi0 * wy[by]) * wz[bz], sy = (gi1 * wx[bx]) * wz[bz], sz = (gi2 * wx[bx]) * wy[by];
					const float w2 = (bx ? sx : -sx) + (by ? sy : -sy) + (bz ? sz : -sz);
					a0[c] = active ? d0 * w1 + e0 * w2 : 0.f;
					a1[c] = active ? d1 * w1 + e1 * w2 : 0.f;
				}
				// runs of consecutive lanes (within this half-warp) that sit in the same grid cell; all 16 lanes of the
				// half-warp reach every shuffle below (no divergent exits), the two half-warps sync independently
				const uint32_t ka = lg.cell[0] | (lg.cell[1] << 16), kb = lg.cell[2];
				const uint32_t pa = __shfl_up_sync(half_mask, ka, 1), pb = __shfl_up_sync(half_mask, kb, 1);
				const bool head = (s == 0) || ka != pa || kb != pb;
				const uint32_t heads = __ballot_sync(half_mask, head);
				bool issue = active;
				if (__popc(heads) <= MERGE_MAX_HEADS) {   // enough merging to pay for the reduction (coarse levels); uniform per half-warp
					const uint32_t hs16 = (heads >> (lane & 16)) & 0xffffu;
					const int run = __popc(hs16 & ((2u << s) - 1u));
					#pragma unroll
					for (int d = 1; d < 16; d <<= 1) {
						const int other = __shfl_down_sync(half_mask, run, d);
						const bool take = (s + d < 16) && other == run;
						#pragma unroll
						for (int c = 0; c < 8; ++c) {
							const float t0 = __shfl_down_sync(half_mask, a0[c], d), t1 = __shfl_down_sync(half_mask, a1[c], d);
							if (take) { a0[c] += t0; a1[c] += t1; }
						}
					}
					issue = head;   // inactive lanes contributed zeros; a head that is itself inactive still carries its run's sum
				}
				if (issue) {
					float* gl = ggrid + (size_t)gm.offset[l] * 2;
					#pragma unroll
					for (int c = 0; c < 8; ++c) {
						if (!(dbg & 1)) red_add_v2f32(gl + lg.idx[c], a0[c], a1[c]);
					}
				}
			}
		}
		__syncwarp();
	}
	// ---- flush per-warp accumulators
	#pragma unroll
	for (int kt = 0; kt < 4; ++kt) {
		#pragma unroll
		for (int r = 0; r < 4; ++r) {
			float v = w1acc[kt][r];
			v += __shfl_xor_sync(0xffffffffu, v, 4);
			v += __shfl_xor_sync(0xffffffffu, v, 8);
			v += __shfl_xor_sync(0xffffffffu, v, 16);
			// second-order term of W1 row 0: TC adds it to the gradient directly, !TC leaves it to k_reduce_partials
			if (g == 0 && v != 0.f) red_add_f32((TC ? grads + pl.off_w1 : w1row0_acc) + 16 * kt + 8 * (r >> 1) + 2 * t + (r & 1), v);
		}
	}
	var_local += __shfl_xor_sync(0xffffffffu, var_local, 1);
	var_local += __shfl_xor_sync(0xffffffffu, var_local, 2);
	var_local += __shfl_xor_sync(0xffffffffu, var_local, 4);
	var_local += __shfl_xor_sync(0xffffffffu, var_local, 8);
	var_local += __shfl_xor_sync(0xffffffffu, var_local, 16);
	if (lane == 0 && var_local != 0.f) red_add_f32(var_acc, var_local);
	if (DELTA) {
		__syncwarp();
		#pragma unroll
		for (int k = 0; k < 9; ++k) {
			float v = lane < 16 ? sDACC[lane * 9 + k] : 0.f;
			v += __shfl_xor_sync(0xffffffffu, v, 1);
			v += __shfl_xor_sync(0xffffffffu, v, 2);
			v += __shfl_xor_sync(0xffffffffu, v, 4);
			v += __shfl_xor_sync(0xffffffffu, v, 8);
			if (lane == 0 && v != 0.f) red_add_f32(delta_acc + k, v);
		}
	}
	if (TC) {
		// ---- weight gradients: tensor memory -> gradient buffer, once per CTA ----
		stage_free();                          // this warp's last group is complete
		umma::fence_before_sync();
		__syncthreads();                       // ... and so are everybody else's
		umma::fence_after_sync();
		if (n_tiles > 0 && warp < 4 && !(dbg & 4)) {
			// warp q reads lanes 32 q .. 32 q + 31; rows 16 q .. 16 q + 15 of every accumulator sit on its first 16 lanes
			const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
			const int m = 16 * warp + lane;
			const bool row = lane < 16;
			float v[16];
			umma::tmem_ld16(tmem + lane_base + TM_C2T, v);      // dC2[o][m]
			if (row) { _Pragma("unroll") for (int o = 0; o < 16; ++o) red_add_f32(grads + pl.off_c2 + o * 64 + m, v[o]); }
			umma::tmem_ld16(tmem + lane_base + TM_W1T, v);      // dW1[o][m]
			if (row) { _Pragma("unroll") for (int o = 0; o < 16; ++o) red_add_f32(grads + pl.off_w1 + o * 64 + m, v[o]); }
			#pragma unroll
			for (int c = 0; c < 64; c += 16) {                  // dC1[m][c ..]
				umma::tmem_ld16(tmem + lane_base + TM_C1 + c, v);
				if (row) { _Pragma("unroll") for (int j = 0; j < 16; j += 4) red_add_v4f32(grads + pl.off_c1 + m * 64 + c + j, v[j], v[j + 1], v[j + 2], v[j + 3]); }
			}
			#pragma unroll
			for (int c = 0; c < 48; c += 16) {                  // dC0[m][c ..]
				umma::tmem_ld16(tmem + lane_base + TM_C0 + c, v);
				if (row) { _Pragma("unroll") for (int j = 0; j < 16; j += 4) red_add_v4f32(grads + pl.off_c0 + m * 48 + c + j, v[j], v[j + 1], v[j + 2], v[j + 3]); }
			}
			#pragma unroll
			for (int c = 0; c < 32; c += 16) {                  // dW0[m][c ..]
				umma::tmem_ld16(tmem + lane_base + TM_W0 + c, v);
				if (row) { _Pragma("unroll") for (int j = 0; j < 16; j += 4) red_add_v4f32(grads + pl.off_w0 + m * pl.sdf_in_w + c + j, v[j], v[j + 1], v[j + 2], v[j + 3]); }
			}
		}
		__threadfence();                       // every thread's gradient reductions are ordered before the arrival counter below
		umma::fence_before_sync();
		__syncthreads();
		if (warp == 0) umma::tmem_dealloc(tmem, TM_COLS_ALLOC);
		// the variance gradient is the batch sum rounded ONCE to half (nerf_network.h:463-474): the last CTA to get here writes it
		__shared__ int s_last;
		if (threadIdx.x == 0) {
			__threadfence();
			const uint32_t arrived = atomicInc(done_counter, gridDim.x - 1);    // wraps to 0 with the last arrival: ready for the next launch
			s_last = arrived == gridDim.x - 1;
			if (s_last) {
				__threadfence();
				grads[pl.off_var] = round_h(*reinterpret_cast<volatile float*>(var_acc));
			}
		}
		// Data-parallel: "this rank's gradients are complete" is posted to the peers, and theirs awaited, by the same last CTA (+ the
		// DeltaNetwork's gradient sums, summed over ranks in place): the sharded optimizer behind this kernel may read every peer's buffer.
		// A kernel of its own would queue behind the next step's sample generation, which fills every SM this kernel frees.
		if (grads_ready.world) {
			__syncthreads();
			if (s_last && warp == 0) { __threadfence_system(); peer_exchange_warp(grads_ready, nullptr, nullptr, nullptr, DELTA ? delta_acc : nullptr, lane); }
		}
	}
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static int num_sms() { return device_sm_count(); }

int launch_network_forward(cudaStream_t stream, const GridMeta& gm, const ParamLayout& pl, const __half* params, const float* coords,
                           const uint32_t* n_ptr, uint32_t n_max, float sdf_bias, __half* out) {
	static bool attr[NEUS_MAX_DEVICES] = {};
	if (ensure_dynamic_smem(k_network_forward, FWD_SMEM, attr)) return -1;
	if (n_max == 0) return 0;
	const uint32_t n_tiles = (n_max + TILE - 1) / TILE;
	uint32_t grid = (uint32_t)num_sms() * 2;
	const uint32_t need = (n_tiles + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
	if (need < grid) grid = need;
	k_network_forward<<<grid, THREADS_PER_CTA, FWD_SMEM, stream>>>(gm, pl, params, coords, n_ptr, n_max, sdf_bias, out);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}

int launch_network_sdf(cudaStream_t stream, const GridMeta& gm, const ParamLayout& pl, const __half* params, const float* pos, uint32_t stride,
                       uint32_t n, float sdf_bias, int as_density, __half* out_h, float* out_f) {
	static bool attr[NEUS_MAX_DEVICES] = {};
	if (ensure_dynamic_smem(k_network_sdf, FWD_SMEM, attr)) return -1;
	if (n == 0) return 0;
	const uint32_t n_tiles = (n + TILE - 1) / TILE;
	uint32_t grid = (uint32_t)num_sms() * 2;
	const uint32_t need = (n_tiles + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
	if (need < grid) grid = need;
	k_network_sdf<<<grid, THREADS_PER_CTA, FWD_SMEM, stream>>>(gm, pl, params, pos, stride, n, sdf_bias, as_density, out_h, out_f);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}

template <bool DELTA, bool TC>
static int launch_fwdbwd_variant(cudaStream_t stream, const GridMeta& gm, const ParamLayout& pl, const __half* params, const float* coords,
                                 const __half* dL_dout, const uint32_t* n_in_ptr, uint32_t n_rows, float sdf_bias, uint32_t indeed_batch_size,
                                 float* grads, float* var_acc, float* w1row0_acc, const StagedOperands& so, float* delta_acc, uint32_t* done_counter, int dbg,
                                 const PeerXfer& grads_ready) {
	using Cfg = BwdCfg<DELTA, TC>;
	static bool attr[NEUS_MAX_DEVICES] = {};
	if (ensure_dynamic_smem(k_network_fwdbwd<DELTA, TC>, Cfg::SMEM, attr)) return -1;
	const uint32_t n_tiles = (n_rows + TILE - 1) / TILE;
	uint32_t grid = (uint32_t)num_sms();
	const uint32_t need = (n_tiles + Cfg::WARPS - 1) / Cfg::WARPS;
	if (need < grid) grid = need;
	k_network_fwdbwd<DELTA, TC><<<grid, Cfg::WARPS * 32, Cfg::SMEM, stream>>>(gm, pl, params, coords, dL_dout, n_in_ptr, n_rows, sdf_bias,
		(float)indeed_batch_size, grads, var_acc, w1row0_acc, so, dbg, delta_acc, done_counter, grads_ready);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}

// Profiling switches of the training kernel (NEUS2_DBG or neus2_set_debug_bits; results are WRONG with any bit set): 1 = no hash-grid
// atomics, 4 = no read-back of the tensor-memory accumulators, 8 = weight-gradient operands staged but never multiplied, 16 = no waits
// for tcgen05 completion.  tools/bench_fwdbwd.py times the kernel on a frozen training state with each of them.
int g_fwdbwd_debug_bits = -1;

// NEUS2_WGRAD=staged selects the global-memory operand staging + split-K GEMM kernel (wgrad.cu) instead of the in-kernel tensor-memory
// accumulators: the A/B switch behind the measurements in DESIGN.md.  Read once per process.
bool weight_gradients_in_kernel() {
	static int tc = -1;
	if (tc < 0) { const char* e = getenv("NEUS2_WGRAD"); tc = (e && !strcmp(e, "staged")) ? 0 : 1; }
	return tc != 0;
}

int launch_network_fwdbwd(cudaStream_t stream, const GridMeta& gm, const ParamLayout& pl, const __half* params, const float* coords,
                          const __half* dL_dout, const uint32_t* n_in_ptr, uint32_t n_rows, float sdf_bias, uint32_t indeed_batch_size,
                          float* grads, float* var_acc, float* w1row0_acc, const StagedOperands& so, float* delta_acc, uint32_t* done_counter,
                          const PeerXfer* grads_ready_exchange) {
	if (n_rows == 0 && !grads_ready_exchange) return 0;
	if (n_rows == 0) { neus_set_error("a data-parallel rank cannot skip the training kernel: its peers wait for it", __FILE__, __LINE__); return -1; }
	PeerXfer gx = {};
	if (grads_ready_exchange) gx = *grads_ready_exchange;
	const bool delta = gm.delta_xform != nullptr && delta_acc != nullptr;
	if (g_fwdbwd_debug_bits < 0) { const char* e = getenv("NEUS2_DBG"); g_fwdbwd_debug_bits = e ? atoi(e) : 0; }
	const int dbg = g_fwdbwd_debug_bits;
	const bool tc = weight_gradients_in_kernel();
	if (!tc && grads_ready_exchange) { neus_set_error("the gradient-ready exchange rides in the training kernel only with in-kernel weight gradients", __FILE__, __LINE__); return -1; }
	if (!tc && !so.H2) { neus_set_error("NEUS2_WGRAD=staged needs the staged-operand buffers (set before the context is created)", __FILE__, __LINE__); return -1; }
	if (delta) {
		if (!tc) return launch_fwdbwd_variant<true, false>(stream, gm, pl, params, coords, dL_dout, n_in_ptr, n_rows, sdf_bias, indeed_batch_size, grads, var_acc, w1row0_acc, so, delta_acc, done_counter, dbg, gx);
		return launch_fwdbwd_variant<true, true>(stream, gm, pl, params, coords, dL_dout, n_in_ptr, n_rows, sdf_bias, indeed_batch_size, grads, var_acc, w1row0_acc, so, delta_acc, done_counter, dbg, gx);
	}
	if (!tc) return launch_fwdbwd_variant<false, false>(stream, gm, pl, params, coords, dL_dout, n_in_ptr, n_rows, sdf_bias, indeed_batch_size, grads, var_acc, w1row0_acc, so, nullptr, done_counter, dbg, gx);
	return launch_fwdbwd_variant<false, true>(stream, gm, pl, params, coords, dL_dout, n_in_ptr, n_rows, sdf_bias, indeed_batch_size, grads, var_acc, w1row0_acc, so, nullptr, done_counter, dbg, gx);
}
