// wgrad.cu -- MLP weight gradients: dW[o][i] = sum_rows dY[row][o] * X[row][i], a skinny GEMM with K = batch
// (2^18 rows), computed split-K over CTAs from the operand matrices staged by k_network_fwdbwd.
//
// Restates the reference's CUTLASS split-K calls (tcnn src/fully_fused_mlp.cu:1020,1055,1066 first order;
// :1176-1187 second order) -- fp32 accumulation here instead of the reference's fp16 accumulators
// (cutlass_matmul.h:83-84).  Both operands are stored row-major [row][width], i.e. "K-outer": fragments are
// fetched with ldmatrix.trans.
#include "kernels.h"

#define WG_THREADS 128
#define WG_ROWS 64            // rows (K) per shared-memory tile
// CTAs per job, proportional to the bytes a job streams per row (160 / 256 / 224 / 160 / 384 B): all jobs finish together.
// Sum = 5 x 148; the partial buffer has WG_MAX_CTAS slots per parameter, slots a job never writes stay zero (allocated zeroed).
#define WG_MAX_CTAS 240
__constant__ uint32_t c_wg_ctas[5] = {100, 160, 140, 100, 240};
static const uint32_t h_wg_ctas[5] = {100, 160, 140, 100, 240};

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_ptr) {
	uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_ptr);
	asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
	             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
	const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
	const int src_size = valid ? 16 : 0;   // src_size 0 -> zero fill
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(dst), "l"(gmem_src), "r"(src_size));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }

// async copy of a [WG_ROWS][W] half tile (rows row0.. of a row-major [n][W] matrix) to shared [WG_ROWS][W+8]; rows >= n -> 0
template <int W>
__device__ __forceinline__ void load_tile_async(__half* dst, const __half* __restrict__ src, uint32_t row0, uint32_t n, int tid) {
	constexpr int V = W / 8;   // 16-byte chunks per row
	for (int i = tid; i < WG_ROWS * V; i += WG_THREADS) {
		const int r = i / V, c = i - r * V;
		const bool valid = row0 + r < n;
		cp_async16(dst + r * (W + 8) + c * 8, src + (size_t)(valid ? row0 + r : 0) * W + c * 8, valid);
	}
}

// One job: partial[NOUT][NIN] over rows [row_begin, row_end) for up to two (dY, X) operand pairs, double-buffered with cp.async.
// If w1row0 != nullptr the dY operand of the SECOND pair is derived on the fly from the staged hidden activations:
//   gH = relu'(h0) * W1[0,:]   (reverse sweep of the second-order term with seed e0, fully_fused_mlp.cu:1163-1174)
template <int NOUT, int NIN>
__device__ __forceinline__ void wgrad_job(const __half* dY0, const __half* X0, const __half* dY1, const __half* X1, const __half* w1row0,
                                          uint32_t row_begin, uint32_t row_end, uint32_t n, float* __restrict__ partial, __half* smem) {
	constexpr int MT = NOUT / 16;                 // m tiles
	constexpr int NT = NIN / 8;                   // n tiles
	constexpr int NT_W = (MT == 4) ? NT : NT / 4; // n tiles per warp
	static_assert(MT == 4 || MT == 1, "unsupported shape");
	static_assert(NT_W % 2 == 0, "n tiles per warp must be even");
	constexpr int STAGE = WG_ROWS * (NOUT + 8) + WG_ROWS * (NIN + 8);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int g = lane >> 2, t = lane & 3;
	const int m0 = (MT == 4) ? 16 * warp : 0;
	const int nt0 = (MT == 4) ? 0 : warp * NT_W;
	float acc[NT_W][4];
	#pragma unroll
	for (int i = 0; i < NT_W; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
	const uint32_t tiles_per_seg = (row_end - row_begin + WG_ROWS - 1) / WG_ROWS;
	const uint32_t n_tiles = tiles_per_seg * (dY1 ? 2u : 1u);
	auto issue = [&](uint32_t ti) {
		const uint32_t seg = ti / tiles_per_seg, r0 = row_begin + (ti - seg * tiles_per_seg) * WG_ROWS;
		__half* sY = smem + (ti & 1) * STAGE;
		__half* sX = sY + WG_ROWS * (NOUT + 8);
		load_tile_async<NOUT>(sY, seg ? dY1 : dY0, r0, n, tid);
		load_tile_async<NIN>(sX, seg ? X1 : X0, r0, n, tid);
		cp_async_commit();
	};
	if (n_tiles) issue(0);
	for (uint32_t ti = 0; ti < n_tiles; ++ti) {
		if (ti + 1 < n_tiles) { issue(ti + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
		__syncthreads();
		__half* sY = smem + (ti & 1) * STAGE;
		const __half* sX = sY + WG_ROWS * (NOUT + 8);
		if (w1row0 && ti >= tiles_per_seg) {
			for (int i = tid; i < WG_ROWS * (NOUT / 2); i += WG_THREADS) {
				const int r = i / (NOUT / 2), c = i - r * (NOUT / 2);
				uint32_t* p = reinterpret_cast<uint32_t*>(sY + r * (NOUT + 8) + 2 * c);
				const uint32_t h = *p;
				const uint32_t w = *reinterpret_cast<const uint32_t*>(w1row0 + 2 * c);
				*p = w & __hgt2_mask(*reinterpret_cast<const __half2*>(&h), __half2half2(__ushort_as_half((unsigned short)0)));
			}
			__syncthreads();
		}
		#pragma unroll
		for (int ks = 0; ks < WG_ROWS / 16; ++ks) {
			uint32_t a[4];
			{   // A(m,k) = dY[k][m]: matrices (k0-7,m0-7) (k0-7,m8-15) (k8-15,m0-7) (k8-15,m8-15)
				const int q = lane >> 3, r = lane & 7;
				ldmatrix_x4_trans(a, sY + (16 * ks + (q >> 1) * 8 + r) * (NOUT + 8) + m0 + (q & 1) * 8);
			}
			#pragma unroll
			for (int p = 0; p < NT_W / 2; ++p) {
				uint32_t b[4];   // B(k,n) = X[k][n]: (k0-7,n0-7) (k8-15,n0-7) (k0-7,n8-15) (k8-15,n8-15)
				const int q = lane >> 3, r = lane & 7;
				ldmatrix_x4_trans(b, sX + (16 * ks + (q & 1) * 8 + r) * (NIN + 8) + 8 * (nt0 + 2 * p) + (q >> 1) * 8);
				mma16816(acc[2 * p], a, b[0], b[1]);
				mma16816(acc[2 * p + 1], a, b[2], b[3]);
			}
		}
		__syncthreads();   // the buffer is refilled two iterations later by issue(ti + 2)... which runs right after this barrier
	}
	#pragma unroll
	for (int i = 0; i < NT_W; ++i) {
		const int col = 8 * (nt0 + i) + 2 * t;
		*reinterpret_cast<float2*>(partial + (size_t)(m0 + g) * NIN + col) = make_float2(acc[i][0], acc[i][1]);
		*reinterpret_cast<float2*>(partial + (size_t)(m0 + g + 8) * NIN + col) = make_float2(acc[i][2], acc[i][3]);
	}
}

__global__ void __launch_bounds__(WG_THREADS)
k_weight_gradients(ParamLayout pl, StagedOperands so, const __half* __restrict__ params, const uint32_t* __restrict__ n_in_ptr,
                   uint32_t n_rows, float* __restrict__ partials, uint32_t n_mlp) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	__half* smem = reinterpret_cast<__half*>(smem_raw);
	__shared__ __align__(16) __half s_w1row0[64];
	if (threadIdx.x < 64) s_w1row0[threadIdx.x] = params[pl.off_w1 + threadIdx.x];
	__syncthreads();
	const uint32_t n = (*n_in_ptr == 0) ? 0 : n_rows;
	uint32_t job = 0, part = blockIdx.x;
	while (job < 4 && part >= c_wg_ctas[job]) { part -= c_wg_ctas[job]; ++job; }
	const uint32_t n_parts = c_wg_ctas[job];
	const uint32_t tiles = (n + WG_ROWS - 1) / WG_ROWS;
	const uint32_t per = (tiles + n_parts - 1) / n_parts;
	const uint32_t rb = min(n, part * per * WG_ROWS), re = min(n, (part + 1) * per * WG_ROWS);
	float* p = partials + (size_t)part * n_mlp;
	switch (job) {
		case 0: wgrad_job<16, 64>(so.dOUT, so.H2, nullptr, nullptr, nullptr, rb, re, n, p + pl.off_c2, smem); break;
		case 1: wgrad_job<64, 64>(so.dH2, so.H1, nullptr, nullptr, nullptr, rb, re, n, p + pl.off_c1, smem); break;
		case 2: wgrad_job<64, 48>(so.dH1, so.CIN, nullptr, nullptr, nullptr, rb, re, n, p + pl.off_c0, smem); break;
		case 3: wgrad_job<16, 64>(so.dSO, so.H0, nullptr, nullptr, nullptr, rb, re, n, p + pl.off_w1, smem); break;
		default: wgrad_job<64, 32>(so.dH0, so.X0, so.H0, so.dI, s_w1row0, rb, re, n, p + pl.off_w0, smem); break;
	}
}

// grads[i] = sum_c partials[c][i] (+ w1 row-0 second-order column sums); variance gradient = half(sum) (nerf_network.h:463-474)
__global__ void k_reduce_partials(ParamLayout pl, const float* __restrict__ partials, uint32_t n_partials, uint32_t n_mlp,
                                  const float* __restrict__ var_acc, const float* __restrict__ w1row0_acc, float* __restrict__ grads) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n_mlp) {
		float s = 0.f;
		for (uint32_t c = 0; c < n_partials; ++c) s += partials[(size_t)c * n_mlp + i];
		if (i >= pl.off_w1 && i < pl.off_w1 + 64) s += w1row0_acc[i - pl.off_w1];
		grads[i] = s;
	}
	if (i == 0) grads[pl.off_var] = round_h(*var_acc);
}

uint32_t weight_gradient_partials() { return WG_MAX_CTAS; }

int launch_weight_gradients(cudaStream_t stream, const ParamLayout& pl, const StagedOperands& so, const __half* params,
                            const uint32_t* n_in_ptr, uint32_t n_rows, float* partials, uint32_t n_partials,
                            const float* var_acc, const float* w1row0_acc, float* grads) {
	const uint32_t n_mlp = pl.off_grid;
	const int smem = 2 * (WG_ROWS * (64 + 8) * 2) * 2;   // two stages of (dY tile + X tile)
	if (n_partials != WG_MAX_CTAS) { neus_set_error("partials buffer must hold weight_gradient_partials() slots", __FILE__, __LINE__); return -1; }
	uint32_t grid = 0;
	for (int j = 0; j < 5; ++j) grid += h_wg_ctas[j];
	k_weight_gradients<<<grid, WG_THREADS, smem, stream>>>(pl, so, params, n_in_ptr, n_rows, partials, n_mlp);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	k_reduce_partials<<<(n_mlp + 255) / 256, 256, 0, stream>>>(pl, partials, n_partials, n_mlp, var_acc, w1row0_acc, grads);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}
