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
#define WG_CTAS_PER_JOB 96

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_ptr) {
	uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_ptr);
	asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
	             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// copy a [WG_ROWS][W] half tile (rows row0.. of a row-major [n][W] matrix) to shared [WG_ROWS][W+8]; rows >= n -> 0
template <int W>
__device__ __forceinline__ void load_tile(__half* dst, const __half* __restrict__ src, uint32_t row0, uint32_t n, int tid) {
	constexpr int V = W / 8;   // uint4 per row
	for (int i = tid; i < WG_ROWS * V; i += WG_THREADS) {
		const int r = i / V, c = i - r * V;
		uint4 v = make_uint4(0, 0, 0, 0);
		if (row0 + r < n) v = reinterpret_cast<const uint4*>(src + (size_t)(row0 + r) * W)[c];
		*reinterpret_cast<uint4*>(dst + r * (W + 8) + c * 8) = v;
	}
}

// One job: partial[NOUT][NIN] over rows [row_begin, row_end) for up to two (dY, X) operand pairs.
template <int NOUT, int NIN>
__device__ __forceinline__ void wgrad_job(const __half* dY0, const __half* X0, const __half* dY1, const __half* X1,
                                          uint32_t row_begin, uint32_t row_end, uint32_t n, float* __restrict__ partial, __half* smem) {
	constexpr int MT = NOUT / 16;                 // m tiles
	constexpr int NT = NIN / 8;                   // n tiles
	constexpr int NT_W = (MT == 4) ? NT : NT / 4; // n tiles per warp
	static_assert(MT == 4 || MT == 1, "unsupported shape");
	static_assert(NT_W % 2 == 0, "n tiles per warp must be even");
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int g = lane >> 2, t = lane & 3;
	const int m0 = (MT == 4) ? 16 * warp : 0;
	const int nt0 = (MT == 4) ? 0 : warp * NT_W;
	__half* sY = smem;                            // [WG_ROWS][NOUT+8]
	__half* sX = smem + WG_ROWS * (NOUT + 8);     // [WG_ROWS][NIN+8]
	float acc[NT_W][4];
	#pragma unroll
	for (int i = 0; i < NT_W; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
	for (int seg = 0; seg < 2; ++seg) {
		const __half* dY = seg ? dY1 : dY0;
		const __half* X = seg ? X1 : X0;
		if (!dY) continue;
		for (uint32_t r0 = row_begin; r0 < row_end; r0 += WG_ROWS) {
			__syncthreads();
			load_tile<NOUT>(sY, dY, r0, n, tid);
			load_tile<NIN>(sX, X, r0, n, tid);
			__syncthreads();
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
		}
	}
	#pragma unroll
	for (int i = 0; i < NT_W; ++i) {
		const int col = 8 * (nt0 + i) + 2 * t;
		*reinterpret_cast<float2*>(partial + (size_t)(m0 + g) * NIN + col) = make_float2(acc[i][0], acc[i][1]);
		*reinterpret_cast<float2*>(partial + (size_t)(m0 + g + 8) * NIN + col) = make_float2(acc[i][2], acc[i][3]);
	}
}

__global__ void __launch_bounds__(WG_THREADS)
k_weight_gradients(ParamLayout pl, StagedOperands so, const uint32_t* __restrict__ n_in_ptr,
                   uint32_t n_rows, float* __restrict__ partials, uint32_t n_mlp) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	__half* smem = reinterpret_cast<__half*>(smem_raw);
	const uint32_t n = (*n_in_ptr == 0) ? 0 : n_rows;
	const uint32_t tiles = (n + WG_ROWS - 1) / WG_ROWS;
	const uint32_t per = (tiles + gridDim.x - 1) / gridDim.x;
	const uint32_t rb = min(n, blockIdx.x * per * WG_ROWS), re = min(n, (blockIdx.x + 1) * per * WG_ROWS);
	float* p = partials + (size_t)blockIdx.x * n_mlp;
	switch (blockIdx.y) {
		case 0: wgrad_job<16, 64>(so.dOUT, so.H2, nullptr, nullptr, rb, re, n, p + pl.off_c2, smem); break;
		case 1: wgrad_job<64, 64>(so.dH2, so.H1, nullptr, nullptr, rb, re, n, p + pl.off_c1, smem); break;
		case 2: wgrad_job<64, 48>(so.dH1, so.CIN, nullptr, nullptr, rb, re, n, p + pl.off_c0, smem); break;
		case 3: wgrad_job<16, 64>(so.dSO, so.H0, nullptr, nullptr, rb, re, n, p + pl.off_w1, smem); break;
		default: wgrad_job<64, 32>(so.dH0, so.X0, so.gH, so.dI, rb, re, n, p + pl.off_w0, smem); break;
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

uint32_t weight_gradient_partials() { return WG_CTAS_PER_JOB; }

int launch_weight_gradients(cudaStream_t stream, const ParamLayout& pl, const StagedOperands& so,
                            const uint32_t* n_in_ptr, uint32_t n_rows, float* partials, uint32_t n_partials,
                            const float* var_acc, const float* w1row0_acc, float* grads) {
	const uint32_t n_mlp = pl.off_grid;
	const int smem = WG_ROWS * (64 + 8) * 2 * 2;
	dim3 grid(n_partials, 5);
	k_weight_gradients<<<grid, WG_THREADS, smem, stream>>>(pl, so, n_in_ptr, n_rows, partials, n_mlp);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	k_reduce_partials<<<(n_mlp + 255) / 256, 256, 0, stream>>>(pl, partials, n_partials, n_mlp, var_acc, w1row0_acc, grads);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}
