// kernels.h -- internal launcher declarations shared between the .cu files and the C-ABI layer.
#pragma once
#include "common.cuh"

// Operand matrices staged by the fused training kernel for the weight-gradient GEMMs (half, row-major [rows][w]).
struct StagedOperands {
	__half* dOUT; // [n][16] dL/drgb (cols 0..2 of the wrapped/rescaled dL/dout, rest 0)
	__half* H2;   // [n][64] colour hidden 2
	__half* dH2;  // [n][64]
	__half* H1;   // [n][64]
	__half* dH1;  // [n][64]
	__half* CIN;  // [n][48] colour-MLP input
	__half* dSO;  // [n][16] gradient of the sdf-MLP output
	__half* H0;   // [n][64] sdf hidden
	__half* dH0;  // [n][64]
	__half* X0;   // [n][32] sdf-MLP input
	__half* dI;   // [n][32] d_density_input           (second-order forward seed)
};

int launch_network_forward(cudaStream_t stream, const GridMeta& gm, const ParamLayout& pl, const __half* params, const float* coords,
                           const uint32_t* n_ptr, uint32_t n_max, float sdf_bias, __half* out);
int launch_network_sdf(cudaStream_t stream, const GridMeta& gm, const ParamLayout& pl, const __half* params, const float* pos, uint32_t stride,
                       uint32_t n, float sdf_bias, int as_density, __half* out_h, float* out_f);
int launch_network_fwdbwd(cudaStream_t stream, const GridMeta& gm, const ParamLayout& pl, const __half* params, const float* coords,
                          const __half* dL_dout, const uint32_t* n_in_ptr, uint32_t n_rows, float sdf_bias, uint32_t indeed_batch_size,
                          float* grads, float* var_acc, float* w1row0_acc, const StagedOperands& so, float* delta_acc, uint32_t* done_counter,
                          const struct PeerXfer* grads_ready_exchange = nullptr /* data-parallel: the kernel's last CTA runs this exchange (peer_sync.cuh) */);
// true (default): the MLP weight gradients are accumulated inside the training kernel (tcgen05 + tensor memory), launch_weight_gradients is not
// needed and the staged-operand buffers are not allocated; false (NEUS2_WGRAD=staged): the round-1 path through wgrad.cu
bool weight_gradients_in_kernel();
extern int g_fwdbwd_debug_bits;

// weight-gradient GEMMs: grads[mlp part] += sum over rows (see wgrad.cu)
int launch_weight_gradients(cudaStream_t stream, const ParamLayout& pl, const StagedOperands& so, const __half* params,
                            const uint32_t* n_in_ptr, uint32_t n_rows, float* partials, uint32_t n_partials,
                            const float* var_acc, const float* w1row0_acc, float* grads);
uint32_t weight_gradient_partials();   // number of per-CTA partial slots `partials` must hold (x n_mlp_params floats)
