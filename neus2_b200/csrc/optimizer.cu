// optimizer.cu -- fused Adam (+ per-parameter step counters, sparse skip, L2 on the MLP weights) and fp16 EMA.
// Restates adam_step (tcnn optimizers/adam.h:52-160) and ema_step_half_precision (optimizers/ema.h:64-77,112-147)
// as ONE pass over the parameters (the reference launches two kernels and re-reads the fp16 params).
// Compiled with -fmad=false so the update is bit-identical to oracle/neus_oracle.cpp:orc_adam_ema_step.
#include "ngp_common.cuh"
#include "ngp_kernels.h"

__global__ void k_adam_ema(uint64_t n, uint64_t n_matrix, float loss_scale, float lr, float beta1, float beta2, float eps, float l2_reg,
                           float ema_decay, float ema_debias_old, float ema_debias_new, const float* __restrict__ grads,
                           float* __restrict__ master, __half* __restrict__ params, float* __restrict__ m1, float* __restrict__ m2,
                           uint32_t* __restrict__ steps, __half* __restrict__ ema) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float gradient = grads[i] / loss_scale;
	const bool is_matrix = i < n_matrix;
	__half p = params[i];
	if (is_matrix || gradient != 0) {            // adam.h:109-112: non-matrix params with zero gradient are skipped entirely
		const float w = master[i];
		if (is_matrix) gradient += l2_reg * w;   // adam.h:132-135
		const float gsq = gradient * gradient;
		const float fm = beta1 * m1[i] + (1 - beta1) * gradient;
		const float sm = beta2 * m2[i] + (1 - beta2) * gsq;
		m1[i] = fm; m2[i] = sm;
		float learning_rate = lr;
		const uint32_t cs = ++steps[i];
		learning_rate *= sqrtf(1 - powf(beta2, (float)cs)) / (1 - powf(beta1, (float)cs));
		const float eff = fminf(fmaxf(learning_rate / (sqrtf(sm) + eps), 0.f), 3.402823466e+38f);
		const float decayed = (1 - 0.f * learning_rate) * w - copysignf(0.f * learning_rate, w);
		const float nw = decayed - eff * fm;
		master[i] = nw;
		p = __float2half_rn(nw);
		params[i] = p;
	}
	const float filtered = (__half2float(ema[i]) * ema_decay * ema_debias_old + __half2float(p) * (1 - ema_decay)) * ema_debias_new;
	ema[i] = __float2half_rn(filtered);
}

int launch_adam_ema(cudaStream_t stream, uint64_t n, uint64_t n_matrix, float loss_scale, float lr, float beta1, float beta2, float eps, float l2_reg,
                    float ema_decay, uint32_t ema_step, const float* grads, float* master, __half* params, float* m1, float* m2, uint32_t* steps, __half* ema) {
	const float ema_debias_old = 1 - (float)pow((double)ema_decay, (double)(ema_step - 1));
	const float ema_debias_new = 1.0f / (1 - (float)pow((double)ema_decay, (double)ema_step));
	const uint32_t blocks = (uint32_t)((n + 255) / 256);
	k_adam_ema<<<blocks, 256, 0, stream>>>(n, n_matrix, loss_scale, lr, beta1, beta2, eps, l2_reg, ema_decay, ema_debias_old, ema_debias_new,
	                                      grads, master, params, m1, m2, steps, ema);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}
