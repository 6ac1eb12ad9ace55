// loss.cu -- NeuS compositing, loss and dL/d(network output) (compiled with -fmad=false).
// Restates compute_loss_kernel_train_nerf_with_global_movement (src/testbed_nerf.cu:1475-1997) for Byte (LDR)
// images: rgb activation Logistic, Huber(0.1)/5 in sRGB space, NeuS logistic-CDF alpha with the learned inv_s,
// mask term, eikonal rows, cos rows.  Float/double literal types follow the reference expression by expression.
// The reference's atomicAdd compaction (:1682) becomes an exclusive scan in ray-slot order (deterministic).
#include "ngp_common.cuh"
#include "ngp_kernels.h"
#include "lookback.cuh"

__device__ __forceinline__ void ray_dir_for_loss(const float* __restrict__ ray, const GlobalMove& gmv, float ray_o[3], float dir[3]) {
	ray_o[0] = ray[0]; ray_o[1] = ray[1]; ray_o[2] = ray[2];
	normalize3(ray + 3, dir);
	#pragma unroll
	for (int k = 0; k < 3; ++k) ray_o[k] -= gmv.first_frame_offset[k];
	global_move(gmv, ray_o, dir);
}

// ---- the order-dependent part of compositing: T *= 1 - alpha, rgb += w * rgb_j, weight_sum += w, stop once T < 1e-4 ----
// Every lane of a warp evaluates one sample of the current 32-sample chunk; the recurrences over the chunk are then replayed in sample
// order by all lanes alike, so every per-ray value is bit-identical to the reference's serial loop.  The replay is what bounds the loss
// kernels (rays carry ~70 samples here, the longest ones several hundred, and whoever comes behind a long ray waits for it): the lanes'
// {alpha, r, g, b} are staged in shared memory and consumed eight samples at a time -- eight broadcast 16-byte loads up front, then eight
// branch-free recurrence steps (the early-out is a sticky predicate), i.e. ~10 cycles per sample on the dependent chain instead of a
// compare -> branch -> four shuffles -> multiply round trip per sample.
struct CompositeState { float T, weight_sum, rgb[3]; uint32_t compacted; bool stop; };
struct SampleCapture { float weight, T, rgb[3]; };   // of one sample: its weight, the transmittance behind it, the colour up to and including it

__device__ __forceinline__ void stage_chunk(float4* __restrict__ stage, const SampleEval& e, int lane) {
	__syncwarp();   // the previous chunk's replay has read its samples
	stage[lane] = make_float4(e.alpha, e.rgb[0], e.rgb[1], e.rgb[2]);
	__syncwarp();
}

// pass 1 (:1567-1627): consumes up to `cnt` samples of the staged chunk, stops (sticky) before the first sample whose incoming T is below 1e-4
template <bool CAPTURE>
__device__ __forceinline__ void replay_composite(const float4* __restrict__ stage, uint32_t cnt, int lane, CompositeState& s, SampleCapture& cap) {
	const float EPSILON = 1e-4f;
	for (uint32_t L = 0; L < cnt && !s.stop; L += 8) {
		float4 v[8];
		#pragma unroll
		for (int k = 0; k < 8; ++k) v[k] = stage[L + k];   // L + k < 32: the chunk buffer is 32 entries and L a multiple of 8; entries >= cnt hold zeros
		#pragma unroll
		for (int k = 0; k < 8; ++k) {
			const bool valid = L + k < cnt;
			s.stop = s.stop || (valid && s.T < EPSILON);
			const bool go = valid && !s.stop;
			const float weight = v[k].x * s.T;
			const float n0 = s.rgb[0] + weight * v[k].y, n1 = s.rgb[1] + weight * v[k].z, n2 = s.rgb[2] + weight * v[k].w;
			const float nws = s.weight_sum + weight;
			const float nT = s.T * (1.f - v[k].x);
			if (go) {
				s.rgb[0] = n0; s.rgb[1] = n1; s.rgb[2] = n2; s.weight_sum = nws; s.T = nT; ++s.compacted;
				if (CAPTURE && lane == (int)(L + k)) { cap.weight = weight; cap.T = nT; cap.rgb[0] = n0; cap.rgb[1] = n1; cap.rgb[2] = n2; }
			}
		}
	}
}

// The same pass-1 recurrences with the dependent chains cut down to what is serial by nature (k_loss_rays): lane j holds sample j of the chunk.
//   1. T_0 .. T_32 = the running product of the staged (1 - alpha_j), one multiplication per sample, by all lanes alike; lane j keeps T_j
//      (the transmittance in front of its sample) and T_{j+1}.  T never grows, so the samples the serial loop consumes are those in front
//      of the first j with T_j < 1e-4 (one ballot).
//   2. weight_j = alpha_j * T_j and weight_j * rgb_j are per-lane products of the very operands the serial loop multiplies.
//   3. weight_sum and the three colour sums add those products in sample order (four independent additions per sample); samples that are
//      not consumed contribute +0, which leaves every (non-negative) running sum unchanged.
// Same values, bit for bit, as replay_composite; ~4 cycles per sample on each dependent chain instead of a compare / predicate / multiply
// round trip.  `a`, `r`, `g`, `b`: this lane's sample (zeros behind the end of the ray); returns how many samples of the chunk were consumed.
__device__ __forceinline__ uint32_t composite_chunk(float4* __restrict__ stage, float a, float r, float g, float b, uint32_t cnt, int lane,
                                                     CompositeState& s, SampleCapture& cap) {
	const float EPSILON = 1e-4f;
	float* keep = reinterpret_cast<float*>(stage);
	__syncwarp();   // whoever used the staging area before is done with it
	keep[lane] = 1.f - a;
	__syncwarp();
	float T = s.T, T_in = s.T, T_out = s.T;
	#pragma unroll
	for (int q = 0; q < 8; ++q) {
		const float4 k4 = reinterpret_cast<const float4*>(keep)[q];
		const float kk[4] = {k4.x, k4.y, k4.z, k4.w};
		#pragma unroll
		for (int u = 0; u < 4; ++u) {
			if (lane == 4 * q + u) T_in = T;
			T *= kk[u];
			if (lane == 4 * q + u) T_out = T;
		}
	}
	const uint32_t in_chunk = cnt >= 32u ? 0xffffffffu : ((1u << cnt) - 1u);
	const uint32_t below = __ballot_sync(0xffffffffu, T_in < EPSILON) & in_chunk;
	const uint32_t consumed = below ? (uint32_t)__ffs((int)below) - 1u : cnt;
	const bool mine = (uint32_t)lane < consumed;
	const float w = a * T_in;
	__syncwarp();   // every lane has read the staged (1 - alpha)
	stage[lane] = mine ? make_float4(w, w * r, w * g, w * b) : make_float4(0.f, 0.f, 0.f, 0.f);
	__syncwarp();
	float ws = s.weight_sum, c0 = s.rgb[0], c1 = s.rgb[1], c2 = s.rgb[2];
	for (uint32_t L = 0; L < consumed; L += 8) {
		float4 v[8];
		#pragma unroll
		for (int k = 0; k < 8; ++k) v[k] = stage[L + k];
		#pragma unroll
		for (int k = 0; k < 8; ++k) {
			c0 += v[k].y; c1 += v[k].z; c2 += v[k].w; ws += v[k].x;
			if (lane == (int)(L + k)) { cap.rgb[0] = c0; cap.rgb[1] = c1; cap.rgb[2] = c2; }
		}
	}
	cap.weight = w; cap.T = T_out;
	s.rgb[0] = c0; s.rgb[1] = c1; s.rgb[2] = c2; s.weight_sum = ws;
	s.compacted += consumed;
	s.stop = below != 0u;
	s.T = below ? __shfl_sync(0xffffffffu, T_in, (int)consumed) : T;   // stopped in front of sample `consumed`; else the product over the whole chunk ((1 - alpha) = 1 behind its end)
	return consumed;
}

// pass 2 (:1800-1830): the same recurrences over `cnt` kept samples, no early-out; lane j keeps the state right behind sample j
__device__ __forceinline__ void replay_capture(const float4* __restrict__ stage, uint32_t cnt, int lane, float& T, float rgb[3], SampleCapture& cap) {
	for (uint32_t L = 0; L < cnt; L += 8) {
		float4 v[8];
		#pragma unroll
		for (int k = 0; k < 8; ++k) v[k] = stage[L + k];
		#pragma unroll
		for (int k = 0; k < 8; ++k) {
			const float weight = v[k].x * T;
			const float n0 = rgb[0] + weight * v[k].y, n1 = rgb[1] + weight * v[k].z, n2 = rgb[2] + weight * v[k].w;
			const float nT = T * (1.f - v[k].x);
			if (L + k < cnt) {
				rgb[0] = n0; rgb[1] = n1; rgb[2] = n2; T = nT;
				if (lane == (int)(L + k)) { cap.weight = weight; cap.T = nT; cap.rgb[0] = n0; cap.rgb[1] = n1; cap.rgb[2] = n2; }
			}
		}
	}
}

// pass 1: composite, one WARP per ray.  Lanes evaluate 32 consecutive samples in parallel (the expensive part: three
// logistic colours, exp(10*var), two logistic CDFs, a division); the order-dependent recurrences (T *= 1-alpha,
// rgb_ray += w*rgb, weight_sum += w, and the T < 1e-4 early-out that fixes the compacted sample count) are then
// replayed in sample order from warp shuffles, so every per-ray value is bit-identical to the serial loop.
// ray_tmp[i] = {compacted (bits), rgb_ray[3], weight_sum, T, dir[3]}
__global__ void __launch_bounds__(128)
k_loss_composite(const uint32_t* __restrict__ n_kept_ptr, GlobalMove gmv, float cos_anneal_ratio,
                 const __half* __restrict__ network_output, const float* __restrict__ rays, const uint32_t* __restrict__ numsteps,
                 const float* __restrict__ coords, float* __restrict__ ray_tmp, uint32_t* __restrict__ compacted_counts) {
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (i >= *n_kept_ptr) return;
	const uint32_t ns = numsteps[i * 2 + 0], base = numsteps[i * 2 + 1];
	const float* cin = coords + (size_t)base * 7;
	const __half* nout = network_output + (size_t)base * 16;
	float T = 1.f;
	float rgb_ray[3] = {0.f, 0.f, 0.f};
	float weight_sum = 0.f;
	uint32_t compacted = 0;
	float ray_o[3], dir[3];
	ray_dir_for_loss(rays + (size_t)i * 6, gmv, ray_o, dir);
	{   // BENT_DIR :1583-1588 (from sample 0)
		float dv[3] = {__half2float(nout[8]) * 2.0f - 1.0f, __half2float(nout[9]) * 2.0f - 1.0f, __half2float(nout[10]) * 2.0f - 1.0f};
		if (ns > 0) normalize3(dv, dir);
	}
	__shared__ float4 s_stage[4][32];
	float4* stage = s_stage[threadIdx.x >> 5];
	CompositeState cs = {T, weight_sum, {0.f, 0.f, 0.f}, 0u, false};
	SampleCapture unused;
	for (uint32_t c0 = 0; c0 < ns && !cs.stop; c0 += 32) {
		const uint32_t idx = c0 + lane;
		SampleEval e;
		e.alpha = 0.f; e.rgb[0] = e.rgb[1] = e.rgb[2] = 0.f;
		if (idx < ns) e = eval_sample(nout + (size_t)idx * 16, cin + (size_t)idx * 7, dir, cos_anneal_ratio);
		stage_chunk(stage, e, lane);
		replay_composite<false>(stage, min(32u, ns - c0), lane, cs, unused);
	}
	T = cs.T; weight_sum = cs.weight_sum; compacted = cs.compacted;
	rgb_ray[0] = cs.rgb[0]; rgb_ray[1] = cs.rgb[1]; rgb_ray[2] = cs.rgb[2];
	if (lane == 0) {
		float* rt = ray_tmp + (size_t)i * 12;
		rt[0] = __uint_as_float(compacted);
		rt[1] = rgb_ray[0]; rt[2] = rgb_ray[1]; rt[3] = rgb_ray[2];
		rt[4] = weight_sum; rt[5] = T;
		rt[6] = dir[0]; rt[7] = dir[1]; rt[8] = dir[2];
		compacted_counts[i] = compacted;
	}
}

#define SCAN_THREADS 1024
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_compacted(const uint32_t* __restrict__ n_ptr, const uint32_t* __restrict__ counts, uint32_t* __restrict__ base, uint32_t* __restrict__ total_out) {
	__shared__ uint32_t warp_sums[32];
	const uint32_t n = *n_ptr;
	const uint32_t per = (n + SCAN_THREADS - 1) / SCAN_THREADS;
	const uint32_t b = min(n, threadIdx.x * per), e = min(n, b + per);
	uint32_t s = 0;
	for (uint32_t i = b; i < e; ++i) s += counts[i];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc = s;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
	if (lane == 31) warp_sums[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		uint32_t w = warp_sums[lane], winc = w;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += y; }
		warp_sums[lane] = winc - w;
		if (lane == 31) *total_out = winc;
	}
	__syncthreads();
	uint32_t off = warp_sums[warp] + inc - s;
	for (uint32_t i = b; i < e; ++i) { base[i] = off; off += counts[i]; }
}

// ---- pieces shared by the warp-per-ray kernels (k_loss_composite / k_loss_gradient) and the by-samples kernels (k_loss_rays / k_loss_samples) ----
// What the ray is compared against (:1629-1676): the training pixel under the ray, composited over the (random) background, in sRGB.
struct RayTarget { float rgbtarget[3]; float bg[3]; float mask_gt; };
__device__ __forceinline__ RayTarget ray_target(uint32_t ray_idx, uint32_t n_rays, uint32_t n_rays_total, Pcg32 rng, const DatasetDev& ds, const LossParamsDev& lp) {
	RayTarget t;
	rng.advance((int64_t)(uint32_t)(ray_idx * N_MAX_RANDOM_SAMPLES_PER_RAY));
	const uint32_t img = image_idx(ray_idx, n_rays, n_rays_total, ds.n_images);
	float xy[2];
	random_image_pos_training(rng, ds.width, ds.height, lp.snap_to_pixel_centers != 0, xy);
	t.bg[0] = lp.background_color[0]; t.bg[1] = lp.background_color[1]; t.bg[2] = lp.background_color[2];
	if (lp.random_bg_color) { t.bg[0] = rng.next_float(); t.bg[1] = rng.next_float(); t.bg[2] = rng.next_float(); }
	#pragma unroll
	for (int k = 0; k < 3; ++k) t.bg[k] = srgb_to_linear(t.bg[k]);
	int px, py; image_pos(xy, ds.width, ds.height, &px, &py);
	float tex[4]; read_rgba(ds, img, px, py, tex);
	#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const float exposure_scale = expf(0.6931471805599453f * 0.0f);
		t.rgbtarget[k] = exposure_scale * tex[k] + (1.0f - tex[3]) * t.bg[k];
		t.rgbtarget[k] = linear_to_srgb(t.rgbtarget[k]);
		t.bg[k] = linear_to_srgb(t.bg[k]);
	}
	t.mask_gt = (float)(tex[3] > 0.9999);
	return t;
}

// Per-ray scalars of pass 2 (:1690-1760): Huber loss and its gradient per channel, the mask term's gradient through the weight sum.
struct RayLoss { float rgb_ray[3]; float lg_grad[3]; float gradient_weight_sum, weight_sum, mean_loss, mask_loss; };
__device__ __forceinline__ RayLoss ray_loss(const float rgb_ray[3], float weight_sum, const RayTarget& t, const LossParamsDev& lp) {
	RayLoss r;
	float lg_loss[3];
	#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const float alpha_h = 0.1f;
		const float difference = rgb_ray[k] - t.rgbtarget[k];
		const float abs_diff = fabsf(difference);
		const float square = 0.5f / alpha_h * difference * difference;
		lg_loss[k] = (abs_diff > alpha_h ? (abs_diff - 0.5f * alpha_h) : square) / 5.0f;
		r.lg_grad[k] = (abs_diff > alpha_h ? (difference > 0 ? 1.0f : -1.0f) : (difference / alpha_h)) / 5.0f;
		r.rgb_ray[k] = rgb_ray[k];
	}
	if (weight_sum >= 1.0 - 1e-4) { weight_sum = (float)(1.0 - 1e-4); r.gradient_weight_sum = 0.0f; }
	else if (weight_sum <= 1e-4) { weight_sum = (float)1e-4; r.gradient_weight_sum = 0.0f; }
	else {
		const float sigmoid_weight_sum = 1.0f / (1.0f + expf(-weight_sum));
		r.gradient_weight_sum = (t.mask_gt - sigmoid_weight_sum) * weight_sum * lp.mask_loss_weight;
	}
	r.weight_sum = weight_sum;
	r.mean_loss = (lg_loss[0] + lg_loss[1] + lg_loss[2]) / 3.0f;
	r.mask_loss = -(t.mask_gt * logf(weight_sum) + (1 - t.mask_gt) * logf(1 - weight_sum));
	return r;
}

// dL/d(network output) of ONE sample (:1800-1957) from its evaluation `e`, its compositing weight, the transmittance right behind it (Tj)
// and the colour accumulated up to and including it (my_r2); `o` is the sample's raw output row (the three colour logits are read again).
// Writes the 16-half row; returns the sample's eikonal term (|grad sdf| - 1)^2.
__device__ __forceinline__ float sample_gradient(const SampleEval& e, const __half* __restrict__ o, float weight, float Tj, const float my_r2[3], const RayLoss& rl,
                                                 const float dir[3], const LossParamsDev& lp, float loss_scale, __half* __restrict__ dout_row) {
	float suffix[3], dloss_by_drgb[3];
	#pragma unroll
	for (int k = 0; k < 3; ++k) { suffix[k] = rl.rgb_ray[k] - my_r2[k]; dloss_by_drgb[k] = weight * rl.lg_grad[k]; }
	float local[16];
	#pragma unroll
	for (int k = 0; k < 16; ++k) local[k] = 0.f;
	#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const float ov = __half2float(o[k]);
		const float density = logistic_f(ov);
		const float deriv = density * (1 - density);
		local[k] = loss_scale * (dloss_by_drgb[k] * deriv + fmaxf(0.0f, 0.0f * ov));
	}
	const float dot = rl.lg_grad[0] * (Tj * e.rgb[0] - suffix[0]) + rl.lg_grad[1] * (Tj * e.rgb[1] - suffix[1]) + rl.lg_grad[2] * (Tj * e.rgb[2] - suffix[2]);
	const float dloss_dalpha = (float)((dot + rl.gradient_weight_sum * (1 - rl.weight_sum)) / (1.0f - e.alpha + 1e-5));
	float dalpha_d_e_minus_sigmoidx, d_e_minus_sigmoidx_dsdf, d_e_minus_sigmoidx_dinvs, d_alpha_d_plus_e, d_plus_e_dinvs, d_plus_e_iter_cos, d_e_minus_sigmoidx_diter_cos;
	if (e.p_div_c <= 0.0f || e.p_div_c >= 1.0f) {
		dalpha_d_e_minus_sigmoidx = d_e_minus_sigmoidx_dsdf = d_e_minus_sigmoidx_dinvs = d_alpha_d_plus_e = d_plus_e_dinvs = d_plus_e_iter_cos = d_e_minus_sigmoidx_diter_cos = 0.0f;
	} else {
		const float plus_sigmoid_x = e.inv_s * e.iter_cos * e.dt;
		const float plus_e = expf(plus_sigmoid_x);
		const float e_minus_sigmoid = expf(-e.next_sdf * e.inv_s);
		d_e_minus_sigmoidx_dsdf = -e.inv_s * e_minus_sigmoid;
		d_e_minus_sigmoidx_dinvs = -e.next_sdf * e_minus_sigmoid;
		const float a = 1 + e_minus_sigmoid;
		const float b = 1 + plus_e * e_minus_sigmoid;
		const float c = (float)(1e-5 + 1 / (1 + plus_e * e_minus_sigmoid));
		const float delta = a * (b * b) * (c * c);
		dalpha_d_e_minus_sigmoidx = -(plus_e / (delta) - 1 / (a * a * c));
		d_alpha_d_plus_e = -e_minus_sigmoid / (delta);
		d_plus_e_dinvs = plus_e * e.iter_cos * e.dt;
		d_plus_e_iter_cos = plus_e * e.inv_s * e.dt;
		d_e_minus_sigmoidx_diter_cos = (float)(-e.inv_s * e_minus_sigmoid * e.dt * 0.5);
	}
	const float dloss_dinvs = dloss_dalpha * (dalpha_d_e_minus_sigmoidx * d_e_minus_sigmoidx_dinvs + d_alpha_d_plus_e * d_plus_e_dinvs);
	const float dloss_dvariance = dloss_dinvs * e.inv_s * 10;
	const float d_iter_cos_true_cos = (e.true_cos >= 0) ? 0.0f : 1.0f;
	const float gradient_norm = (float)sqrt((double)(e.g[0] * e.g[0] + e.g[1] * e.g[1] + e.g[2] * e.g[2]) + 1e-6);
	const float pos_gradient_norm_inv = 1 - 1 / gradient_norm;
	const float dloss_dnormal_norm = dloss_dalpha * (dalpha_d_e_minus_sigmoidx * d_e_minus_sigmoidx_diter_cos + d_plus_e_iter_cos * d_alpha_d_plus_e) * d_iter_cos_true_cos;
	const float dloss_dsdf = dloss_dalpha * dalpha_d_e_minus_sigmoidx * d_e_minus_sigmoidx_dsdf;
	local[3] = loss_scale * dloss_dsdf;
	#pragma unroll
	for (int k = 0; k < 3; ++k) local[4 + k] = lp.ek_loss_weight * 2 * lp.loss_scale * pos_gradient_norm_inv * e.g[k];
	local[7] = loss_scale * dloss_dvariance;
	#pragma unroll
	for (int k = 0; k < 3; ++k) local[8 + k] = loss_scale * dloss_dnormal_norm * dir[k];
	uint32_t packed[8];
	#pragma unroll
	for (int k = 0; k < 8; ++k) packed[k] = pack_h2(local[2 * k], local[2 * k + 1]);
	reinterpret_cast<uint4*>(dout_row)[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
	reinterpret_cast<uint4*>(dout_row)[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
	return (gradient_norm - 1.0f) * (gradient_norm - 1.0f);
}

__device__ __forceinline__ SampleEval zero_eval() {
	SampleEval e;
	e.alpha = 0.f; e.rgb[0] = e.rgb[1] = e.rgb[2] = 0.f; e.p_div_c = 0.f; e.inv_s = 0.f; e.iter_cos = 0.f; e.dt = 0.f; e.next_sdf = 0.f; e.true_cos = 0.f; e.sdf = 0.f;
	e.g[0] = e.g[1] = e.g[2] = 0.f;
	return e;
}

// pass 2: loss + dL/doutput + compaction (:1629-1957), one WARP per ray (see k_loss_composite).
__global__ void __launch_bounds__(128)
k_loss_gradient(uint32_t n_rays, uint32_t n_rays_total, Pcg32 rng_in, DatasetDev ds, LossParamsDev lp, uint32_t max_samples_compacted,
                const uint32_t* __restrict__ n_kept_ptr, const __half* __restrict__ network_output, const uint32_t* __restrict__ ray_indices,
                uint32_t* __restrict__ numsteps, const float* __restrict__ coords, const float* __restrict__ ray_tmp,
                const uint32_t* __restrict__ compacted_base_arr, float* __restrict__ coords_compacted, __half* __restrict__ dloss_doutput,
                float* __restrict__ loss_output, float* __restrict__ ek_loss_output, float* __restrict__ mask_loss_output) {
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (i >= *n_kept_ptr) return;
	const uint32_t ns = numsteps[i * 2 + 0], base = numsteps[i * 2 + 1];
	const float* cin = coords + (size_t)base * 7;
	const __half* nout = network_output + (size_t)base * 16;
	const float* rt = ray_tmp + (size_t)i * 12;
	uint32_t compacted = __float_as_uint(rt[0]);
	float rgb_ray[3] = {rt[1], rt[2], rt[3]};
	float weight_sum = rt[4];
	float T = rt[5];
	const float dir[3] = {rt[6], rt[7], rt[8]};

	const RayTarget tg = ray_target(ray_indices[i], n_rays, n_rays_total, rng_in, ds, lp);
	if (compacted == ns) {
		#pragma unroll
		for (int k = 0; k < 3; ++k) rgb_ray[k] += T * tg.bg[k];
	}
	__syncwarp();   // every lane has read numsteps before lane 0 rewrites it
	const uint32_t compacted_base = compacted_base_arr[i];
	compacted = min(max_samples_compacted - min(max_samples_compacted, compacted_base), compacted);
	if (lane == 0) {
		numsteps[i * 2 + 0] = compacted;
		numsteps[i * 2 + 1] = compacted_base;
	}
	if (compacted == 0) return;

	float* cout = coords_compacted + (size_t)compacted_base * 7;
	__half* dout = dloss_doutput + (size_t)compacted_base * 16;
	const RayLoss rl = ray_loss(rgb_ray, weight_sum, tg, lp);
	if (lane == 0) {
		loss_output[i] = rl.mean_loss / (float)n_rays;
		mask_loss_output[i] = rl.mask_loss;
	}
	float ek_acc = 0.f;

	const float loss_scale = lp.loss_scale / n_rays;
	float rgb_ray2[3] = {0.f, 0.f, 0.f};
	T = 1.f;
	__shared__ float4 s_stage[4][32];
	float4* stage = s_stage[threadIdx.x >> 5];
	for (uint32_t c0 = 0; c0 < compacted; c0 += 32) {
		const uint32_t cnt = min(32u, compacted - c0);
		// compact the network inputs: cnt*7 contiguous floats, coalesced
		for (uint32_t k = lane; k < cnt * 7; k += 32) cout[(size_t)c0 * 7 + k] = cin[(size_t)c0 * 7 + k];
		const uint32_t j = c0 + lane;
		const bool valid = lane < (int)cnt;
		SampleEval e = zero_eval();
		const __half* o = nout + (size_t)j * 16;
		if (valid) e = eval_sample(o, cin + (size_t)j * 7, dir, lp.cos_anneal_ratio);
		// replay the recurrences in sample order; lane L keeps the state right after its own sample
		SampleCapture my = {0.f, 0.f, {0.f, 0.f, 0.f}};
		stage_chunk(stage, e, lane);
		replay_capture(stage, cnt, lane, T, rgb_ray2, my);
		if (!valid) continue;
		ek_acc += sample_gradient(e, o, my.weight, my.T, my.rgb, rl, dir, lp, loss_scale, dout + (size_t)j * 16);
	}
	// eikonal log value: warp reduction (the reference sums serially; fp32 order differs, it is a reported scalar only)
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) ek_acc += __shfl_xor_sync(0xffffffffu, ek_acc, o);
	if (lane == 0) ek_loss_output[i] = ek_acc / ((float)compacted * (float)n_rays);
}

// ---- the loss by SAMPLES instead of by rays --------------------------------------------------------------------------------------
// In steady state a step has ~2,500 kept rays of ~270 marched / ~75 kept samples each (up to 400 / 300: profiles/r02_ray_lengths.json): one
// warp per ray leaves most of the GPU idle and every kernel waits for its longest ray.  Only the compositing recurrence is serial along a
// ray; everything expensive (the logistic / exp / division chains of eval_sample and the gradient expressions) is per sample.  So:
//   * the first network pass, which evaluates eval_sample for its early-out anyway, stores {alpha, r, g, b} per sample (forward_rays.cu);
//   * k_loss_rays (one warp per ray) only REPLAYS the recurrences over those values (a 16-byte load per sample, ~10 cycles per sample on the
//     dependent chain), stores each kept sample's {weight, T behind it, colour up to it}, takes the ray's offset into the compacted batch
//     from a prefix sum across the blocks of the launch (lookback.cuh; ray-slot order, as k_scan_compacted), and leaves a 48-byte record
//     per ray plus, per compacted slot, which sample and which ray it belongs to;
//   * k_loss_samples (one THREAD per compacted sample, perfectly balanced) evaluates the sample again and writes its dL/doutput row and its
//     compacted coordinates.
// Every value is computed by the same expressions in the same order as in k_loss_composite / k_loss_gradient (the tests compare bit for bit).
struct RayRecord { float rgb_ray[3]; float lg_grad[3]; float gradient_weight_sum, weight_sum; float dir[3]; float pad; };
static_assert(sizeof(RayRecord) == 48, "three 16-byte loads per record");

__device__ __forceinline__ float eikonal_term(uint2 g_bits) {   // (|grad sdf| - 1)^2 from output columns 4..6, as in sample_gradient
	const __half* gh = reinterpret_cast<const __half*>(&g_bits);
	const float g0 = __half2float(gh[0]), g1 = __half2float(gh[1]), g2 = __half2float(gh[2]);
	const float gradient_norm = (float)sqrt((double)(g0 * g0 + g1 * g1 + g2 * g2) + 1e-6);
	return (gradient_norm - 1.0f) * (gradient_norm - 1.0f);
}

#define LF_RAYS 8
__global__ void __launch_bounds__(LF_RAYS * 32)
k_loss_rays(uint32_t n_rays, uint32_t n_rays_total, Pcg32 rng_in, DatasetDev ds, LossParamsDev lp, uint32_t max_samples_compacted,
            const uint32_t* __restrict__ n_kept_ptr, const __half* __restrict__ network_output, const float4* __restrict__ composite_in,
            const uint32_t* __restrict__ ray_indices, uint32_t* __restrict__ numsteps, LookbackScan scan, LossScratch ws,
            float* __restrict__ loss_output, float* __restrict__ ek_loss_output, float* __restrict__ mask_loss_output, uint32_t* __restrict__ out_counter) {
	__shared__ uint32_t s_tile, s_count[LF_RAYS], s_base[LF_RAYS];
	__shared__ float4 s_stage[LF_RAYS][32];
	if (threadIdx.x == 0) s_tile = lookback_take_ticket(scan);
	__syncthreads();
	const uint32_t tile = s_tile;
	const uint32_t n_kept = *n_kept_ptr;
	const uint32_t n_tiles = (n_kept + LF_RAYS - 1) / LF_RAYS;
	if (tile >= n_tiles) {
		if (n_kept == 0 && tile == 0 && threadIdx.x == 0) *out_counter = 0u;
		return;
	}
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	float4* stage = s_stage[warp];
	const uint32_t i = tile * LF_RAYS + warp;
	const bool active = i < n_kept;

	uint32_t ns = 0, base = 0;
	CompositeState cs = {1.f, 0.f, {0.f, 0.f, 0.f}, 0u, false};
	float ek_lane = 0.f;
	float dir[3] = {0.f, 0.f, 0.f};
	RayTarget tg;
	if (active) {
		ns = numsteps[i * 2 + 0]; base = numsteps[i * 2 + 1];
		const __half* nout = network_output + (size_t)base * 16;
		const float4* comp = composite_in + base;
		tg = ray_target(ray_indices[i], n_rays, n_rays_total, rng_in, ds, lp);   // independent of the network: its loads overlap the replay
		if (ns > 0) {   // BENT_DIR :1583-1588: the view direction comes from sample 0's output (a kept ray always has one)
			const float dv[3] = {__half2float(nout[8]) * 2.0f - 1.0f, __half2float(nout[9]) * 2.0f - 1.0f, __half2float(nout[10]) * 2.0f - 1.0f};
			normalize3(dv, dir);
		}
		const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
		float4 v = lane < (int)ns ? comp[lane] : zero4;                                                                   // chunk 0
		uint2 gq = lane < (int)ns ? *reinterpret_cast<const uint2*>(nout + (size_t)lane * 16 + 4) : make_uint2(0u, 0u);  // output columns 4..7
		for (uint32_t c0 = 0; c0 < ns && !cs.stop; c0 += 32) {
			const uint32_t idx = c0 + lane;
			const float4 v_cur = v;
			const uint2 g_cur = gq;
			if (idx + 32 < ns) {   // the next chunk's loads are in flight during this chunk's replay
				v = comp[idx + 32];
				gq = *reinterpret_cast<const uint2*>(nout + (size_t)(idx + 32) * 16 + 4);
			} else { v = zero4; gq = make_uint2(0u, 0u); }
			SampleCapture cap = {0.f, 0.f, {0.f, 0.f, 0.f}};
			const uint32_t consumed = composite_chunk(stage, v_cur.x, v_cur.y, v_cur.z, v_cur.w, min(32u, ns - c0), lane, cs, cap);
			if ((uint32_t)lane < consumed) {   // this lane's sample was consumed: what pass 2 needs from the recurrences
				ws.prefix4[(size_t)base + idx] = make_float4(cap.weight, cap.T, cap.rgb[0], cap.rgb[1]);
				ws.prefix1[(size_t)base + idx] = cap.rgb[2];
				ek_lane += eikonal_term(g_cur);
			}
		}
	}
	uint32_t compacted = cs.compacted;
	float T = cs.T, weight_sum = cs.weight_sum;
	float rgb_ray[3] = {cs.rgb[0], cs.rgb[1], cs.rgb[2]};

	// ---- where the ray's samples go in the compacted batch ----
	if (lane == 0) s_count[warp] = compacted;
	__syncthreads();
	if (warp == 0) {
		const uint32_t c = lane < LF_RAYS ? s_count[lane] : 0u;
		uint32_t inc = c;
		#pragma unroll
		for (int o = 1; o < LF_RAYS; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
		const uint32_t aggregate = __shfl_sync(0xffffffffu, inc, LF_RAYS - 1);
		const uint32_t before = lookback_exclusive(scan, tile, aggregate, lane);
		if (lane < LF_RAYS) s_base[lane] = before + inc - c;
		if (tile == n_tiles - 1 && lane == 0) *out_counter = before + aggregate;   // total compacted samples (un-clamped), as k_scan_compacted
	}
	__syncthreads();
	if (!active) return;

	// ---- per-ray part of pass 2 (:1629-1760) ----
	if (compacted == ns) {
		#pragma unroll
		for (int k = 0; k < 3; ++k) rgb_ray[k] += T * tg.bg[k];
	}
	const uint32_t compacted_base = s_base[warp];
	const uint32_t unclamped = compacted;
	compacted = min(max_samples_compacted - min(max_samples_compacted, compacted_base), compacted);
	if (lane == 0) {
		numsteps[i * 2 + 0] = compacted;
		numsteps[i * 2 + 1] = compacted_base;
	}
	if (compacted == 0) return;
	const RayLoss rl = ray_loss(rgb_ray, weight_sum, tg, lp);
	if (lane == 0) {
		loss_output[i] = rl.mean_loss / (float)n_rays;
		mask_loss_output[i] = rl.mask_loss;
		RayRecord rec;
		#pragma unroll
		for (int k = 0; k < 3; ++k) { rec.rgb_ray[k] = rl.rgb_ray[k]; rec.lg_grad[k] = rl.lg_grad[k]; rec.dir[k] = dir[k]; }
		rec.gradient_weight_sum = rl.gradient_weight_sum; rec.weight_sum = rl.weight_sum; rec.pad = 0.f;
		const float4* src = reinterpret_cast<const float4*>(&rec);
		float4* dst = reinterpret_cast<float4*>(ws.ray_rec) + (size_t)i * 3;
		dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
	}
	if (compacted != unclamped) {   // the ray that straddles the end of the batch: its eikonal sum covers the samples that fit only
		const __half* nout = network_output + (size_t)base * 16;
		ek_lane = 0.f;
		for (uint32_t c0 = 0; c0 < compacted; c0 += 32)
			if (c0 + lane < compacted) ek_lane += eikonal_term(*reinterpret_cast<const uint2*>(nout + (size_t)(c0 + lane) * 16 + 4));
	}
	for (uint32_t c0 = 0; c0 < compacted; c0 += 32) {   // compacted slot -> (sample, ray)
		const uint32_t j = c0 + lane;
		if (j < compacted) { ws.slot_sample[compacted_base + j] = base + j; ws.slot_ray[compacted_base + j] = i; }
	}
	// eikonal log value: warp reduction (the reference sums serially; fp32 order differs, it is a reported scalar only)
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) ek_lane += __shfl_xor_sync(0xffffffffu, ek_lane, o);
	if (lane == 0) ek_loss_output[i] = ek_lane / ((float)compacted * (float)n_rays);
}

__global__ void __launch_bounds__(256)
k_loss_samples(uint32_t n_rays, LossParamsDev lp, uint32_t max_samples_compacted, const uint32_t* __restrict__ total_ptr,
               const __half* __restrict__ network_output, const float* __restrict__ coords, LossScratch ws,
               float* __restrict__ coords_compacted, __half* __restrict__ dloss_doutput) {
	const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= min(*total_ptr, max_samples_compacted)) return;
	const uint32_t s = ws.slot_sample[m], i = ws.slot_ray[m];
	RayRecord rec;
	{
		const float4* src = reinterpret_cast<const float4*>(ws.ray_rec) + (size_t)i * 3;
		float4* dst = reinterpret_cast<float4*>(&rec);
		dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
	}
	RayLoss rl;
	#pragma unroll
	for (int k = 0; k < 3; ++k) { rl.rgb_ray[k] = rec.rgb_ray[k]; rl.lg_grad[k] = rec.lg_grad[k]; }
	rl.gradient_weight_sum = rec.gradient_weight_sum; rl.weight_sum = rec.weight_sum; rl.mean_loss = 0.f; rl.mask_loss = 0.f;
	const __half* o = network_output + (size_t)s * 16;
	const float* cin = coords + (size_t)s * 7;
	const SampleEval e = eval_sample(o, cin, rec.dir, lp.cos_anneal_ratio);
	const float4 p4 = ws.prefix4[s];
	const float my_r2[3] = {p4.z, p4.w, ws.prefix1[s]};
	const float loss_scale = lp.loss_scale / n_rays;
	(void)sample_gradient(e, o, p4.x, p4.y, my_r2, rl, rec.dir, lp, loss_scale, dloss_doutput + (size_t)m * 16);
	float* cout = coords_compacted + (size_t)m * 7;
	#pragma unroll
	for (int k = 0; k < 7; ++k) cout[k] = cin[k];
}

int launch_compute_loss(cudaStream_t stream, uint32_t n_rays_local, uint32_t n_rays, uint32_t n_rays_total, const uint64_t rng_state[2], const DatasetDev& ds, const Aabb& bb,
                        const GlobalMove& gmv, const LossParamsDev& lp, uint32_t max_samples_compacted, const uint32_t* n_kept_ptr,
                        const __half* network_output, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords,
                        float* ray_tmp, uint32_t* scan_tmp, float* coords_compacted, __half* dloss_doutput,
                        float* loss_output, float* ek_loss_output, float* mask_loss_output, uint32_t* out_counter,
                        const float4* composite_in, const LossScratch* scratch, unsigned long long* lookback_desc, uint32_t* lookback_ticket, uint32_t lookback_epoch) {
	(void)bb;
	if (n_rays_local == 0) return 0;
	Pcg32 rng; rng.state = rng_state[0]; rng.inc = rng_state[1];
	if (composite_in && scratch && lookback_desc) {
		LookbackScan scan = {lookback_desc, lookback_ticket, lookback_epoch};
		k_loss_rays<<<(n_rays_local + LF_RAYS - 1) / LF_RAYS, LF_RAYS * 32, 0, stream>>>(n_rays, n_rays_total, rng, ds, lp, max_samples_compacted, n_kept_ptr, network_output, composite_in,
		        ray_indices, numsteps, scan, *scratch, loss_output, ek_loss_output, mask_loss_output, out_counter);
		COUNT_LAUNCH(1);
		CUDA_CHECK(cudaGetLastError());
		k_loss_samples<<<(max_samples_compacted + 255) / 256, 256, 0, stream>>>(n_rays, lp, max_samples_compacted, out_counter, network_output, coords, *scratch, coords_compacted, dloss_doutput);
		COUNT_LAUNCH(1);
		CUDA_CHECK(cudaGetLastError());
		return 0;
	}
	const uint32_t blocks = (n_rays_local + 3) / 4;   // one warp per ray
	uint32_t* counts = scan_tmp;
	uint32_t* cbase = scan_tmp + n_rays_local;
	k_loss_composite<<<blocks, 128, 0, stream>>>(n_kept_ptr, gmv, lp.cos_anneal_ratio, network_output, rays, numsteps, coords, ray_tmp, counts);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	k_scan_compacted<<<1, SCAN_THREADS, 0, stream>>>(n_kept_ptr, counts, cbase, out_counter);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	k_loss_gradient<<<blocks, 128, 0, stream>>>(n_rays, n_rays_total, rng, ds, lp, max_samples_compacted, n_kept_ptr, network_output, ray_indices,
	                                           numsteps, coords, ray_tmp, cbase, coords_compacted, dloss_doutput, loss_output, ek_loss_output, mask_loss_output);
	COUNT_LAUNCH(1);
	CUDA_CHECK(cudaGetLastError());
	return 0;
}
