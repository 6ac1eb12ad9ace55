// l2_microbench.cu -- what the B200's L2 sustains for the access patterns of the hash-grid kernels, as a roofline
// denominator for them: random 4-byte gathers over a 21 MB fp16 table (the forward gather) and random
// 8-byte vector reductions over a 42 MB fp32 table (the gradient scatter).  Both tables fit the 126 MB L2, so HBM is not
// the bound; the kernels below issue the accesses with as much memory-level parallelism as the hardware takes and nothing
// else, so their rate is the ceiling a real kernel can approach.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2_microbench tools/l2_microbench.cu && tools/l2_microbench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int ILP>
__global__ void k_gather(const uint32_t* __restrict__ table, uint32_t mask, uint32_t iters, uint32_t* __restrict__ sink) {
	uint32_t acc = 0, s = hash32(blockIdx.x * blockDim.x + threadIdx.x);
	for (uint32_t it = 0; it < iters; ++it) {
		uint32_t v[ILP];
		#pragma unroll
		for (int k = 0; k < ILP; ++k) { s = s * 1664525u + 1013904223u; v[k] = __ldg(table + (hash32(s) & mask)); }
		#pragma unroll
		for (int k = 0; k < ILP; ++k) acc ^= v[k];
	}
	if (acc == 0x12345678u) sink[0] = acc;
}

// 8 corners of a random cell: the hash-grid pattern (2 x 2 x 2 neighbourhood through the spatial hash)
__global__ void k_gather_corners(const uint32_t* __restrict__ table, uint32_t mask, uint32_t iters, uint32_t* __restrict__ sink) {
	uint32_t acc = 0, s = hash32(blockIdx.x * blockDim.x + threadIdx.x);
	for (uint32_t it = 0; it < iters; ++it) {
		s = s * 1664525u + 1013904223u;
		const uint32_t h = hash32(s), x = h & 1023u, y = (h >> 10) & 1023u, z = (h >> 20) & 1023u;
		uint32_t v[8];
		#pragma unroll
		for (int c = 0; c < 8; ++c) v[c] = __ldg(table + (((x + (c & 1)) ^ ((y + ((c >> 1) & 1)) * 2654435761u) ^ ((z + (c >> 2)) * 805459861u)) & mask));
		#pragma unroll
		for (int c = 0; c < 8; ++c) acc ^= v[c];
	}
	if (acc == 0x12345678u) sink[0] = acc;
}

__global__ void k_red(float2* __restrict__ table, uint32_t mask, uint32_t iters) {
	uint32_t s = hash32(blockIdx.x * blockDim.x + threadIdx.x);
	for (uint32_t it = 0; it < iters; ++it) {
		#pragma unroll
		for (int k = 0; k < 8; ++k) { s = s * 1664525u + 1013904223u; atomicAdd(table + (hash32(s) & mask), make_float2(1.f, 1.f)); }
	}
}

template <typename F> static float time_ms(F f, int reps = 5) {
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	f(); cudaDeviceSynchronize();
	float best = 1e30f;
	for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
	return best;
}

int main() {
	int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	const uint32_t n_entries = 1u << 22;                 // 4 Mi x 4 B = 16 MiB (closest power of two to the 21 MB of fp16 tables)
	uint32_t* table; float2* grads; uint32_t* sink;
	cudaMalloc(&table, (size_t)n_entries * 4); cudaMemset(table, 1, (size_t)n_entries * 4);
	cudaMalloc(&grads, (size_t)n_entries * 8); cudaMemset(grads, 0, (size_t)n_entries * 8);   // 32 MiB of float2
	cudaMalloc(&sink, 4);
	const int blocks = sms * 8, threads = 256;
	const uint32_t iters = 256;
	printf("{\"sms\": %d", sms);
	{
		float ms = time_ms([&] { k_gather<8><<<blocks, threads>>>(table, n_entries - 1, iters, sink); });
		double n = (double)blocks * threads * iters * 8;
		printf(", \"gather4B_random_Gops\": %.2f, \"gather4B_random_sector_GBps\": %.1f", n / ms / 1e6, n * 32 / ms / 1e6);
	}
	{
		float ms = time_ms([&] { k_gather<16><<<blocks, threads>>>(table, n_entries - 1, iters, sink); });
		double n = (double)blocks * threads * iters * 16;
		printf(", \"gather4B_random_ilp16_Gops\": %.2f", n / ms / 1e6);
	}
	{
		float ms = time_ms([&] { k_gather_corners<<<blocks, threads>>>(table, n_entries - 1, iters, sink); });
		double n = (double)blocks * threads * iters * 8;
		printf(", \"gather4B_hash_corners_Gops\": %.2f", n / ms / 1e6);
	}
	{
		float ms = time_ms([&] { k_red<<<blocks, threads>>>(grads, n_entries - 1, iters / 4); });
		double n = (double)blocks * threads * (iters / 4) * 8;
		printf(", \"red_v2f32_random_Gops\": %.2f", n / ms / 1e6);
	}
	printf("}\n");
	return 0;
}
