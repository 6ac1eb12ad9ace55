// TEST INFRASTRUCTURE ONLY.  Compiles the REFERENCE's low-discrepancy / random helpers
// (/root/reference/include/neural-graphics-primitives/random_val.cuh: sobol, nested_uniform_scramble_base2, ld_random_val(_2d),
// ld_random_pixel_offset, cylindrical_to_dir, random_val* on the reference's own pcg32.h) as host code against the stand-in headers of
// oracle/ref_host_shim/ and exposes them through a C ABI (oracle/Makefile.ref -> oracle/_ref/libref_random_host.so).
#include <string>
#include <sstream>
#include <neural-graphics-primitives/random_val.cuh>

RefHostIdx threadIdx, blockIdx;
RefHostDim blockDim, gridDim;

extern "C" {
uint32_t refr_sobol(uint32_t index, uint32_t dim) { return ngp::sobol(index, dim); }
uint32_t refr_scramble(uint32_t x, uint32_t seed) { return ngp::nested_uniform_scramble_base2(x, seed); }
float refr_ld_random_val(uint32_t index, uint32_t seed, uint32_t dim) { return ngp::ld_random_val(index, seed, dim); }
void refr_ld_random_val_2d(uint32_t index, uint32_t seed, float* out2) { auto v = ngp::ld_random_val_2d(index, seed); out2[0] = v.x(); out2[1] = v.y(); }
void refr_ld_random_pixel_offset(uint32_t spp, float* out2) { auto v = ngp::ld_random_pixel_offset(spp); out2[0] = v.x(); out2[1] = v.y(); }
// the reference's pcg32 driven through random_val / random_val_2d / random_val_3d after an advance (as the kernels do: rng.advance(i * N))
void refr_pcg32_stream(uint64_t seed, int64_t advance, uint32_t n, float* floats, uint32_t* uints) {
	ngp::default_rng_t rng{seed};
	rng.advance(advance);
	ngp::default_rng_t rng2 = rng;
	for (uint32_t i = 0; i < n; ++i) floats[i] = ngp::random_val(rng);
	for (uint32_t i = 0; i < n; ++i) uints[i] = ngp::random_uint(rng2);
}
}
