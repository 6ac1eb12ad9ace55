// TEST INFRASTRUCTURE ONLY.  Compiles the REFERENCE's dataset coordinate conventions (NerfDataset::nerf_matrix_to_ngp,
// ngp_matrix_to_nerf, nerf_direction_to_ngp, nerf_ray_to_ngp -- inline members of /root/reference/include/neural-graphics-primitives/
// nerf_loader.h:112-175) as host code against the stand-in headers of oracle/ref_host_shim/ (oracle/Makefile.ref ->
// oracle/_ref/libref_loader_host.so).  The real header is taken by absolute path: the shim directory holds a trimmed stand-in of the same
// name for the kernel builds.
#include <tiny-cuda-nn/common.h>
#include "/root/reference/include/neural-graphics-primitives/nerf_loader.h"

RefHostIdx threadIdx, blockIdx;
RefHostDim blockDim, gridDim;

extern "C" {
float refl_nerf_scale(void) { return ngp::NERF_SCALE; }
// m12: 3x4 row-major; mode 0 = default axis cycle, 1 = from_na, 2 = from_mitsuba; inverse != 0 -> ngp_matrix_to_nerf
void refl_matrix(const float* m12, float scale, const float* offset3, int mode, int inverse, float* out12) {
	ngp::NerfDataset ds;
	ds.scale = scale; ds.offset = Eigen::Vector3f{offset3[0], offset3[1], offset3[2]};
	ds.from_na = mode == 1; ds.from_mitsuba = mode == 2;
	Eigen::Matrix<float, 3, 4> m;
	for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) m(r, c) = m12[r * 4 + c];
	const Eigen::Matrix<float, 3, 4> o = inverse ? ds.ngp_matrix_to_nerf(m) : ds.nerf_matrix_to_ngp(m);
	for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) out12[r * 4 + c] = o(r, c);
}
void refl_direction(const float* d3, int from_mitsuba, float* out3) {
	ngp::NerfDataset ds; ds.from_mitsuba = from_mitsuba != 0;
	const Eigen::Vector3f o = ds.nerf_direction_to_ngp({d3[0], d3[1], d3[2]});
	for (int k = 0; k < 3; ++k) out3[k] = o[k];
}
}
