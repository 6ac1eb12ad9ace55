// TEST INFRASTRUCTURE ONLY.  Compiles the REFERENCE's device helper headers
// (/root/reference/include/neural-graphics-primitives/common_device.cuh: srgb transfer functions, image_pos, read_rgba, pixel_to_ray,
// fov conversions; bounding_box.cuh: BoundingBox::ray_intersect / contains) as host code against the stand-in headers of
// oracle/ref_host_shim/ and exposes them through a C ABI (oracle/Makefile.ref -> oracle/_ref/libref_ngp_host.so).
#include <cassert>
#include <sstream>
#include <string>
#include <vector>
#include <neural-graphics-primitives/bounding_box.cuh>

RefHostIdx threadIdx, blockIdx;
RefHostDim blockDim, gridDim;

extern "C" {
float refn_srgb_to_linear(float x) { return ngp::srgb_to_linear(x); }
float refn_linear_to_srgb(float x) { return ngp::linear_to_srgb(x); }
float refn_fov_to_focal_length(int resolution, float degrees) { return ngp::fov_to_focal_length(resolution, degrees); }
float refn_focal_length_to_fov(int resolution, float focal) { return ngp::focal_length_to_fov(resolution, focal); }

void refn_ray_intersect(const float* aabb6, const float* o, const float* d, float* out2) {
	ngp::BoundingBox bb({aabb6[0], aabb6[1], aabb6[2]}, {aabb6[3], aabb6[4], aabb6[5]});
	const Eigen::Vector2f t = bb.ray_intersect({o[0], o[1], o[2]}, {d[0], d[1], d[2]});
	out2[0] = t.x(); out2[1] = t.y();
}
int refn_contains(const float* aabb6, const float* p) {
	ngp::BoundingBox bb({aabb6[0], aabb6[1], aabb6[2]}, {aabb6[3], aabb6[4], aabb6[5]});
	return bb.contains({p[0], p[1], p[2]}) ? 1 : 0;
}
void refn_image_pos(const float* xy, int w, int h, int* out2) {
	const Eigen::Vector2i p = ngp::image_pos({xy[0], xy[1]}, {w, h});
	out2[0] = p.x(); out2[1] = p.y();
}
void refn_read_rgba(int px, int py, int w, int h, const uint8_t* pixels, uint32_t img, float* out4) {
	const Eigen::Array4f v = ngp::read_rgba(Eigen::Vector2i{px, py}, Eigen::Vector2i{w, h}, pixels, ngp::EImageDataType::Byte, img);
	for (int k = 0; k < 4; ++k) out4[k] = v[k];
}
// camera12: 3x4 row-major camera-to-world.  parallax_shift = {0, 0, z}: no head offset, as the NeuS paths call it.
void refn_pixel_to_ray(uint32_t spp, int px, int py, int w, int h, const float* focal2, const float* camera12, const float* screen_center2, int snap, float* out6) {
	Eigen::Matrix<float, 3, 4> cam;
	for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) cam(r, c) = camera12[r * 4 + c];
	const ngp::Ray ray = ngp::pixel_to_ray(spp, {px, py}, {w, h}, {focal2[0], focal2[1]}, cam, {screen_center2[0], screen_center2[1]}, {0.f, 0.f, 1.f}, snap != 0);
	for (int k = 0; k < 3; ++k) { out6[k] = ray.o[k]; out6[3 + k] = ray.d[k]; }
}
}
