// TEST INFRASTRUCTURE ONLY.  Compiles the REFERENCE's rotation-6d / DeltaNetwork kernels
// (/root/reference/include/neural-graphics-primitives/common_operation.cuh:37-160, 416-470, 551-623, 788-845) as host code against the
// stand-in headers of oracle/ref_host_shim/ and exposes them through a C ABI, one simulated thread at a time.  Built by
// oracle/Makefile.ref into oracle/_ref/libref_delta_host.so where the reference tree is mounted; tests/test_ref_delta_host.py checks the
// oracle's restatement against it and tests/golden/make_delta_golden.py freezes its outputs into tests/golden/delta_reference.npz.
#include <neural-graphics-primitives/common_operation.cuh>

RefHostIdx threadIdx, blockIdx;
RefHostDim blockDim, gridDim;

static inline __half from_bits(uint16_t b) { __half h; std::memcpy(&h.h, &b, 2); return h; }
static inline uint16_t to_bits(const __half& h) { uint16_t b; std::memcpy(&b, &h.h, 2); return b; }

extern "C" {

// rotation_6d_to_matrix<__half>: R row-major out
void refd_rotation_matrix(const uint16_t* rot6_bits, float* R9) {
	__half r[8];
	for (int k = 0; k < 6; ++k) r[k] = from_bits(rot6_bits[k]);
	Eigen::Matrix3f m;
	rotation_6d_to_matrix<__half>(r, m);
	for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R9[i * 3 + j] = m(i, j);
}

// add_global_movement_with_rotation_6d<float, __half> on [n][6] = (pos3, warped dir3), as DeltaNetwork::forward_impl launches it
void refd_apply(uint32_t n, const float* in6, const uint16_t* rot6_bits, const uint16_t* trans_bits, const float* off3, float* out6) {
	__half r[8], t[4];
	for (int k = 0; k < 6; ++k) r[k] = from_bits(rot6_bits[k]);
	for (int k = 0; k < 3; ++k) t[k] = from_bits(trans_bits[k]);
	const Eigen::Vector3f off(off3 ? off3[0] : 0.f, off3 ? off3[1] : 0.f, off3 ? off3[2] : 0.f);
	tcnn::MatrixView<float> vin(const_cast<float*>(in6), 1, 6), vout(out6, 1, 6);
	for (uint32_t i = 0; i < n; ++i) { threadIdx.x = i; add_global_movement_with_rotation_6d<float, __half>(n, r, t, off, vin, vout); }
	threadIdx.x = 0;
}

// add_loss_to_rotation_6d_each<__half, float>: per-sample gradients as half bit patterns: rot [n][6], trans [n][3], input [n][3]
void refd_backward_each(uint32_t n, const float* dL3, const float* pos3, const uint16_t* rot6_bits, const uint16_t* trans_bits, const float* off3,
                        uint16_t* rot_grad, uint16_t* trans_grad, uint16_t* input_grad) {
	__half r[8], t[4];
	for (int k = 0; k < 6; ++k) r[k] = from_bits(rot6_bits[k]);
	for (int k = 0; k < 3; ++k) t[k] = from_bits(trans_bits[k]);
	const Eigen::Vector3f off(off3 ? off3[0] : 0.f, off3 ? off3[1] : 0.f, off3 ? off3[2] : 0.f);
	__half* rg = new __half[(size_t)n * 6]; __half* tg = new __half[(size_t)n * 3]; __half* ig = new __half[(size_t)n * 3];
	tcnn::MatrixView<float> loss(const_cast<float*>(dL3), 1, 3), input(const_cast<float*>(pos3), 1, 3);
	tcnn::MatrixView<__half> vrg(rg, 1, 6), vtg(tg, 1, 3), vig(ig, 1, 3);
	for (uint32_t i = 0; i < n; ++i) { threadIdx.x = i; add_loss_to_rotation_6d_each<__half, float>(n, loss, input, r, t, off, vrg, vtg, vig); }
	threadIdx.x = 0;
	for (size_t k = 0; k < (size_t)n * 6; ++k) rot_grad[k] = to_bits(rg[k]);
	for (size_t k = 0; k < (size_t)n * 3; ++k) { trans_grad[k] = to_bits(tg[k]); input_grad[k] = to_bits(ig[k]); }
	delete[] rg; delete[] tg; delete[] ig;
}

// accumulate_global_movement_rotation_6d_kernel<__half> (in place on acc_rot[12] / acc_trans[4]) and the read-only
// save_global_movement_rotation_6d_kernel<__half>
void refd_accumulate(const uint16_t* rot6_bits, const uint16_t* trans_bits, uint16_t* acc_rot_bits, uint16_t* acc_trans_bits, int save_only, uint16_t* saved_rot_bits, uint16_t* saved_trans_bits) {
	__half r[8], t[4], ar[12], at[4], sr[12], st[4];
	for (int k = 0; k < 6; ++k) r[k] = from_bits(rot6_bits[k]);
	for (int k = 0; k < 3; ++k) t[k] = from_bits(trans_bits[k]);
	for (int k = 0; k < 9; ++k) ar[k] = from_bits(acc_rot_bits[k]);
	for (int k = 0; k < 3; ++k) at[k] = from_bits(acc_trans_bits[k]);
	threadIdx.x = 0;
	if (save_only) {
		save_global_movement_rotation_6d_kernel<__half>(1u, r, t, ar, at, sr, st);
		for (int k = 0; k < 9; ++k) saved_rot_bits[k] = to_bits(sr[k]);
		for (int k = 0; k < 3; ++k) saved_trans_bits[k] = to_bits(st[k]);
	} else {
		accumulate_global_movement_rotation_6d_kernel<__half>(1u, r, t, ar, at);
		for (int k = 0; k < 9; ++k) acc_rot_bits[k] = to_bits(ar[k]);
		for (int k = 0; k < 3; ++k) acc_trans_bits[k] = to_bits(at[k]);
	}
}

}  // extern "C"
