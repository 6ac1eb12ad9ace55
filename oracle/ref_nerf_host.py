"""TEST INFRASTRUCTURE ONLY: ctypes access to oracle/_ref/libref_nerf_host.so -- the kernels of the REFERENCE's src/testbed_nerf.cu
(lines 1-2396) compiled as host code by oracle/Makefile.ref and run one simulated thread at a time, in ascending thread order.
Used by tests/golden/make_nerf_kernels_golden.py and tests/test_golden_nerf_kernels.py."""
import ctypes as C
import os

import numpy as np

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_nerf_host.so")
_lib = None


def available():
    return os.path.exists(LIB) and os.path.isdir("/root/reference/src")


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


IDENTITY_ROT_H = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float16).view(np.uint16)
ZERO_TRANS_H = np.zeros(4, dtype=np.float16).view(np.uint16)
UNIT_AABB = np.array([0, 0, 0, 1, 1, 1], dtype=np.float32)


def generate_training_samples(n_rays, max_samples, n_rays_total, rng_state, data, bitfield, snap=True, cone_angle=0.0,
                              rotation=IDENTITY_ROT_H, transition=ZERO_TRANS_H, first_frame_offset=None, aabb=UNIT_AABB):
    px, focal, pp, xf = _c(data["pixels"], np.uint8), _c(data["focal"], np.float32), _c(data["principal"], np.float32), _c(data["xforms"], np.float32)
    n_img, h, w, _ = px.shape
    off = np.zeros(3, np.float32) if first_frame_offset is None else _c(first_frame_offset, np.float32)
    cnt = np.zeros(2, np.uint32); ri = np.zeros(n_rays, np.uint32); rays = np.zeros((n_rays, 6), np.float32)
    ns = np.zeros((n_rays, 2), np.uint32); coords = np.zeros((max_samples, 7), np.float32)
    lib().refk_generate_training_samples(C.c_uint32(n_rays), _p(_c(aabb, np.float32)), C.c_uint32(max_samples), C.c_uint32(n_rays_total),
                                         C.c_uint64(int(rng_state[0])), C.c_uint64(int(rng_state[1])), C.c_uint32(n_img), w, h, _p(px), _p(focal), _p(pp), _p(xf),
                                         _p(_c(bitfield, np.uint8)), 1 if snap else 0, C.c_float(cone_angle), _p(_c(rotation, np.uint16)), _p(_c(transition, np.uint16)),
                                         _p(off), _p(cnt), _p(ri), _p(rays), _p(ns), _p(coords))
    return dict(n_kept=int(cnt[0]), n_samples=int(cnt[1]), ray_indices=ri, rays=rays, numsteps=ns, coords=coords)


def compute_loss(n_rays, n_kept, n_rays_total, rng_state, max_samples_compacted, data, network_output_bits, ray_indices, rays, numsteps, coords,
                 loss_scale=128.0, background_color=(0, 0, 0), random_bg_color=True, snap=True, mask_loss_weight=0.0, ek_loss_weight=0.01,
                 cos_anneal_ratio=1.0, variance=0.3, rotation=IDENTITY_ROT_H, transition=ZERO_TRANS_H, first_frame_offset=None, aabb=UNIT_AABB):
    px, focal, pp, xf = _c(data["pixels"], np.uint8), _c(data["focal"], np.float32), _c(data["principal"], np.float32), _c(data["xforms"], np.float32)
    n_img, h, w, _ = px.shape
    off = np.zeros(3, np.float32) if first_frame_offset is None else _c(first_frame_offset, np.float32)
    ns = _c(numsteps, np.uint32).copy()
    B = max_samples_compacted
    cc = np.zeros((B, 7), np.float32); dl = np.zeros((B, 16), np.uint16)
    loss = np.zeros(n_rays, np.float32); ek = np.zeros(n_rays, np.float32); ml = np.zeros(n_rays, np.float32); comp = np.zeros(1, np.uint32)
    lib().refk_compute_loss(C.c_uint32(n_rays), C.c_uint32(n_kept), _p(_c(aabb, np.float32)), C.c_uint32(n_rays_total), C.c_uint64(int(rng_state[0])),
                            C.c_uint64(int(rng_state[1])), C.c_uint32(B), C.c_float(loss_scale), _p(_c(background_color, np.float32)), 1 if random_bg_color else 0,
                            1 if snap else 0, C.c_float(mask_loss_weight), C.c_float(ek_loss_weight), C.c_float(cos_anneal_ratio), C.c_float(variance),
                            C.c_uint32(n_img), w, h, _p(px), _p(focal), _p(pp), _p(xf), _p(_c(network_output_bits, np.uint16)), _p(_c(ray_indices, np.uint32)),
                            _p(_c(rays, np.float32)), _p(ns), _p(_c(coords, np.float32)), _p(cc), _p(dl), _p(loss), _p(ek), _p(ml),
                            _p(_c(rotation, np.uint16)), _p(_c(transition, np.uint16)), _p(off), _p(comp))
    return dict(numsteps=ns, coords_compacted=cc, dloss=dl, loss=loss, ek_loss=ek, mask_loss=ml, n_compacted=int(comp[0]))


def grid_samples(n, rng_state, step, grid_in, n_cascades, thresh, aabb=UNIT_AABB):
    pos = np.zeros((n, 3), np.float32); idx = np.zeros(n, np.uint32)
    lib().refk_grid_samples(C.c_uint32(n), C.c_uint64(int(rng_state[0])), C.c_uint64(int(rng_state[1])), C.c_uint32(step), _p(_c(aabb, np.float32)),
                            _p(_c(grid_in, np.float32)), C.c_uint32(n_cascades), C.c_float(thresh), _p(pos), _p(idx))
    return pos, idx


def splat_ema(indices, density_bits, decay, grid, ema_step=0):
    """in place on `grid` (float32 [n_elements])"""
    tmp = np.zeros(grid.shape[0], np.float32)
    lib().refk_splat_ema(C.c_uint32(len(indices)), _p(_c(indices, np.uint32)), _p(_c(density_bits, np.uint16)), C.c_uint32(grid.shape[0]), C.c_float(decay),
                         C.c_uint32(ema_step), _p(tmp), _p(grid))


def bitfield(grid, n_cascades_used, mean):
    out = np.zeros(128 ** 3 // 8 * 8, np.uint8)
    lib().refk_bitfield(_p(_c(grid, np.float32)), C.c_uint32(n_cascades_used), C.c_float(mean), _p(out))
    return out


def render_trace(net_forward, width, height, sample_index, focal, camera, screen_center, bitfield, snap=False, min_transmittance=0.01, variance=0.3,
                 cos_anneal_ratio=1.0, cone_angle=0.0, rotation=IDENTITY_ROT_H, transition=ZERO_TRANS_H, first_frame_offset=None,
                 render_aabb=UNIT_AABB, train_aabb=UNIT_AABB):
    """init_rays_with_payload_kernel_nerf + advance_pos_nerf + the trace loop (compact / generate_next_nerf_network_inputs / network /
    composite_kernel_nerf).  net_forward(n, coords_ptr, out_ptr) is called with ctypes pointers: coords float [n][7] -> out half [n][16]."""
    NETFN = C.CFUNCTYPE(None, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_uint16))
    cb = NETFN(net_forward)
    n = width * height
    off = np.zeros(3, np.float32) if first_frame_offset is None else _c(first_frame_offset, np.float32)
    pay = np.zeros((n, 8), np.float32); rgba = np.zeros((n, 4), np.float32); depth = np.zeros(n, np.float32); ns = np.zeros(n, np.uint32)
    f = lib().refk_render_trace
    f.restype = C.c_uint32
    hits = f(width, height, C.c_uint32(sample_index), _p(_c(focal, np.float32)), _p(_c(np.asarray(camera, np.float32).reshape(12), np.float32)),
             _p(_c(screen_center, np.float32)), 1 if snap else 0, _p(_c(render_aabb, np.float32)), _p(_c(train_aabb, np.float32)), _p(_c(rotation, np.uint16)),
             _p(_c(transition, np.uint16)), _p(off), _p(_c(bitfield, np.uint8)), C.c_float(cone_angle), C.c_float(min_transmittance), C.c_float(variance),
             C.c_float(cos_anneal_ratio), cb, _p(pay), _p(rgba), _p(depth), _p(ns))
    return dict(payload=pay, rgba=rgba, depth=depth, n_steps=ns, n_hit=int(hits))
