"""ctypes binding of the CPU oracle (oracle/neus_oracle.cpp).

TEST INFRASTRUCTURE ONLY -- see the header of neus_oracle.cpp.  Imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by the product package neus2_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)
h16p = C.POINTER(C.c_uint16)
u8p = C.POINTER(C.c_uint8)


def build(force=False):
    so = os.path.join(_HERE, "libneus_oracle.so")
    src = os.path.join(_HERE, "neus_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_per_level_scale.restype = C.c_float
        _LIB.orc_per_level_scale.argtypes = [C.c_float, C.c_int, C.c_uint32, C.c_uint32]
        _LIB.orc_pcg32_next_float.restype = C.c_float
        _LIB.orc_pcg32_next_uint.restype = C.c_uint32
        _LIB.orc_trainer_seed_word.restype = C.c_uint32
        _LIB.orc_grid_index.restype = C.c_uint32
        _LIB.orc_morton3D.restype = C.c_uint32
        _LIB.orc_morton3D_invert.restype = C.c_uint32
        _LIB.orc_net_n_params.restype = C.c_uint64
        _LIB.orc_update_bitfield.restype = C.c_float
        _LIB.orc_srgb_to_linear.restype = C.c_float
        _LIB.orc_srgb_to_linear.argtypes = [C.c_float]
        _LIB.orc_linear_to_srgb.restype = C.c_float
        _LIB.orc_linear_to_srgb.argtypes = [C.c_float]
    return _LIB


def _p(a, t):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be C-contiguous"
    return a.ctypes.data_as(t)


class NetCfg(C.Structure):
    _fields_ = [
        ("n_levels", C.c_uint32),
        ("offsets", C.c_uint32 * 33),
        ("resolutions", C.c_uint32 * 32),
        ("scales", C.c_float * 32),
        ("valid_level", C.c_uint32),
        ("sdf_hidden_layers", C.c_uint32),
        ("rgb_hidden_layers", C.c_uint32),
        ("sdf_bias", C.c_float),
    ]


class Dataset(C.Structure):
    _fields_ = [
        ("n_images", C.c_uint32),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("pixels", u8p),
        ("focal", f32p),
        ("principal", f32p),
        ("xforms", f32p),
    ]


class LossParams(C.Structure):
    _fields_ = [
        ("loss_scale", C.c_float),
        ("background_color", C.c_float * 3),
        ("random_bg_color", C.c_int),
        ("snap_to_pixel_centers", C.c_int),
        ("mask_loss_weight", C.c_float),
        ("ek_loss_weight", C.c_float),
        ("cos_anneal_ratio", C.c_float),
    ]


# ---------------------------------------------------------------------------
# RNG
# ---------------------------------------------------------------------------
class Pcg32:
    """tcnn pcg32 (state, inc) pair; all arithmetic happens in the C oracle."""

    def __init__(self, seed=None, state=None):
        self.st = np.zeros(2, dtype=np.uint64)
        if state is not None:
            self.st[:] = state
        else:
            lib().orc_pcg32_seed(C.c_uint64(seed), _p(self.st, u64p))

    def copy(self):
        return Pcg32(state=self.st.copy())

    def next_uint(self):
        return int(lib().orc_pcg32_next_uint(_p(self.st, u64p)))

    def next_float(self):
        return float(lib().orc_pcg32_next_float(_p(self.st, u64p)))

    def advance(self, delta=1 << 32):
        lib().orc_pcg32_advance(_p(self.st, u64p), C.c_int64(delta))


def trainer_seed_word(seed=1337):
    return int(lib().orc_trainer_seed_word(C.c_uint32(seed)))


def generate_random_uniform(rng, n, lower, upper):
    out = np.empty(n, dtype=np.float32)
    lib().orc_generate_random_uniform(_p(rng.st, u64p), C.c_uint64(n), _p(out, f32p), C.c_float(lower), C.c_float(upper))
    return out


def xavier_uniform(rng, fan_out, fan_in):
    out = np.empty(fan_out * fan_in, dtype=np.float32)
    lib().orc_xavier_uniform(_p(rng.st, u64p), C.c_uint32(fan_out), C.c_uint32(fan_in), _p(out, f32p))
    return out


# ---------------------------------------------------------------------------
# Grid / network configuration
# ---------------------------------------------------------------------------
def make_cfg(n_levels=14, base_resolution=16, top_resolution=2048.0, log2_hashmap_size=19, aabb_scale=1,
             valid_level=None, sdf_hidden_layers=1, rgb_hidden_layers=2, sdf_bias=-0.1, per_level_scale=None):
    L = lib()
    cfg = NetCfg()
    cfg.n_levels = n_levels
    if per_level_scale is None:
        per_level_scale = L.orc_per_level_scale(C.c_float(top_resolution), aabb_scale, base_resolution, n_levels)
    off = np.zeros(33, dtype=np.uint32)
    res = np.zeros(32, dtype=np.uint32)
    sc = np.zeros(32, dtype=np.float32)
    L.orc_grid_tables(C.c_uint32(n_levels), C.c_uint32(base_resolution), C.c_float(per_level_scale), C.c_uint32(log2_hashmap_size),
                      _p(off, u32p), _p(res, u32p), _p(sc, f32p))
    for i in range(33):
        cfg.offsets[i] = int(off[i])
    for i in range(32):
        cfg.resolutions[i] = int(res[i])
        cfg.scales[i] = float(sc[i])
    cfg.valid_level = n_levels if valid_level is None else valid_level
    cfg.sdf_hidden_layers = sdf_hidden_layers
    cfg.rgb_hidden_layers = rgb_hidden_layers
    cfg.sdf_bias = sdf_bias
    cfg.per_level_scale_py = float(per_level_scale)
    return cfg


def cfg_tables(cfg):
    L = cfg.n_levels
    return (np.array(cfg.offsets[: L + 1], dtype=np.uint32), np.array(cfg.resolutions[:L], dtype=np.uint32),
            np.array(cfg.scales[:L], dtype=np.float32))


def net_n_params(cfg):
    return int(lib().orc_net_n_params(C.byref(cfg)))


def param_layout(cfg):
    """Offsets (in elements) of [sdf-MLP, colour-MLP, grid, variance] in the reference order (nerf_network.h:741-785)."""
    sdf_in = (3 + 2 * cfg.n_levels + 15) // 16 * 16
    n_sdf = 64 * sdf_in + (cfg.sdf_hidden_layers - 1) * 64 * 64 + 16 * 64
    n_rgb = 64 * 48 + (cfg.rgb_hidden_layers - 1) * 64 * 64 + 16 * 64
    n_grid = int(cfg.offsets[cfg.n_levels]) * 2
    return dict(sdf=0, rgb=n_sdf, grid=n_sdf + n_rgb, variance=n_sdf + n_rgb + n_grid, total=n_sdf + n_rgb + n_grid + 4,
                sdf_in=sdf_in, n_sdf=n_sdf, n_rgb=n_rgb, n_grid=n_grid)


def grid_index(hashmap_size, res, x, y, z):
    return int(lib().orc_grid_index(C.c_uint32(hashmap_size), C.c_uint32(res), C.c_uint32(x), C.c_uint32(y), C.c_uint32(z)))


def grid_forward(cfg, pos, grid_h, want_dydx=True, stride=None):
    """pos: float32 [n, stride>=3]; grid_h: uint16 (half bits) grid params. Returns enc SoA [2L,n] (half bits), dy_dx [2L,n,3]."""
    n = pos.shape[0]
    stride = pos.shape[1] if stride is None else stride
    E = 2 * cfg.n_levels
    off, res, sc = cfg_tables(cfg)
    enc = np.zeros((E, n), dtype=np.uint16)
    dydx = np.zeros((E, n, 3), dtype=np.float32) if want_dydx else None
    lib().orc_grid_forward(C.c_uint32(n), _p(pos, f32p), C.c_uint32(stride), _p(grid_h, h16p), C.c_uint32(cfg.n_levels),
                           _p(off, u32p), _p(res, u32p), _p(sc, f32p), C.c_uint32(cfg.valid_level), _p(enc, h16p), _p(dydx, f32p))
    return enc, dydx


def grid_backward(cfg, pos, dL_dy_soa, grid_grad):
    n = pos.shape[0]
    off, res, sc = cfg_tables(cfg)
    lib().orc_grid_backward(C.c_uint32(n), _p(pos, f32p), C.c_uint32(pos.shape[1]), _p(dL_dy_soa, h16p), C.c_uint32(cfg.n_levels),
                            _p(off, u32p), _p(res, u32p), _p(sc, f32p), C.c_uint32(cfg.valid_level), _p(grid_grad, f32p))


def grid_bwdbwd_grid(cfg, pos, v, dL_dy_soa, grid_grad):
    n = pos.shape[0]
    off, res, sc = cfg_tables(cfg)
    lib().orc_grid_bwdbwd_grid(C.c_uint32(n), _p(pos, f32p), C.c_uint32(pos.shape[1]), _p(v, f32p), _p(dL_dy_soa, h16p),
                               C.c_uint32(cfg.n_levels), _p(off, u32p), _p(res, u32p), _p(sc, f32p), C.c_uint32(cfg.valid_level),
                               _p(grid_grad, f32p))


def grid_backward_input(cfg, dL_dy_soa, dy_dx):
    E, n = dL_dy_soa.shape
    out = np.zeros((n, 3), dtype=np.float32)
    lib().orc_grid_backward_input(C.c_uint32(n), C.c_uint32(E), _p(dL_dy_soa, h16p), _p(dy_dx, f32p), _p(out, f32p))
    return out


def grid_bwdbwd_dLdoutput(cfg, v, dy_dx):
    E, n = dy_dx.shape[:2]
    out = np.zeros((E, n), dtype=np.uint16)
    lib().orc_grid_bwdbwd_dLdoutput(C.c_uint32(n), C.c_uint32(E), _p(v, f32p), _p(dy_dx, f32p), _p(out, h16p))
    return out


# ---------------------------------------------------------------------------
# MLP
# ---------------------------------------------------------------------------
def mlp_forward(weights_h, x_h, in_w, n_hidden, out_w=16):
    n = x_h.shape[0]
    hidden = np.zeros((n_hidden, n, 64), dtype=np.uint16)
    out = np.zeros((n, out_w), dtype=np.uint16)
    lib().orc_mlp_forward(C.c_uint32(n), C.c_uint32(in_w), C.c_uint32(n_hidden), C.c_uint32(out_w), _p(weights_h, h16p),
                          _p(x_h, h16p), _p(hidden, h16p), _p(out, h16p))
    return hidden, out


def mlp_backward(weights_h, x_h, hidden, dL_dout_h, in_w, n_hidden, out_w=16, want_din=True, wgrad=None):
    n = x_h.shape[0]
    din = np.zeros((n, in_w), dtype=np.uint16) if want_din else None
    lib().orc_mlp_backward(C.c_uint32(n), C.c_uint32(in_w), C.c_uint32(n_hidden), C.c_uint32(out_w), _p(weights_h, h16p),
                           _p(x_h, h16p), _p(hidden, h16p), _p(dL_dout_h, h16p), _p(din, h16p), _p(wgrad, f32p), None)
    return din


def mlp_bwdbwd(weights_h, hidden, d_in_h, dL_dout_h, in_w, n_hidden, wgrad, out_w=16):
    n = d_in_h.shape[0]
    lib().orc_mlp_bwdbwd(C.c_uint32(n), C.c_uint32(in_w), C.c_uint32(n_hidden), C.c_uint32(out_w), _p(weights_h, h16p),
                         _p(hidden, h16p), _p(d_in_h, h16p), _p(dL_dout_h, h16p), _p(wgrad, f32p))


# ---------------------------------------------------------------------------
# Network
# ---------------------------------------------------------------------------
def net_forward(cfg, coords, params_h, debug=False):
    n = coords.shape[0]
    out = np.zeros((n, 16), dtype=np.uint16)
    if not debug:
        lib().orc_net_forward(C.byref(cfg), C.c_uint32(n), _p(coords, f32p), _p(params_h, h16p), _p(out, h16p))
        return out
    E = 2 * cfg.n_levels
    lay = param_layout(cfg)
    enc = np.zeros((E, n), dtype=np.uint16)
    normal = np.zeros((n, 3), dtype=np.float32)
    sdf_in = np.zeros((n, lay["sdf_in"]), dtype=np.uint16)
    rgb_in = np.zeros((n, 48), dtype=np.uint16)
    sdf_hidden = np.zeros((cfg.sdf_hidden_layers, n, 64), dtype=np.uint16)
    rgb_hidden = np.zeros((cfg.rgb_hidden_layers, n, 64), dtype=np.uint16)
    lib().orc_net_forward_debug(C.byref(cfg), C.c_uint32(n), _p(coords, f32p), _p(params_h, h16p), _p(out, h16p),
                                _p(enc, h16p), _p(normal, f32p), _p(sdf_in, h16p), _p(rgb_in, h16p), _p(sdf_hidden, h16p),
                                _p(rgb_hidden, h16p))
    return out, dict(enc=enc, normal=normal, sdf_in=sdf_in, rgb_in=rgb_in, sdf_hidden=sdf_hidden, rgb_hidden=rgb_hidden)


def net_backward(cfg, coords, params_h, dL_dout_h, indeed_batch_size):
    n = coords.shape[0]
    grads = np.zeros(net_n_params(cfg), dtype=np.float32)
    lib().orc_net_backward(C.byref(cfg), C.c_uint32(n), _p(coords, f32p), _p(params_h, h16p), _p(dL_dout_h, h16p),
                           C.c_uint32(indeed_batch_size), _p(grads, f32p))
    return grads


def _h(a):
    """float array -> the half bit patterns the reference stores (uint16)."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).astype(np.float16).view(np.uint16))


def delta_matrix(rot6, trans):
    """rotation_6d_to_matrix + the transition as floats: (R [3][3] row-major, t [3])."""
    R = np.zeros(9, dtype=np.float32); t = np.zeros(3, dtype=np.float32)
    lib().orc_delta_matrix(_p(_h(rot6), h16p), _p(_h(trans), h16p), _p(R, f32p), _p(t, f32p))
    return R.reshape(3, 3), t


def delta_apply(coords, rot6, trans, first_frame_offset=None):
    """DeltaNetwork forward on [n][7] coordinates (position, dt, warped direction)."""
    coords = np.ascontiguousarray(coords, dtype=np.float32)
    out = np.zeros_like(coords)
    off = None if first_frame_offset is None else np.ascontiguousarray(first_frame_offset, dtype=np.float32)
    lib().orc_delta_apply(C.c_uint32(coords.shape[0]), _p(coords, f32p), _p(_h(rot6), h16p), _p(_h(trans), h16p),
                          _p(off, f32p) if off is not None else None, _p(out, f32p))
    return out


def delta_backward(coords, dL_dpos, rot6, trans, first_frame_offset=None):
    """DeltaNetwork backward: grad[12] = [d transition 3 | 0 | d rotation-6d 6 | 0 0] as (float)half."""
    coords = np.ascontiguousarray(coords, dtype=np.float32)
    dL_dpos = np.ascontiguousarray(dL_dpos, dtype=np.float32)
    off = None if first_frame_offset is None else np.ascontiguousarray(first_frame_offset, dtype=np.float32)
    grad = np.zeros(12, dtype=np.float32)
    lib().orc_delta_backward(C.c_uint32(coords.shape[0]), _p(coords, f32p), _p(dL_dpos, f32p), _p(_h(rot6), h16p), _p(_h(trans), h16p),
                             _p(off, f32p) if off is not None else None, _p(grad, f32p))
    return grad


def delta_backward_each(coords, dL_dpos, rot6, trans, first_frame_offset=None):
    """Per-sample DeltaNetwork gradients [n][9] = [d transition 3 | d rotation-6d 6], each as (float)half (before the batch sum)."""
    coords = np.ascontiguousarray(coords, dtype=np.float32)
    dL_dpos = np.ascontiguousarray(dL_dpos, dtype=np.float32)
    off = None if first_frame_offset is None else np.ascontiguousarray(first_frame_offset, dtype=np.float32)
    out = np.zeros((coords.shape[0], 9), dtype=np.float32)
    lib().orc_delta_backward_each(C.c_uint32(coords.shape[0]), _p(coords, f32p), _p(dL_dpos, f32p), _p(_h(rot6), h16p), _p(_h(trans), h16p),
                                  _p(off, f32p) if off is not None else None, _p(out, f32p))
    return out


def net_backward_delta(cfg, coords, params_h, dL_dout_h, indeed_batch_size, rot6, trans, first_frame_offset=None):
    """backward_impl with m_use_delta && m_train_delta: (canonical grads, delta grad[12], dL/d deformed position [n][3])."""
    n = coords.shape[0]
    coords = np.ascontiguousarray(coords, dtype=np.float32)
    grads = np.zeros(net_n_params(cfg), dtype=np.float32)
    dgrad = np.zeros(12, dtype=np.float32)
    dpos = np.zeros((n, 3), dtype=np.float32)
    off = None if first_frame_offset is None else np.ascontiguousarray(first_frame_offset, dtype=np.float32)
    lib().orc_net_backward_delta(C.byref(cfg), C.c_uint32(n), _p(coords, f32p), _p(params_h, h16p), _p(dL_dout_h, h16p),
                                 C.c_uint32(indeed_batch_size), _p(_h(rot6), h16p), _p(_h(trans), h16p),
                                 _p(off, f32p) if off is not None else None, _p(grads, f32p), _p(dgrad, f32p), _p(dpos, f32p))
    return grads, dgrad, dpos


def accumulate_global_movement(rot6, trans, acc_rot, acc_trans):
    """accumulated <- delta o accumulated; returns (R [3][3], t [3]) as the floats of the stored halves."""
    r = _h(np.asarray(acc_rot, dtype=np.float32).reshape(9)); t = _h(acc_trans)
    lib().orc_accumulate_global_movement(_p(_h(rot6), h16p), _p(_h(trans), h16p), _p(r, h16p), _p(t, h16p))
    return r.view(np.float16).astype(np.float32).reshape(3, 3), t.view(np.float16).astype(np.float32)


def net_sdf(cfg, pos, params_h, as_density=False):
    n = pos.shape[0]
    out = np.zeros(n, dtype=np.uint16)
    lib().orc_net_sdf(C.byref(cfg), C.c_uint32(n), _p(pos, f32p), C.c_uint32(pos.shape[1]), _p(params_h, h16p),
                      C.c_int(1 if as_density else 0), _p(out, h16p))
    return out


def adam_ema_step(n_matrix, loss_scale, lr, beta1, beta2, eps, l2_reg, ema_decay, ema_step, grads_f32, master, params_h, m1, m2, steps, ema_h):
    n = master.shape[0]
    lib().orc_adam_ema_step(C.c_uint64(n), C.c_uint64(n_matrix), C.c_float(loss_scale), C.c_float(lr), C.c_float(beta1), C.c_float(beta2),
                            C.c_float(eps), C.c_float(l2_reg), C.c_float(ema_decay), C.c_uint32(ema_step), _p(grads_f32, f32p),
                            _p(master, f32p), _p(params_h, h16p), _p(m1, f32p), _p(m2, f32p), _p(steps, u32p), _p(ema_h, h16p))


# ---------------------------------------------------------------------------
# ngp layer
# ---------------------------------------------------------------------------
def make_dataset(pixels, focal, principal, xforms):
    """pixels uint8 [n,h,w,4]; focal [n,2]; principal [n,2] (relative); xforms [n,3,4] float32 (row-major)."""
    ds = Dataset()
    ds.n_images = pixels.shape[0]
    ds.height = pixels.shape[1]
    ds.width = pixels.shape[2]
    ds.pixels = _p(pixels, u8p)
    ds.focal = _p(focal, f32p)
    ds.principal = _p(principal, f32p)
    ds.xforms = _p(xforms, f32p)
    ds._keep = (pixels, focal, principal, xforms)
    return ds


IDENTITY_ROT_H = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], dtype=np.float16).view(np.uint16)
ZERO_TRANS_H = np.zeros(4, dtype=np.float16).view(np.uint16)
UNIT_AABB = np.array([0, 0, 0, 1, 1, 1], dtype=np.float32)


def generate_training_samples(n_rays, max_samples, n_rays_total, rng, ds, bitfield, aabb=UNIT_AABB, snap=True, cone_angle=0.0,
                              rotation=IDENTITY_ROT_H, transition=ZERO_TRANS_H, first_frame_offset=None):
    ffo = np.zeros(3, dtype=np.float32) if first_frame_offset is None else first_frame_offset
    numsteps_all = np.zeros(n_rays, dtype=np.uint32)
    ray_indices = np.zeros(n_rays, dtype=np.uint32)
    rays = np.zeros((n_rays, 6), dtype=np.float32)
    numsteps = np.zeros((n_rays, 2), dtype=np.uint32)
    coords = np.zeros((max_samples, 7), dtype=np.float32)
    counters = np.zeros(2, dtype=np.uint32)
    lib().orc_generate_training_samples(C.c_uint32(n_rays), _p(aabb, f32p), C.c_uint32(max_samples), C.c_uint32(n_rays_total),
                                        _p(rng.st, u64p), C.byref(ds), _p(bitfield, u8p), C.c_int(1 if snap else 0),
                                        C.c_float(cone_angle), _p(rotation, h16p), _p(transition, h16p), _p(ffo, f32p),
                                        _p(numsteps_all, u32p), _p(ray_indices, u32p), _p(rays, f32p), _p(numsteps, u32p),
                                        _p(coords, f32p), _p(counters, u32p))
    return dict(numsteps_all=numsteps_all, ray_indices=ray_indices, rays=rays, numsteps=numsteps, coords=coords,
                n_kept=int(counters[0]), n_samples=int(counters[1]))


def compute_loss(n_rays, n_kept, n_rays_total, rng, max_samples_compacted, lp, ds, network_output_h, ray_indices, rays, numsteps,
                 coords, aabb=UNIT_AABB, rotation=IDENTITY_ROT_H, transition=ZERO_TRANS_H, first_frame_offset=None):
    ffo = np.zeros(3, dtype=np.float32) if first_frame_offset is None else first_frame_offset
    numsteps = numsteps.copy()
    coords_compacted = np.zeros((max_samples_compacted, 7), dtype=np.float32)
    dloss = np.zeros((max_samples_compacted, 16), dtype=np.uint16)
    loss = np.zeros(n_rays, dtype=np.float32)
    ek = np.zeros(n_rays, dtype=np.float32)
    mask = np.zeros(n_rays, dtype=np.float32)
    counters = np.zeros(1, dtype=np.uint32)
    lib().orc_compute_loss(C.c_uint32(n_rays), C.c_uint32(n_kept), _p(aabb, f32p), C.c_uint32(n_rays_total), _p(rng.st, u64p),
                           C.c_uint32(max_samples_compacted), C.byref(lp), C.byref(ds), _p(network_output_h, h16p),
                           _p(ray_indices, u32p), _p(rays, f32p), _p(numsteps, u32p), _p(coords, f32p), _p(coords_compacted, f32p),
                           _p(dloss, h16p), _p(loss, f32p), _p(ek, f32p), _p(mask, f32p), _p(rotation, h16p), _p(transition, h16p),
                           _p(ffo, f32p), _p(counters, u32p))
    return dict(numsteps=numsteps, coords_compacted=coords_compacted, dloss=dloss, loss=loss, ek_loss=ek, mask_loss=mask,
                n_compacted=int(counters[0]))


def fill_rollover(n_elements, n_input, dl_h=None, coords=None):
    lib().orc_fill_rollover(C.c_uint32(n_elements), C.c_uint32(7), C.c_uint32(n_input), _p(dl_h, h16p), C.c_uint32(16), _p(coords, f32p))


def generate_grid_samples(n, rng, step, grid_in, n_cascades, thresh, aabb=UNIT_AABB):
    pos = np.zeros((n, 3), dtype=np.float32)
    idx = np.zeros(n, dtype=np.uint32)
    lib().orc_generate_grid_samples(C.c_uint32(n), _p(rng.st, u64p), C.c_uint32(step), _p(aabb, f32p), _p(grid_in, f32p),
                                    C.c_uint32(n_cascades), C.c_float(thresh), _p(pos, f32p), _p(idx, u32p))
    return pos, idx


def splat_and_decay(indices, density_h, decay, grid):
    n_elements = grid.shape[0]
    tmp = np.zeros(n_elements, dtype=np.float32)
    lib().orc_splat_and_decay(C.c_uint32(indices.shape[0]), _p(indices, u32p), _p(density_h, h16p), C.c_uint32(n_elements),
                              C.c_float(decay), _p(tmp, f32p), _p(grid, f32p))


def update_bitfield(grid, n_cascades_used=1):
    bitfield = np.zeros(128 * 128 * 128 // 8 * 8, dtype=np.uint8)
    mean = lib().orc_update_bitfield(_p(grid, f32p), C.c_uint32(n_cascades_used), _p(bitfield, u8p))
    return bitfield, float(mean)


# ---------------------------------------------------------------------------
# inference render (a11) and marching cubes (a12)
# ---------------------------------------------------------------------------
class RenderCam(C.Structure):
    _fields_ = [
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("focal", C.c_float * 2),
        ("screen_center", C.c_float * 2),
        ("camera", C.c_float * 12),
        ("sample_index", C.c_uint32),
        ("snap_to_pixel_centers", C.c_int32),
        ("min_transmittance", C.c_float),
        ("cone_angle_constant", C.c_float),
        ("cos_anneal_ratio", C.c_float),
    ]


def make_render_cam(width, height, focal, screen_center, camera, sample_index=0, snap=True, min_transmittance=0.01,
                    cone_angle=0.0, cos_anneal_ratio=1.0):
    c = RenderCam()
    c.width, c.height = width, height
    c.focal[:] = [float(focal[0]), float(focal[1])]
    c.screen_center[:] = [float(screen_center[0]), float(screen_center[1])]
    c.camera[:] = [float(v) for v in np.asarray(camera, dtype=np.float32).reshape(-1)]
    c.sample_index = sample_index
    c.snap_to_pixel_centers = 1 if snap else 0
    c.min_transmittance = min_transmittance
    c.cone_angle_constant = cone_angle
    c.cos_anneal_ratio = cos_anneal_ratio
    return c


def ld_random_val(index, seed):
    lib().orc_ld_random_val.restype = C.c_float
    return float(lib().orc_ld_random_val(C.c_uint32(index), C.c_uint32(seed)))


def sobol(index, dim):
    lib().orc_sobol.restype = C.c_uint32
    return int(lib().orc_sobol(C.c_uint32(index), C.c_uint32(dim)))


def ld_random_pixel_offset(spp):
    out = np.zeros(2, dtype=np.float32)
    lib().orc_ld_random_pixel_offset(C.c_uint32(spp), _p(out, f32p))
    return out


def render_init_rays(cam, bitfield, render_aabb=UNIT_AABB, rotation=IDENTITY_ROT_H, transition=ZERO_TRANS_H, first_frame_offset=None):
    ffo = np.zeros(3, dtype=np.float32) if first_frame_offset is None else first_frame_offset
    payload = np.zeros((cam.width * cam.height, 8), dtype=np.float32)
    lib().orc_render_init_rays(C.byref(cam), _p(render_aabb, f32p), _p(rotation, h16p), _p(transition, h16p), _p(ffo, f32p),
                               _p(bitfield, u8p), _p(payload, f32p))
    return payload


def render_trace(cfg, params_h, cam, bitfield, payload, render_aabb=UNIT_AABB, train_aabb=UNIT_AABB):
    n = cam.width * cam.height
    rgba = np.zeros((n, 4), dtype=np.float32)
    depth = np.zeros(n, dtype=np.float32)
    n_samples = np.zeros(n, dtype=np.uint32)
    lib().orc_render_trace(C.byref(cfg), _p(params_h, h16p), C.byref(cam), _p(render_aabb, f32p), _p(train_aabb, f32p),
                           _p(bitfield, u8p), _p(payload, f32p), _p(rgba, f32p), _p(depth, f32p), _p(n_samples, u32p))
    return rgba, depth, n_samples


def render_shade_accumulate_tonemap(rgba, depth, accumulate, sample_count, background_color, exposure=0.0, to_srgb=True):
    n = rgba.shape[0]
    out = np.zeros((n, 4), dtype=np.float32)
    depth_out = np.zeros(n, dtype=np.float32)
    bg = np.asarray(background_color, dtype=np.float32)
    lib().orc_render_shade_accumulate_tonemap(C.c_uint32(n), _p(rgba, f32p), _p(depth, f32p), C.c_float(sample_count), _p(accumulate, f32p),
                                              _p(bg, f32p), C.c_float(exposure), C.c_int(1 if to_srgb else 0), _p(out, f32p), _p(depth_out, f32p))
    return out, depth_out


def render(cfg, params_h, cam, bitfield, background_color=(0, 0, 0, 1), to_srgb=True, spp=1, **kw):
    """Testbed::render_to_cpu (python_api.cu:123-170) for spp samples; returns [h, w, 4] float32."""
    acc = np.zeros((cam.width * cam.height, 4), dtype=np.float32)
    out = None
    for s in range(spp):
        cam.sample_index = s
        payload = render_init_rays(cam, bitfield, **kw)
        rgba, depth, _ = render_trace(cfg, params_h, cam, bitfield, payload)
        out, _ = render_shade_accumulate_tonemap(rgba, depth, acc, float(s), background_color, 0.0, to_srgb)
    return out.reshape(cam.height, cam.width, 4)


def sdf_on_lattice(cfg, params_h, res, render_aabb=UNIT_AABB, train_aabb=UNIT_AABB, z0=0, z1=None):
    res = np.asarray(res, dtype=np.uint32)
    z1 = int(res[2]) if z1 is None else z1
    out = np.zeros((z1 - z0, int(res[1]), int(res[0])), dtype=np.float32)
    lib().orc_sdf_on_lattice(C.byref(cfg), _p(params_h, h16p), _p(res, u32p), _p(render_aabb, f32p), _p(train_aabb, f32p),
                             C.c_uint32(z0), C.c_uint32(z1), _p(out, f32p))
    return out


def marching_cubes(density, aabb=UNIT_AABB, thresh=0.0):
    """density: float32 [rz, ry, rx].  Returns (verts [nv,3] float32, indices [nt,3] uint32), deterministic numbering."""
    res = np.array([density.shape[2], density.shape[1], density.shape[0]], dtype=np.uint32)
    counts = np.zeros(2, dtype=np.uint32)
    density = np.ascontiguousarray(density, dtype=np.float32)
    lib().orc_marching_cubes(_p(res, u32p), _p(aabb, f32p), C.c_float(thresh), _p(density, f32p), _p(counts, u32p), None, None)
    verts = np.zeros((int(counts[0]), 3), dtype=np.float32)
    idx = np.zeros(int(counts[1]), dtype=np.uint32)
    lib().orc_marching_cubes(_p(res, u32p), _p(aabb, f32p), C.c_float(thresh), _p(density, f32p), _p(counts, u32p), _p(verts, f32p), _p(idx, u32p))
    return verts, idx.reshape(-1, 3)


def mesh_vertex_normals(verts, faces):
    """compute_mesh_1ring's normals (marching_cubes.cu:331-366, 720-727): area-weighted, NOT normalised.  verts [nv,3] f32, faces [nt,3] u32."""
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    idx = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1)
    out = np.zeros((len(verts), 3), dtype=np.float32)
    lib().orc_mesh_vertex_normals(C.c_uint32(len(verts)), C.c_uint32(idx.size // 3), _p(idx, u32p), _p(verts, f32p), _p(out, f32p))
    return out


def mesh_color_inputs(verts, train_aabb=UNIT_AABB):
    """generate_nerf_network_inputs_from_positions (testbed_nerf.cu:847-853): NerfCoordinates [nv,7] for compute_mesh_vertex_colors."""
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    out = np.zeros((len(verts), 7), dtype=np.float32)
    lib().orc_mesh_color_inputs(C.c_uint32(len(verts)), _p(verts, f32p), _p(np.ascontiguousarray(train_aabb, dtype=np.float32), f32p), _p(out, f32p))
    return out


def mesh_extract_colors(out16_bits, from_linear=False):
    """extract_srgb_with_activation (testbed_nerf.cu:532-545) on network outputs [n,16] (half bits) -> [n,3] float."""
    o = np.ascontiguousarray(out16_bits, dtype=np.uint16)
    out = np.zeros((len(o), 3), dtype=np.float32)
    lib().orc_mesh_extract_colors(C.c_uint32(len(o)), _p(o, h16p), C.c_int(1 if from_linear else 0), _p(out, f32p))
    return out
