"""Shared helpers for the parity tests: seeded parameters / inputs in the reference's value ranges."""
import numpy as np

from oracle import oracle as O


def base_cfg(**kw):
    return O.make_cfg(**kw)


def random_params(cfg, seed=0, grid_scale=0.05, mlp_scale=None):
    """fp32 master params in the reference order; grid values large enough that the encoding is not all-subnormal."""
    rs = np.random.RandomState(seed)
    lay = O.param_layout(cfg)
    p = np.zeros(lay["total"], dtype=np.float32)
    SI = lay["sdf_in"]
    # sdf MLP: geometric-init-like first layer (only xyz columns large) plus small noise elsewhere; last layer ~ sqrt(pi)/8
    W0 = rs.normal(0, 0.05, (64, SI)).astype(np.float32)
    W0[:, :3] = rs.normal(0, np.sqrt(2) / 8, (64, 3))
    W0[:, 3 + 2 * cfg.n_levels:] = 0
    W1 = (np.sqrt(np.pi) / 8 + rs.normal(0, 0.05, (16, 64))).astype(np.float32)
    p[0:64 * SI] = W0.ravel()
    p[64 * SI:64 * SI + 1024] = W1.ravel()
    o = lay["rgb"]
    for (a, b, shape) in [(0, 3072, (64, 48)), (3072, 7168, (64, 64)), (7168, 8192, (16, 64))]:
        s = np.sqrt(6.0 / (shape[0] + shape[1])) if mlp_scale is None else mlp_scale
        p[o + a:o + b] = rs.uniform(-s, s, b - a)
    p[lay["grid"]:lay["variance"]] = rs.uniform(-grid_scale, grid_scale, lay["n_grid"])
    p[lay["variance"]:] = 0.3
    return p


def to_half_bits(p32):
    return p32.astype(np.float16).view(np.uint16)


def random_coords(n, seed=0, along_rays=True):
    """NerfCoordinates [n,7]: warped pos in (0,1), dt=0 (aabb_scale 1), warped dir in (0,1)."""
    rs = np.random.RandomState(seed)
    c = np.zeros((n, 7), dtype=np.float32)
    if along_rays:
        n_rays = max(1, n // 64)
        o = rs.uniform(0.1, 0.9, (n_rays, 3))
        d = rs.normal(0, 1, (n_rays, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        t = (np.arange(n) % 64) * (np.sqrt(3) / 1024)
        r = np.minimum(np.arange(n) // 64, n_rays - 1)
        pos = o[r] + t[:, None] * d[r]
        c[:, :3] = np.clip(pos, 0.001, 0.999)
        c[:, 4:7] = (d[r] + 1) * 0.5
    else:
        c[:, :3] = rs.uniform(0, 1, (n, 3))
        d = rs.normal(0, 1, (n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        c[:, 4:7] = (d + 1) * 0.5
    return c


def half_ulp_diff(a_bits, b_bits):
    """|a-b| in units of half ulps (via the monotone integer mapping of binary16)."""
    def key(x):
        x = x.astype(np.int32)
        return np.where(x & 0x8000, 0x8000 - (x & 0x7FFF) - 0, 0x8000 + (x & 0x7FFF))
    return np.abs(key(a_bits) - key(b_bits))
