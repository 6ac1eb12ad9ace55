"""CPU checks of the oracle's render / marching-cubes restatement (no GPU): structural properties the domain offers, since
the reference ships no golden vectors for this layer and its ngp sources do not compile here."""
import numpy as np

from oracle import oracle as O
from tests import util as U


def test_ld_sequence_properties():
    assert O.sobol(0, 0) == 0 and O.sobol(0, 1) == 0
    # dimension 0 = van der Corput; dimension 1 = published Sobol direction numbers (first eight)
    assert [O.sobol(1 << k, 0) for k in range(4)] == [0x80000000, 0x40000000, 0x20000000, 0x10000000]
    assert [O.sobol(1 << k, 1) for k in range(8)] == [0x80000000, 0xc0000000, 0xa0000000, 0xf0000000, 0x88000000, 0xcc000000, 0xaa000000, 0xff000000]
    # linearity over GF(2)
    for a, b in [(5, 9), (1234, 77), (0xffff, 0x10000)]:
        assert O.sobol(a ^ b, 1) == O.sobol(a, 1) ^ O.sobol(b, 1)
    v = np.array([O.ld_random_val(i, 786433 * 17) for i in range(2048)])
    assert v.min() >= 0 and v.max() <= 1 and abs(v.mean() - 0.5) < 0.01
    # stratification of a scrambled (0,1)-sequence: 2048 points put exactly 128 in each of 16 equal bins
    assert np.array_equal(np.histogram(v, bins=16, range=(0, 1))[0], np.full(16, 128))
    assert np.array_equal(O.ld_random_pixel_offset(0), np.array([0.5, 0.5], dtype=np.float32))
    off = np.stack([O.ld_random_pixel_offset(s) for s in range(1, 65)])
    assert off.min() >= 0 and off.max() < 1


def test_marching_cubes_oracle_sphere_and_plane():
    r = 40
    z, y, x = np.meshgrid(*(np.arange(r) / r,) * 3, indexing="ij")
    dens = (0.3 - np.sqrt((x - .5) ** 2 + (y - .5) ** 2 + (z - .5) ** 2)).astype(np.float32)
    v, f = O.marching_cubes(dens)
    assert np.abs(np.linalg.norm(v - 0.5, axis=1) - 0.3).max() < 3e-4
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
    u, c = np.unique(e, axis=0, return_counts=True)
    assert (c == 2).all() and len(v) - len(u) + len(f) == 2
    # orientation: with inside = density > thresh the table winds triangles so that normals point from inside to outside
    # consistently (same sign for every face of a convex body)
    n = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    big = np.linalg.norm(n, axis=1) > 1e-5          # slivers next to lattice points have no reliable normal
    s = np.sign((n[big] * (v[f[big]].mean(1) - 0.5)).sum(1))
    assert big.mean() > 0.9 and abs(s.sum()) == big.sum()
    # an axis-aligned plane between lattice layers: one vertex per (x,y) column, two triangles per cell
    dens = np.broadcast_to((0.5 - (np.arange(8) + 0.25) / 8).astype(np.float32)[:, None, None], (8, 6, 5)).copy()
    v, f = O.marching_cubes(dens)
    assert len(v) == 6 * 5 and len(f) == 2 * 5 * 4
    assert np.allclose(v[:, 2], (3 + 0.75) / 8, atol=1e-6)
    # no surface -> nothing
    v, f = O.marching_cubes(np.ones((4, 4, 4), dtype=np.float32))
    assert len(v) == 0 and len(f) == 0


def test_render_oracle_on_a_ball():
    cfg = U.base_cfg()
    lay = O.param_layout(cfg)
    p = U.random_params(cfg, seed=1, grid_scale=1e-4)
    p[lay["variance"]:] = 0.6
    ph = U.to_half_bits(p)
    # occupancy: cells within 0.32 of the centre
    g = (np.arange(128) + 0.5) / 128
    xx, yy, zz = np.meshgrid(g, g, g, indexing="ij")
    occ = ((xx - .5) ** 2 + (yy - .5) ** 2 + (zz - .5) ** 2) < 0.32 ** 2
    grid = np.zeros(128 ** 3, dtype=np.float32)
    ii, jj, kk = np.nonzero(occ)
    # vectorised morton code
    def expand(v):
        v = v.astype(np.uint64)
        v = (v * 0x00010001) & 0xFF0000FF
        v = (v * 0x00000101) & 0x0F00F00F
        v = (v * 0x00000011) & 0xC30C30C3
        v = (v * 0x00000005) & 0x49249249
        return v
    m = (expand(ii) | (expand(jj) << 1) | (expand(kk) << 2)).astype(np.int64)
    grid[m] = 1.0
    bitfield, _ = O.update_bitfield(grid)
    W, H = 20, 14
    cam = np.array([[1, 0, 0, 0.5], [0, -1, 0, 0.5], [0, 0, -1, 2.0]], dtype=np.float32)
    rc = O.make_render_cam(W, H, (30.0, 30.0), (0.5, 0.5), cam)
    pay = O.render_init_rays(rc, bitfield)
    alive = pay[:, 7] > 0
    assert 20 < alive.sum() < W * H
    # every live ray starts inside an occupied cell, at or after the near plane, on a unit direction
    assert np.allclose(np.linalg.norm(pay[alive, 3:6], axis=1), 1, atol=1e-6) and (pay[alive, 6] > 0.2).all()
    pos = pay[alive, :3] + pay[alive, 3:6] * pay[alive, 6:7]
    assert (np.linalg.norm(pos - 0.5, axis=1) < 0.32 + 2 * np.sqrt(3) / 128).all()
    rgba, depth, n = O.render_trace(cfg, ph, rc, bitfield, pay)
    assert (n[~alive] == 0).all() and (n[alive] > 0).all() and n.max() <= 1024
    assert rgba[:, 3].max() <= 1.0 + 1e-6 and (rgba >= 0).all()
    sat = rgba[:, 3] > 0.99
    assert np.allclose(rgba[sat, 3], 1.0)      # saturated rays are renormalised (composite_kernel_nerf, testbed_nerf.cu:1150-1153)
    # shading, accumulation and tone mapping: black background leaves alpha, linear output is the sRGB->linear of the ray colour
    acc = np.zeros((W * H, 4), dtype=np.float32)
    out, _ = O.render_shade_accumulate_tonemap(rgba, depth, acc, 0.0, (0, 0, 0, 0), 0.0, False)
    assert np.allclose(out[:, 3], rgba[:, 3])
    lin = np.array([[O.lib().orc_srgb_to_linear(float(c)) for c in px[:3]] for px in rgba])
    assert np.allclose(out[:, :3], lin, atol=1e-7)
    # a second identical sample leaves the running mean unchanged
    out2, _ = O.render_shade_accumulate_tonemap(rgba, depth, acc, 1.0, (0, 0, 0, 0), 0.0, False)
    assert np.allclose(out2, out, atol=1e-6)
