"""Size-independent properties at BASELINE.json's full sizes (batch 2^18, 2^20 inference samples, 1920x1080 frames, 512^3
lattice), where the CPU oracle is too slow to be the checker: permutation invariance, linearity, tile/slab independence,
topology invariants.  The small-size tests (test_gpu_network / _step / _render_mesh) carry the oracle comparisons."""
import numpy as np
import pytest

from tests import util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import neus2_b200 as N
    c = N.Context(N.default_config())            # base.json sizes: B = 2^18, 10.56 M parameters
    cfg = U.base_cfg()
    p = U.random_params(cfg, seed=21)
    p[c.layout["variance"]:] = 0.5
    c.set_params(p)
    yield c
    c.close()


def test_forward_is_a_pure_per_sample_function_at_2pow20(ctx):
    """Row i of the output depends on row i of the input only: permuting 2^20+3 samples (not a multiple of the 16-sample tile)
    permutes the outputs bit for bit -- no cross-talk between the 16 rows of an mma tile, between warps, or at the ragged tail."""
    import torch
    n = (1 << 20) + 3
    coords = torch.from_numpy(U.random_coords(n, seed=5)).cuda()
    out = ctx.forward(coords)
    perm = torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    out_p = ctx.forward(coords[perm].contiguous())
    assert torch.equal(out_p.view(torch.int16), out[perm].view(torch.int16))
    sdf = ctx.sdf(coords[:, :3].contiguous())
    assert torch.equal(sdf.view(torch.int16), out[:, 3].contiguous().view(torch.int16))      # the SDF-only kernel is the same function


def test_backward_is_linear_in_the_upstream_gradient_at_batch_2pow18(ctx):
    import torch
    B = 1 << 18
    coords = torch.from_numpy(U.random_coords(B, seed=6)).cuda()
    rs = np.random.RandomState(7)
    d1 = np.zeros((B, 16), dtype=np.float16)
    d2 = np.zeros((B, 16), dtype=np.float16)
    d1[:, :11] = rs.normal(0, 2.0 ** -6, (B, 11)).astype(np.float16)
    d2[:, :11] = rs.normal(0, 2.0 ** -6, (B, 11)).astype(np.float16)
    # sums exactly representable in half keep the test about the kernel, not about input rounding
    d1 = (np.round(d1.astype(np.float32) * 4096) / 4096).astype(np.float16)
    d2 = (np.round(d2.astype(np.float32) * 4096) / 4096).astype(np.float16)
    d12 = (d1.astype(np.float32) + d2.astype(np.float32)).astype(np.float16)
    assert np.array_equal(d12.astype(np.float32), d1.astype(np.float32) + d2.astype(np.float32))
    g = []
    for d in (d1, d2, d12):
        g.append(ctx.backward(coords, torch.from_numpy(d).cuda(), B, B, B).clone())
    lay = ctx.layout
    for name, a, b in (("mlp", 0, lay["grid"]), ("grid", lay["grid"], lay["variance"])):
        s = g[0][a:b] + g[1][a:b]
        rel = ((g[2][a:b] - s).norm() / s.norm()).item()
        assert rel < 2e-3, (name, rel)                       # half rounding of the intermediate gradients is the only non-linearity
    # sparsity pattern of the hash-grid gradient is a function of the coordinates only
    z = [(x[lay["grid"]:lay["variance"]] == 0) for x in g]
    assert torch.equal(z[0] | z[1], z[0]) or (z[2] >= (z[0] & z[1])).all()


def test_adam_leaves_untouched_grid_entries_alone_and_is_deterministic(ctx):
    import torch
    master, params, ema, grads = ctx.buffers()
    p0 = ctx.get_params()
    try:
        g = torch.zeros_like(grads)
        idx = torch.arange(ctx.layout["grid"], ctx.layout["variance"], 7, device="cuda")
        g[idx] = 0.25
        g[: ctx.layout["grid"]] = 0.0                         # MLP weights: zero gradient still decays through L2 (adam.h:132-135)
        grads.copy_(g)
        before = master.clone()
        ctx.optimizer_step(128.0)
        after = master.clone()
        changed = after != before
        untouched = torch.ones_like(changed)
        untouched[idx] = False
        untouched[: ctx.layout["grid"]] = False
        assert not bool(changed[untouched].any())             # sparse skip: adam.h:109-112
        assert bool(changed[idx].all())
        ctx.set_params(p0)
        grads.copy_(g)
        ctx.optimizer_step(128.0)
        assert torch.equal(master, after)                     # bitwise reproducible
    finally:
        ctx.set_params(p0)


def test_render_tiles_and_lattice_slabs_at_full_size(ctx):
    import torch
    from neus2_b200 import synthetic
    ctx.update_density_grid()
    cam = synthetic.camera_ring(1, radius=1.6, seed=3)[0]
    W, H = 1920, 1080
    rp = ctx.render_params(W, H, camera_matrix=cam, focal_length=(1500.0, 1500.0), use_inference_params=0)
    whole = ctx.render(rp).clone()
    assert whole.shape == (H, W, 4) and bool(torch.isfinite(whole).all())
    st = ctx.render_stats()
    assert st["rays_alive"] > 10000 and st["samples"] >= st["rays_alive"]
    parts = torch.empty_like(whole)
    for y0 in range(0, H, 270):
        rp.tile_x0, rp.tile_y0, rp.tile_w, rp.tile_h = 0, y0, W, 270
        parts[y0:y0 + 270] = ctx.render(rp)
    assert torch.equal(parts, whole)
    res = (512, 512, 512)
    lat = ctx.sdf_lattice(res, use_inference_params=False)
    slab = ctx.sdf_lattice(res, z0=200, z1=264, use_inference_params=False)
    assert torch.equal(slab, lat[200:264])
    v, f = ctx.marching_cubes(lat)
    if len(v):
        assert int(f.max()) == len(v) - 1 and int(f.min()) == 0
        e = torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]).sort(dim=1).values
        _, cnt = torch.unique(e[:, 0] * len(v) + e[:, 1], return_counts=True)
        assert int(cnt.max()) <= 2                             # a level set of a lattice function is a 2-manifold with boundary at the box


def test_marching_cubes_512_sphere_euler_characteristic(ctx):
    import torch
    r = 512
    ax = torch.arange(r, device="cuda", dtype=torch.float32) / r
    zz, yy, xx = torch.meshgrid(ax, ax, ax, indexing="ij")
    dens = (0.37 - torch.sqrt((xx - 0.5) ** 2 + (yy - 0.5) ** 2 + (zz - 0.5) ** 2)).contiguous()
    del zz, yy, xx
    v, f = ctx.marching_cubes(dens)
    e = torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]).sort(dim=1).values
    uniq, cnt = torch.unique(e[:, 0] * len(v) + e[:, 1], return_counts=True)
    assert bool((cnt == 2).all()) and len(v) - len(uniq) + len(f) == 2 and len(v) > 500_000
