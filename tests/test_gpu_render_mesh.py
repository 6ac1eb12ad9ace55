"""GPU parity of the inference render and of the SDF lattice + marching cubes (a12) against the CPU oracle.

Integer / pure-fp32 work must match bit-exactly: which pixels enter the occupancy grid, the ray origin / direction / first
sample distance (Sobol jitter + DDA skip), marching-cubes vertex positions and triangle indices.  Colours go through the
fp16 network (<= 2 half-ulp per output against the oracle, tests/test_gpu_network.py) and a discrete saturation test, so
images are compared with a tolerance and the per-ray sample counts on all but a small fraction of rays.
"""
import numpy as np
import pytest

from oracle import oracle as O
from tests import util as U

pytestmark = pytest.mark.gpu

W, H = 56, 40
G3 = 128 ** 3


@pytest.fixture(scope="module")
def scene():
    import torch
    import neus2_b200 as N
    from neus2_b200 import synthetic
    cfg = U.base_cfg()
    ctx = N.Context(N.default_config(batch_size=1 << 14))
    ctx.init_params()                       # geometric init: the SDF of a ball around the cube centre
    p = ctx.get_params()
    p[ctx.layout["variance"]:] = 0.55       # inv_s = exp(5.5): a surface sharp enough that rays saturate
    rs = np.random.RandomState(5)
    g0, g1 = ctx.layout["grid"], ctx.layout["variance"]
    p[g0:g1] = rs.uniform(-0.02, 0.02, g1 - g0)   # some high-frequency detail so that every level matters
    ctx.set_params(p)
    ctx.update_density_grid()               # occupancy from the network itself (training_step 0: uniform samples of every cell)
    torch.cuda.synchronize()
    bitfield = ctx.debug_buffer("bitfield", "|u1", (G3 // 8 * 8,)).cpu().numpy().copy()
    assert 0 < np.unpackbits(bitfield[:G3 // 8]).sum() < G3
    xf = synthetic.camera_ring(3, radius=1.25, seed=3)
    yield dict(ctx=ctx, cfg=cfg, p=p, ph=U.to_half_bits(p), bitfield=bitfield, xforms=xf)
    ctx.close()


def _params(scene, view=0, **kw):
    ctx = scene["ctx"]
    focal = np.array([70.0, 68.0], dtype=np.float32)
    kw.setdefault("use_inference_params", 0)
    return ctx.training_view_params(scene["xforms"][view], focal, (0.52, 0.47), (W, H), **kw), focal


def _oracle_cam(scene, rp, sample_index=0):
    return O.make_render_cam(rp.width, rp.height, list(rp.focal_length), list(rp.screen_center), np.array(list(rp.camera_matrix), dtype=np.float32),
                             sample_index=sample_index, snap=bool(rp.snap_to_pixel_centers), min_transmittance=rp.min_transmittance)


def test_sobol_and_pixel_offset_oracle_properties():
    # Sobol dim 0 is the bit reversal, dim 1 starts 0x80000000, 0xc0000000, 0xa0000000, 0xf0000000 (published direction numbers)
    assert [O.sobol(1 << k, 1) for k in range(5)] == [0x80000000, 0xc0000000, 0xa0000000, 0xf0000000, 0x88000000]
    assert all(O.sobol(i, 0) == int(f"{i:032b}"[::-1], 2) for i in (1, 2, 3, 77, 0xdeadbeef))
    assert np.array_equal(O.ld_random_pixel_offset(0), np.array([0.5, 0.5], dtype=np.float32))


@pytest.mark.parametrize("sample_index,snap", [(0, 1), (3, 0)])
def test_render_init_rays_bit_exact(scene, sample_index, snap):
    import torch
    ctx = scene["ctx"]
    rp, _ = _params(scene, view=1, snap_to_pixel_centers=snap)
    rp.sample_index = sample_index
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    N = W * H
    from neus2_b200 import check
    import ctypes as C
    check(ctx._lib.neus2_render(ctx._h, None, C.byref(rp), C.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize()
    n_alive = int(ctx.debug_buffer("render_counters", "<u4", (4,)).cpu().numpy()[0])
    q = ctx.debug_buffer("render_queue", "<f4", (N, 8)).cpu().numpy()[:n_alive]
    idx = q[:, 7].copy().view(np.uint32)
    order = np.argsort(idx)
    q, idx = q[order], idx[order]
    pay = O.render_init_rays(_oracle_cam(scene, rp, sample_index), scene["bitfield"])
    alive_ref = np.nonzero(pay[:, 7] > 0)[0]
    assert n_alive == len(alive_ref) and n_alive > 50
    assert np.array_equal(idx, alive_ref.astype(np.uint32))
    assert np.array_equal(q[:, :7].view(np.uint32), pay[alive_ref, :7].view(np.uint32)), "origin / direction / first-sample t must be bit-identical"


def test_render_image_matches_oracle(scene):
    import torch
    ctx = scene["ctx"]
    rp, _ = _params(scene, view=0, min_transmittance=0.01, background_color=(0.2, 0.4, 0.9, 1.0))
    img = ctx.render(rp, spp=1).cpu().numpy()
    torch.cuda.synchronize()
    n_gpu = ctx.debug_buffer("render_ray_samples", "<u4", (W * H,)).cpu().numpy().copy()
    stats = ctx.render_stats()
    cam = _oracle_cam(scene, rp)
    pay = O.render_init_rays(cam, scene["bitfield"])
    rgba, depth, n_ref = O.render_trace(scene["cfg"], scene["ph"], cam, scene["bitfield"], pay)
    acc = np.zeros((W * H, 4), dtype=np.float32)
    ref, _ = O.render_shade_accumulate_tonemap(rgba, depth, acc, 0.0, (0.2, 0.4, 0.9, 1.0), 0.0, True)
    ref = ref.reshape(H, W, 4)
    alive = pay[:, 7] > 0
    assert stats["rays_alive"] == int(alive.sum())
    assert stats["samples"] == int(n_gpu[alive].sum())
    # sample counts: identical except where a half-ulp of the network output moves the saturation test by one step
    same = (n_gpu[alive] == n_ref[alive]).mean()
    assert same > 0.97, same
    assert np.abs(n_gpu[alive].astype(np.int64) - n_ref[alive].astype(np.int64)).max() <= 2
    hit = (rgba[:, 3] > 0.5).reshape(H, W)
    assert hit.sum() > 100, "the test view must see the object"
    d = np.abs(img - ref)
    assert np.percentile(d, 99) < 2e-3 and d.max() < 3e-2, (np.percentile(d, [50, 99, 100]))
    # pixels that never entered the grid are exactly the background
    bg = ref[~alive.reshape(H, W)]
    assert np.array_equal(img[~alive.reshape(H, W)], bg)


def test_render_spp_accumulation_and_linear_output(scene):
    ctx = scene["ctx"]
    rp, _ = _params(scene, view=2, snap_to_pixel_centers=0, to_srgb=0)
    img = ctx.render(rp, spp=3).cpu().numpy()
    ref = O.render(scene["cfg"], scene["ph"], _oracle_cam(scene, rp), scene["bitfield"], background_color=(0, 0, 0, 1), to_srgb=False, spp=3)
    d = np.abs(img - ref)
    assert np.percentile(d, 99) < 2e-3 and d.max() < 3e-2, np.percentile(d, [50, 99, 100])


def test_render_tiles_equal_whole_image(scene):
    """Tiles of one frame are independent units (the multi-GPU split of BASELINE.json config 5): bit-identical to the whole."""
    import torch
    ctx = scene["ctx"]
    rp, _ = _params(scene, view=0)
    whole = ctx.render(rp, spp=1).clone()
    parts = torch.empty_like(whole)
    for (x0, y0, w, h) in [(0, 0, W, 13), (0, 13, 20, H - 13), (20, 13, W - 20, H - 13)]:
        rp.tile_x0, rp.tile_y0, rp.tile_w, rp.tile_h = x0, y0, w, h
        parts[y0:y0 + h, x0:x0 + w] = ctx.render(rp, spp=1)
    assert torch.equal(whole, parts)


def test_render_rejects_bad_tile(scene):
    import neus2_b200 as N
    ctx = scene["ctx"]
    rp, _ = _params(scene)
    rp.tile_w = W + 1
    with pytest.raises(N.Neus2Error):
        ctx.render(rp)


# ---- SDF lattice + marching cubes ----------------------------------------------------------------------------------------
def test_sdf_lattice_matches_oracle_and_sdf_operator(scene):
    import torch
    ctx = scene["ctx"]
    res = (48, 32, 16)
    aabb = np.array([0.1, 0.2, 0.3, 0.9, 0.8, 0.7], dtype=np.float32)
    lat = ctx.sdf_lattice(res, aabb, use_inference_params=False)
    ref = O.sdf_on_lattice(scene["cfg"], scene["ph"], res, render_aabb=aabb)
    got = lat.cpu().numpy()
    d = U.half_ulp_diff(got.astype(np.float16).view(np.uint16), ref.astype(np.float16).view(np.uint16))
    assert np.percentile(d, 99) <= 2 and d.max() <= 8, np.percentile(d, [50, 99, 100])
    # a z-slab equals the same rows of the whole lattice bit for bit (slabs are the multi-GPU unit)
    slab = ctx.sdf_lattice(res, aabb, z0=5, z1=11, use_inference_params=False)
    assert torch.equal(slab, lat[5:11])
    # and the lattice kernel equals the plain SDF operator on explicitly generated positions
    zz, yy, xx = np.meshgrid(np.arange(res[2]), np.arange(res[1]), np.arange(res[0]), indexing="ij")
    pos = np.stack([xx, yy, zz], -1).reshape(-1, 3).astype(np.float32) * (1.0 / np.array(res, dtype=np.float32))
    pos = (pos * (aabb[3:] - aabb[:3]) + aabb[:3]).astype(np.float32)
    sdf = ctx.sdf(torch.from_numpy(pos).cuda(), use_inference_params=False).float().reshape(res[2], res[1], res[0])
    assert (sdf - lat).abs().max().item() <= 2e-3   # positions here are rounded by numpy (no fma): same values up to 1 ulp of the input


def _canon(verts, faces):
    return verts, faces


@pytest.mark.parametrize("res", [(32, 32, 32), (40, 24, 18), (16, 16, 17)])
def test_marching_cubes_bit_exact_on_network_sdf(scene, res):
    import torch
    ctx = scene["ctx"]
    lat = ctx.sdf_lattice(res, use_inference_params=False)
    dens = -lat if res[0] == 40 else lat          # both orientations of the inside test
    dens = dens.contiguous()
    v, f = ctx.marching_cubes(dens, thresh=0.0)
    vr, fr = O.marching_cubes(dens.cpu().numpy(), thresh=0.0)
    assert len(vr) > 20
    assert v.shape == vr.shape and f.shape == fr.shape
    assert np.array_equal(v.cpu().numpy().view(np.uint32), vr.view(np.uint32))
    assert np.array_equal(f.cpu().numpy(), fr.astype(np.int64))


def test_marching_cubes_sphere_is_closed_manifold_at_256(scene):
    """Size-independent properties at a lattice far beyond what the oracle is run on: an analytic ball gives a closed
    2-manifold (every edge shared by exactly two triangles, Euler characteristic 2) with all vertices on the sphere."""
    import torch
    ctx = scene["ctx"]
    r = 256
    ax = torch.arange(r, device="cuda", dtype=torch.float32) / r
    zz, yy, xx = torch.meshgrid(ax, ax, ax, indexing="ij")
    dens = (0.31 - torch.sqrt((xx - 0.5) ** 2 + (yy - 0.52) ** 2 + (zz - 0.48) ** 2)).contiguous()
    v, f = ctx.marching_cubes(dens)
    assert len(v) > 100000
    rad = torch.sqrt((v[:, 0] - 0.5) ** 2 + (v[:, 1] - 0.52) ** 2 + (v[:, 2] - 0.48) ** 2)
    assert (rad - 0.31).abs().max().item() < 1e-4
    e = torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]).sort(dim=1).values
    key = e[:, 0] * len(v) + e[:, 1]
    uniq, cnt = torch.unique(key, return_counts=True)
    assert bool((cnt == 2).all())
    assert len(v) - len(uniq) + len(f) == 2
    assert int(f.max()) == len(v) - 1 and int(f.min()) == 0


def test_marching_cubes_empty_and_full(scene):
    import torch
    ctx = scene["ctx"]
    for val in (-1.0, 1.0):
        dens = torch.full((16, 16, 16), val, device="cuda")
        v, f = ctx.marching_cubes(dens)
        assert len(v) == 0 and len(f) == 0


def test_render_camera_facing_away_is_pure_background(scene):
    """No ray enters the box: the queue stays empty, the persistent kernel exits at once, every pixel is the (sRGB -> linear -> sRGB)
    background with the background's alpha."""
    import torch
    ctx = scene["ctx"]
    cam = scene["xforms"][0].copy()
    cam[:, 2] *= -1.0          # look the other way
    cam[:, 0] *= -1.0          # keep the frame right-handed
    rp = ctx.render_params(W, H, camera_matrix=cam, focal_length=(70.0, 70.0), background_color=(0.25, 0.5, 0.75, 1.0), use_inference_params=0)
    img = ctx.render(rp).cpu().numpy()
    assert ctx.render_stats() == dict(rays_alive=0, samples=0)
    ref, _ = O.render_shade_accumulate_tonemap(np.zeros((1, 4), np.float32), np.zeros(1, np.float32), np.zeros((1, 4), np.float32), 0.0, (0.25, 0.5, 0.75, 1.0), 0.0, True)
    assert np.allclose(img.reshape(-1, 4), ref[0], atol=1e-6)
    assert np.array_equal(img, np.broadcast_to(img[0, 0], img.shape))


def test_operators_accept_empty_and_single_sample_batches(scene):
    import torch
    ctx = scene["ctx"]
    empty = torch.zeros((0, 7), dtype=torch.float32, device="cuda")
    assert ctx.forward(empty).shape == (0, 16)
    one = torch.from_numpy(U.random_coords(1, seed=3)).cuda()
    many = torch.from_numpy(U.random_coords(40, seed=3)).cuda()
    many[0] = one[0]
    assert torch.equal(ctx.forward(one)[0].view(torch.int16), ctx.forward(many)[0].view(torch.int16))
    assert ctx.sdf(one[:, :3].contiguous()).shape == (1,)


def test_render_interleaved_rows_equal_whole_image(scene):
    """One call renders a GPU's share of a frame dealt in row blocks (row_block / row_stride): bit-identical to those rows of the whole."""
    import torch
    ctx = scene["ctx"]
    rp, _ = _params(scene, view=0)
    whole = ctx.render(rp, spp=1).clone()
    block, stride = 6, 3
    for r in range(stride):
        rows = [y for y in range(H) if (y // block) % stride == r]
        rp.tile_x0, rp.tile_w, rp.tile_y0, rp.tile_h = 0, W, r * block, len(rows)
        rp.row_block, rp.row_stride = block, stride
        part = ctx.render(rp, spp=1)
        assert torch.equal(part, whole[torch.tensor(rows, device="cuda")])
    rp.tile_h += block * stride            # runs past the last image row
    with pytest.raises(Exception):
        ctx.render(rp)


def test_inference_params_before_the_first_optimizer_step(scene):
    """Trainer::params_inference() is the training copy until the optimizer has stepped once (tcnn trainer.h:78-109): a freshly
    initialised model renders / meshes the same with either parameter selection."""
    import torch
    ctx = scene["ctx"]
    assert ctx.get_state()["optimizer_step"] == 0
    c = torch.from_numpy(U.random_coords(100, seed=8)).cuda()
    assert torch.equal(ctx.forward(c, use_inference_params=True).view(torch.int16), ctx.forward(c, use_inference_params=False).view(torch.int16))
    assert torch.equal(ctx.sdf_lattice((16, 16, 16), use_inference_params=True), ctx.sdf_lattice((16, 16, 16), use_inference_params=False))


def test_mesh_attributes_match_oracle(scene):
    """Vertex normals (compute_mesh_1ring) and vertex colours (compute_mesh_vertex_colors) of the extracted mesh, through the C ABI."""
    ctx = scene["ctx"]
    lat = ctx.sdf_lattice((32, 32, 32), use_inference_params=False).contiguous()
    v, f = ctx.marching_cubes(lat, thresh=0.0)
    nrm, col = ctx.mesh_attributes(len(v), use_inference_params=False)
    vr, fr = v.cpu().numpy(), f.cpu().numpy().astype(np.uint32)
    assert len(vr) > 200
    n_ref = O.mesh_vertex_normals(vr, fr)
    scale = np.abs(n_ref).max()
    # same face normals (one fused multiply-add per component on both sides); the float atomics add them in arrival order
    assert np.abs(nrm.cpu().numpy() - n_ref).max() <= 1e-5 * scale
    coords = O.mesh_color_inputs(vr)                      # aabb_scale 1: the training box is the unit cube
    c_ref = O.mesh_extract_colors(O.net_forward(scene["cfg"], coords, scene["ph"]), from_linear=False)
    d = np.abs(col.cpu().numpy() - c_ref)
    assert np.percentile(d, 99) <= 2e-3 and d.max() <= 3e-2, (np.percentile(d, [50, 99]), d.max())   # <= 2 half-ulps before the logistic
