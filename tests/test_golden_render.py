"""CPU pin of the oracle's inference tracer against golden vectors produced by the REFERENCE's own tracer kernels
(src/testbed_nerf.cu: init_rays_with_payload_kernel_nerf, advance_pos_nerf, compact_kernel_nerf, generate_next_nerf_network_inputs,
composite_kernel_nerf) compiled as host code from where they lie (oracle/Makefile.ref, tests/golden/make_render_golden.py), with the
oracle's network forward as the network on both sides.  Which pixels hit the occupancy grid, where their rays start and which rays reach
the shading stage are exact; ray set-up and depth agree to float rounding; colours to the network's half precision (the reference
marches several steps per compaction with plain IEEE position updates, the oracle one step with fused ones: a sample position that moves
by one ulp can move a half-precision network output by one half-ulp)."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_nerf_host as RK

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "render_reference.npz")


def _mod():
    spec = importlib.util.spec_from_file_location("make_render_golden", os.path.join(HERE, "golden", "make_render_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_tracer_matches_reference_kernels():
    g, mod = np.load(GOLD), _mod()
    s = mod.setup()
    for k, v in enumerate(s["views"]):
        rc = O.make_render_cam(mod.W, mod.H, v["focal"], v["screen_center"], v["camera"], sample_index=v["spp"], snap=v["snap"], min_transmittance=v["min_transmittance"])
        pay = O.render_init_rays(rc, s["bitfield"])
        ref_pay = g[f"v{k}_payload"]
        assert np.array_equal(pay[:, 7], ref_pay[:, 7]) and 200 < pay[:, 7].sum() < mod.W * mod.H      # the same pixels enter the occupancy grid
        alive = pay[:, 7] > 0
        assert np.abs(pay[alive] - ref_pay[alive]).max() < 5e-7                                        # origin, direction, start t
        rgba, depth, ns = O.render_trace(s["cfg"], s["params"], rc, s["bitfield"], pay)
        ref_rgba, ref_depth = g[f"v{k}_rgba"], g[f"v{k}_depth"]
        assert np.array_equal(rgba[:, 3] > 0, ref_rgba[:, 3] > 0) and int((rgba[:, 3] > 0).sum()) == int(g[f"v{k}_n_hit"])   # the same rays reach shade_kernel_nerf
        d = np.abs(rgba - ref_rgba)
        assert np.percentile(d, 99.9) < 2e-3 and d.max() < 2e-2, (np.percentile(d, [50, 99, 99.9]), d.max())
        dd = np.abs(depth - ref_depth)                                                                 # depth of the maximum-weight sample: two samples of
        assert np.percentile(dd, 99.5) < 2e-6 and dd.max() < 2e-2, (np.percentile(dd, 99.5), dd.max()) # near-equal weight may swap on a half-ulp of colour
        # the reference's payload.n_steps also counts the marching step on which a ray leaves the box or saturates
        hit = ref_rgba[:, 3] > 0
        assert np.all(np.abs(g[f"v{k}_n_steps"][hit].astype(np.int64) - ns[hit].astype(np.int64)) <= 8)


def test_mesh_lattice_positions_match_reference_kernel():
    """generate_grid_samples_nerf_uniform: where get_density_on_grid samples the SDF for marching cubes (cell corners, not centres)."""
    import ctypes as C
    g, mod = np.load(GOLD), _mod()
    for k, (res, ra, ta) in enumerate(mod.LATTICES):
        pos = np.zeros((res[2], res[1], res[0], 3), np.float32)
        O.lib().orc_lattice_positions(np.array(res, np.uint32).ctypes.data_as(C.c_void_p), np.array(ra, np.float32).ctypes.data_as(C.c_void_p),
                                      np.array(ta, np.float32).ctypes.data_as(C.c_void_p), pos.ctypes.data_as(C.c_void_p))
        ref = g[f"lattice{k}"]
        if k != 1:
            assert np.array_equal(pos.view(np.uint32), ref.view(np.uint32))         # unit / power-of-two boxes: nothing to fuse
        assert np.abs(pos - ref).max() <= 1.2e-7
        first = (np.array(ra[:3], np.float32) - np.array(ta[:3], np.float32)) / (np.array(ta[3:], np.float32) - np.array(ta[:3], np.float32))
        assert np.allclose(ref[0, 0, 0], first, atol=1e-7)                            # the first sample sits ON the render box's minimum


def test_shade_accumulate_tonemap_match_reference_kernels():
    """shade_kernel_nerf -> accumulate_kernel -> tonemap_kernel (linear colour space, identity curve): tests/golden/shade_reference.npz"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_shade_golden", os.path.join(HERE, "golden", "make_shade_golden.py"))
    sm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sm)
    g = np.load(os.path.join(HERE, "golden", "shade_reference.npz"))
    rgba, depth, accs = sm.inputs()
    for k, (sample_count, to_srgb, bg, expo) in enumerate(sm.CASES):
        acc = accs[k].copy()
        img, dout = O.render_shade_accumulate_tonemap(rgba, depth, acc, sample_count, bg, exposure=expo, to_srgb=bool(to_srgb))
        assert np.array_equal(dout, g[f"s{k}_depth"])                                  # depth only where alpha > 0.2, 1e10 elsewhere
        assert np.abs(acc - g[f"s{k}_accumulate"]).max() <= 1.2e-7                     # running mean: one fused multiply-add in the oracle
        assert np.abs(img - g[f"s{k}_image"]).max() <= 5e-7
        if sample_count == 0:
            assert np.array_equal(img.view(np.uint32), g[f"s{k}_image"].view(np.uint32))   # first sample: nothing to fuse, bit-exact incl. the sRGB curve


@pytest.mark.skipif(not RK.available(), reason="needs the reference tree (oracle/_ref/libref_nerf_host.so)")
def test_committed_fixture_is_what_the_reference_source_produces(tmp_path):
    g, mod = np.load(GOLD), _mod()
    fresh_path = str(tmp_path / "fresh.npz")
    mod.main(fresh_path)
    fresh = np.load(fresh_path)
    assert sorted(fresh.files) == sorted(g.files)
    for name in g.files:
        assert np.array_equal(fresh[name], g[name]), name
    spec = importlib.util.spec_from_file_location("make_shade_golden", os.path.join(HERE, "golden", "make_shade_golden.py"))
    sm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sm)
    sm.main(fresh_path)
    fresh, g = np.load(fresh_path), np.load(os.path.join(HERE, "golden", "shade_reference.npz"))
    for name in g.files:
        assert np.array_equal(fresh[name], g[name]), name
