"""CPU pin of the oracle's mesh attributes (vertex normals, the network inputs and the colour extraction of compute_mesh_vertex_colors) against
golden vectors produced by the REFERENCE's own kernels (src/marching_cubes.cu accumulate_1ring; src/testbed_nerf.cu
generate_nerf_network_inputs_from_positions, extract_srgb_with_activation) compiled as host code from where they lie (oracle/Makefile.ref,
tests/golden/make_mesh_attributes_golden.py).  Integer content (1-ring valences) is exact; floats are compared to rounding: the oracle writes
the fused multiply-adds nvcc contracts (cross product, squared norm), the host build of the reference rounds every product."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "mesh_attributes_reference.npz")


def _mod(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, "golden", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_vertex_normals_and_color_inputs_match_reference_kernels():
    g, mod, mc = np.load(GOLD), _mod("make_mesh_attributes_golden"), _mod("make_mc_golden")
    s = int(g["stride"])
    assert int(g["n_cases"]) == len(mc.CASES)
    for k, case in enumerate(mc.CASES):
        v, f = O.marching_cubes(mc.lattice(case), np.array(case["aabb"], np.float32), case["thresh"])     # == the reference's mesh (test_golden_mc.py)
        assert (len(v), len(f)) == tuple(int(c) for c in g[f"m{k}_counts"])
        n = O.mesh_vertex_normals(v, f)
        ref = g[f"m{k}_normals"]
        scale = np.abs(ref).max()
        assert scale > 1e-4 and np.abs(n[::s] - ref).max() <= 2e-6 * scale          # sums of <= 8 face normals, each one fused multiply-add apart
        # every vertex of a closed surface lies on >= 3 triangles; valences from the reference's homogeneous weight (2 per incident triangle)
        val = np.bincount(f.reshape(-1).astype(np.int64), minlength=len(v))
        assert np.array_equal(2 * val, g[f"m{k}_valence2"].astype(np.int64))
        # normals point out of the solid (sdf > thresh inside): the big sphere of the lattice is centred at (0.5, 0.45, 0.5) of the unit lattice
        lo, hi = np.array(case["aabb"][:3], np.float32), np.array(case["aabb"][3:], np.float32)
        c = lo + (hi - lo) * np.array([0.5, 0.45, 0.5], np.float32)
        near = np.linalg.norm((v - c) / (hi - lo), axis=1) < 0.35
        # gen_faces' winding faces inwards (save_mesh reverses it when writing); (pb-pa)x(pa-pc) is MINUS that winding's normal: outwards
        assert np.mean(np.sum(n[near] * (v[near] - c), axis=1) > 0) > 0.97
        for a, box in enumerate(mod.TRAIN_AABBS):
            co = O.mesh_color_inputs(v, np.array(box, np.float32))
            ref = g[f"m{k}_coords_aabb{a}"]
            if k == 0:      # same vertices bit for bit (unit box): warped position and dt are the same operations, bit-exact
                assert np.array_equal(co[::s, :4].view(np.uint32), ref[:, :4].view(np.uint32))
            else:           # the vertices themselves are one fused multiply-add apart there (test_golden_mc.py)
                assert np.abs(co[::s, :4] - ref[:, :4]).max() <= 1.2e-7
            assert np.abs(co[::s, 4:] - ref[:, 4:]).max() <= 2.4e-7                              # direction: fused squared norm (up to 2 ulps of 0.5..1)
            assert np.all(co[:, 3] == 0)


def test_colour_extraction_matches_reference_kernel():
    g, mod = np.load(GOLD), _mod("make_mesh_attributes_golden")
    rgbd = mod.network_outputs(4096, 7)
    out16 = np.zeros((4096, 16), np.float16)
    out16[:, :7] = rgbd
    for lin in (0, 1):
        c = O.mesh_extract_colors(out16.view(np.uint16), from_linear=bool(lin))
        ref = g[f"colors_from_linear{lin}"]
        assert np.array_equal(c.view(np.uint32), ref.view(np.uint32))
        assert 0.0 <= c.min() and c.max() <= 1.0 and c.std() > 0.2


@pytest.mark.skipif(not (os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_mc_host.so")) and os.path.isdir("/root/reference/src")),
                    reason="needs the reference tree (oracle/_ref/libref_mc_host.so, libref_nerf_host.so)")
def test_committed_fixture_is_what_the_reference_source_produces(tmp_path):
    g, mod = np.load(GOLD), _mod("make_mesh_attributes_golden")
    fresh_path = str(tmp_path / "fresh.npz")
    mod.main(fresh_path)
    fresh = np.load(fresh_path)
    assert sorted(fresh.files) == sorted(g.files)
    for name in g.files:
        assert np.array_equal(fresh[name], g[name]), name


# ---- the files: save_mesh (src/marching_cubes.cu:824-960) ------------------------------------------------------------------------------------
REF_PLY_COMMENT = b"comment output from https://github.com/NVlabs/instant-ngp"


def test_obj_and_ply_files_are_byte_identical_to_the_reference_writer(tmp_path):
    """neus2_b200.mesh.save against the bytes the reference's own save_mesh wrote for the same seeded mesh (dataset coordinates (v - offset) / scale,
    reversed winding, normalised normals, clamped colours, printf formats).  The one intended difference is the PLY comment line."""
    import hashlib
    from neus2_b200 import mesh
    g, mod = np.load(os.path.join(HERE, "golden", "mesh_files_reference.npz")), _mod("make_mesh_files_golden")
    V, N, Cc, F = mod.seeded_mesh()
    for ext in ("obj", "ply"):
        path = str(tmp_path / ("mesh." + ext))
        mesh.save(path, V, N, Cc, F, nerf_scale=mod.NERF_SCALE, nerf_offset=mod.NERF_OFFSET)
        data = open(path, "rb").read()
        if ext == "ply":
            assert b"comment output from neus2_b200" in data
            data = data.replace(b"comment output from neus2_b200", REF_PLY_COMMENT)
        assert len(data) == int(g[ext + "_size"]) and data[:1500] == g[ext + "_head"].tobytes()
        assert np.array_equal(np.frombuffer(hashlib.sha256(data).digest(), np.uint8), g[ext + "_sha"])


@pytest.mark.skipif(not (os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_savemesh_host.so")) and os.path.isdir("/root/reference/src")),
                    reason="needs the reference tree (oracle/_ref/libref_savemesh_host.so)")
def test_committed_mesh_file_fixture_is_what_the_reference_source_produces(tmp_path):
    g, mod = np.load(os.path.join(HERE, "golden", "mesh_files_reference.npz")), _mod("make_mesh_files_golden")
    fresh_path = str(tmp_path / "fresh.npz")
    mod.main(fresh_path)
    fresh = np.load(fresh_path)
    assert sorted(fresh.files) == sorted(g.files)
    for name in g.files:
        assert np.array_equal(fresh[name], g[name]), name
