"""CPU pin of the oracle's marching cubes against golden vectors produced by the REFERENCE's own kernels (src/marching_cubes.cu:
gen_vertices, gen_faces and their case tables) compiled as host code from where they lie (oracle/Makefile.ref,
tests/golden/make_mc_golden.py).  Run with its threads in lattice order the reference numbers vertices and triangles exactly as the
oracle's (and the product's) deterministic scan does: index buffers are bit-identical in the same order, vertex positions bit-identical
for the unit box and within one ulp otherwise (one fused multiply-add)."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "mc_reference.npz")


def _mod():
    spec = importlib.util.spec_from_file_location("make_mc_golden", os.path.join(HERE, "golden", "make_mc_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_marching_cubes_matches_reference_kernels_bit_for_bit():
    g, mod = np.load(GOLD), _mod()
    assert int(g["n_cases"]) == len(mod.CASES)
    for k, case in enumerate(mod.CASES):
        v, f = O.marching_cubes(mod.lattice(case), np.array(case["aabb"], np.float32), case["thresh"])
        assert (len(v), f.size) == tuple(int(c) for c in g[f"m{k}_counts"]) and len(v) > 500
        assert np.array_equal(f.reshape(-1)[:192], g[f"m{k}_idx_head"])
        assert np.array_equal(mod.digest(f.reshape(-1).astype(np.uint32)), g[f"m{k}_idx_sha"])            # every triangle, same order, same winding
        if k == 0:      # unit box: box_min + x * scale has nothing to fuse, positions are bit-identical
            assert np.array_equal(v[:64].view(np.uint32), g[f"m{k}_verts_head"].view(np.uint32))
            assert np.array_equal(mod.digest(v), g[f"m{k}_verts_sha"])
        else:           # the oracle mirrors nvcc's fused multiply-add there, the host build of the reference rounds twice
            assert np.abs(v - g[f"m{k}_verts"]).max() <= 1.2e-7


@pytest.mark.skipif(not (os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_mc_host.so")) and os.path.isdir("/root/reference/src")),
                    reason="needs the reference tree (oracle/_ref/libref_mc_host.so)")
def test_committed_fixture_is_what_the_reference_source_produces(tmp_path):
    g, mod = np.load(GOLD), _mod()
    fresh_path = str(tmp_path / "fresh.npz")
    mod.main(fresh_path)
    fresh = np.load(fresh_path)
    assert sorted(fresh.files) == sorted(g.files)
    for name in g.files:
        assert np.array_equal(fresh[name], g[name]), name
