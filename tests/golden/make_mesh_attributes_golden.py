"""Generates tests/golden/mesh_attributes_reference.npz by running the REFERENCE's own mesh-attribute kernels compiled as host code from where
they lie (oracle/Makefile.ref): accumulate_1ring as compute_mesh_1ring launches it (src/marching_cubes.cu:331-366, 720-727; libref_mc_host.so) on the
meshes of make_mc_golden.py's lattices, and compute_mesh_vertex_colors' kernels generate_nerf_network_inputs_from_positions / extract_srgb_with_activation
(src/testbed_nerf.cu:847-853, 532-545; libref_nerf_host.so).  CPU only; needs /root/reference:

    make -f oracle/Makefile.ref all && python tests/golden/make_mesh_attributes_golden.py tests/golden/mesh_attributes_reference.npz

Meshes are rebuilt by the test from the same seeded lattices (the oracle's marching cubes is pinned bit for bit to the reference's by
tests/test_golden_mc.py); stored here: every 5th vertex normal, all 1-ring valences, every 5th network input, the colours of seeded outputs.
"""
import ctypes as C
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
MC_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_mc_host.so")
NERF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_nerf_host.so")
TRAIN_AABBS = [(0, 0, 0, 1, 1, 1), (-0.5, -0.5, -0.5, 1.5, 1.5, 1.5)]     # aabb_scale 1 and 2
STRIDE = 5


def _mc():
    spec = importlib.util.spec_from_file_location("make_mc_golden", os.path.join(HERE, "make_mc_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def reference_mesh(L, mc, case):
    d = mc.lattice(case)
    aabb = np.array(case["aabb"], np.float32)
    verts = np.zeros((200000, 3), np.float32); idx = np.zeros(1200000, np.uint32); n = np.zeros(2, np.uint32)
    rx, ry, rz = case["res"]
    L.refm_marching_cubes(_p(aabb), rx, ry, rz, C.c_float(case["thresh"]), _p(d), _p(verts), C.c_uint32(len(verts)), _p(idx), C.c_uint32(len(idx)), _p(n))
    return np.ascontiguousarray(verts[:int(n[0])]), np.ascontiguousarray(idx[:int(n[1])])


def network_outputs(n, seed):
    """Seeded stand-ins for the network's colour outputs (pre-activation), [n][7] float as Network::inference hands them to the kernel;
    values are half-representable so that the half [n][16] layout of the product / oracle carries the same numbers."""
    rs = np.random.RandomState(seed)
    return rs.normal(0, 2.5, (n, 7)).astype(np.float16).astype(np.float32)


def main(out_path):
    L, K, mc = C.CDLL(MC_LIB), C.CDLL(NERF_LIB), _mc()
    out = {"n_cases": np.array(len(mc.CASES)), "stride": np.array(STRIDE)}
    for k, case in enumerate(mc.CASES):
        v, idx = reference_mesh(L, mc, case)
        nv, nt = len(v), len(idx) // 3
        pos4 = np.zeros((nv, 4), np.float32); nrm = np.zeros((nv, 3), np.float32)
        L.refm_compute_mesh_1ring(C.c_uint32(nv), C.c_uint32(nt), _p(idx), _p(v), _p(pos4), _p(nrm))
        out[f"m{k}_counts"] = np.array([nv, nt], np.uint32)
        out[f"m{k}_normals"] = nrm[::STRIDE].copy()
        out[f"m{k}_valence2"] = pos4[:, 3].astype(np.uint8)                     # 2 per incident triangle
        out[f"m{k}_ring_sum"] = pos4[::STRIDE, :3].copy()
        for a, box in enumerate(TRAIN_AABBS):
            coords = np.zeros((nv, 7), np.float32)
            K.refk_mesh_color_inputs(C.c_uint32(nv), _p(np.array(box, np.float32)), _p(v), _p(coords))
            out[f"m{k}_coords_aabb{a}"] = coords[::STRIDE].copy()
        print("case", k, "verts", nv, "triangles", nt)
    n = 4096
    rgbd = network_outputs(n, 7)
    for lin in (0, 1):
        rgb = np.zeros((n, 3), np.float32)
        K.refk_extract_srgb(C.c_uint32(n), _p(rgbd), lin, _p(rgb))
        out[f"colors_from_linear{lin}"] = rgb
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "mesh_attributes_reference.npz"))
