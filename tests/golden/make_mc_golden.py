"""Generates tests/golden/mc_reference.npz by running the REFERENCE's own marching-cubes kernels (src/marching_cubes.cu: gen_vertices
:276-329, gen_faces :377-719 with its case tables) compiled as host code from where they lie (oracle/Makefile.ref ->
oracle/_ref/libref_mc_host.so), launched count-then-emit as marching_cubes_gpu does (:794-822), simulated threads in lattice order.
CPU only; needs /root/reference:

    make -f oracle/Makefile.ref all && python tests/golden/make_mc_golden.py tests/golden/mc_reference.npz

The lattices are rebuilt from seeds by the test (function below); vertices and indices are stored as SHA-256 digests plus counts and probes.
"""
import ctypes as C
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_mc_host.so")

CASES = [dict(res=(32, 48, 16), aabb=(0, 0, 0, 1, 1, 1), thresh=0.0, seed=0), dict(res=(16, 16, 32), aabb=(0.1, 0.2, 0.0, 0.9, 0.8, 1.0), thresh=0.013, seed=1)]


def lattice(case):
    rx, ry, rz = case["res"]
    z, y, x = np.meshgrid(np.arange(rz), np.arange(ry), np.arange(rx), indexing="ij")
    P = np.stack([(x + 0.5) / rx, (y + 0.5) / ry, (z + 0.5) / rz], -1)
    rs = np.random.RandomState(case["seed"])
    d = (0.3 - np.linalg.norm(P - np.array([0.5, 0.45, 0.5]), axis=-1)).astype(np.float32) + (rs.rand(rz, ry, rx).astype(np.float32) - 0.5) * 0.02
    return np.ascontiguousarray(np.maximum(d, (0.12 - np.linalg.norm(P - np.array([0.8, 0.8, 0.5]), axis=-1)).astype(np.float32)))


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8).copy()


def main(out_path):
    L = C.CDLL(LIB)
    out = {"n_cases": np.array(len(CASES))}
    for k, case in enumerate(CASES):
        d = lattice(case)
        aabb = np.array(case["aabb"], np.float32)
        verts = np.zeros((200000, 3), np.float32); idx = np.zeros(1200000, np.uint32); n = np.zeros(2, np.uint32)
        rx, ry, rz = case["res"]
        L.refm_marching_cubes(aabb.ctypes.data_as(C.c_void_p), rx, ry, rz, C.c_float(case["thresh"]), d.ctypes.data_as(C.c_void_p), verts.ctypes.data_as(C.c_void_p),
                              C.c_uint32(len(verts)), idx.ctypes.data_as(C.c_void_p), C.c_uint32(len(idx)), n.ctypes.data_as(C.c_void_p))
        nv, ni = int(n[0]), int(n[1])
        out[f"m{k}_counts"] = n.copy()
        out[f"m{k}_verts_sha"] = digest(verts[:nv]); out[f"m{k}_idx_sha"] = digest(idx[:ni])
        out[f"m{k}_verts_head"] = verts[:64].copy(); out[f"m{k}_idx_head"] = idx[:192].copy()
        if k > 0:
            out[f"m{k}_verts"] = verts[:nv].copy()      # non-unit box: positions are compared to rounding (min + x * scale: plain IEEE here, fused in the oracle)
        print("case", k, "verts", nv, "triangles", ni // 3)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "mc_reference.npz"))
