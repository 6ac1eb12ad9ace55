"""Generates tests/golden/mesh_files_reference.npz by running the REFERENCE's own save_mesh (src/marching_cubes.cu:824-960, compiled as host code from where
it lies: oracle/Makefile.ref -> oracle/_ref/libref_savemesh_host.so) on a seeded mesh: the .obj and .ply bytes it wrote (SHA-256 + the first lines).
CPU only; needs /root/reference:   make -f oracle/Makefile.ref all && python tests/golden/make_mesh_files_golden.py
"""
import ctypes as C
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_savemesh_host.so")
NERF_SCALE, NERF_OFFSET = 0.37, (0.5, 0.4, 0.6)


def seeded_mesh():
    rs = np.random.RandomState(0)
    nv = 500
    V = rs.uniform(0, 1, (nv, 3)).astype(np.float32)
    N = rs.normal(0, 1, (nv, 3)).astype(np.float32)
    N[3] = 0                                                              # a vertex no triangle touched: normalized() of zero stays zero
    Cc = rs.uniform(-0.1, 1.1, (nv, 3)).astype(np.float32)               # out-of-range colours are clamped by the writers
    F = rs.randint(0, nv, (900, 3)).astype(np.uint32)
    return V, N, Cc, F


def main(out_path):
    L = C.CDLL(LIB)
    V, N, Cc, F = seeded_mesh()
    off = np.array(NERF_OFFSET, np.float32)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for ext in ("obj", "ply"):
            path = os.path.join(d, "mesh." + ext)
            assert L.refsm_save_mesh(p(V), p(N), p(Cc), C.c_uint32(len(V)), p(F.reshape(-1)), C.c_uint32(F.size), path.encode(), C.c_float(NERF_SCALE), p(off)) == 0
            data = open(path, "rb").read()
            out[ext + "_sha"] = np.frombuffer(hashlib.sha256(data).digest(), np.uint8).copy()
            out[ext + "_size"] = np.array(len(data))
            out[ext + "_head"] = np.frombuffer(data[:1500], np.uint8).copy()
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "mesh_files_reference.npz"))
