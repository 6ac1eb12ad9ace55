"""Generates tests/golden/shade_reference.npz by running the REFERENCE's own frame kernels: shade_kernel_nerf (src/testbed_nerf.cu
:2148-2181), accumulate_kernel and tonemap_kernel (src/render_buffer.cu :217-252, :474-501), compiled as host code from where they lie
(oracle/Makefile.ref -> oracle/_ref/libref_nerf_host.so, libref_rb_host.so).  CPU only; needs /root/reference:

    make -f oracle/Makefile.ref all && python tests/golden/make_shade_golden.py tests/golden/shade_reference.npz
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
W, H = 24, 10
CASES = [(0.0, 1, (0, 0, 0, 1), 0.0), (3.0, 0, (0.2, 0.4, 0.6, 1.0), 0.5), (1.0, 1, (1, 1, 1, 0.5), -1.0)]   # sample_count, to_srgb, background, exposure


def inputs():
    rs = np.random.RandomState(0)
    n = W * H
    rgba = rs.rand(n, 4).astype(np.float32)
    rgba[::3] = 0
    rgba[:, 3] = np.where(rgba[:, 3] < 0.1, 0, rgba[:, 3])
    depth = rs.rand(n).astype(np.float32) * 3
    acc = [rs.rand(n, 4).astype(np.float32) if c[0] > 0 else np.zeros((n, 4), np.float32) for c in CASES]
    return rgba, depth, acc


def main(out_path):
    LK = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_nerf_host.so"))
    LB = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_rb_host.so"))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rgba, depth, accs = inputs()
    n = W * H
    out = {}
    for k, (sample_count, to_srgb, bg, expo) in enumerate(CASES):
        frame = np.zeros((n, 4), np.float32); dbuf = np.full(n, 1e10, np.float32)
        LK.refk_shade(C.c_uint32(n), p(rgba), p(depth), p(frame), p(dbuf))
        acc = accs[k].copy(); img = np.zeros((n, 4), np.float32); bga = np.array(bg, np.float32)
        LB.refb_accumulate_tonemap(W, H, p(frame), p(acc), C.c_float(sample_count), p(bga), C.c_float(expo), to_srgb, p(img))
        out[f"s{k}_image"], out[f"s{k}_accumulate"], out[f"s{k}_depth"] = img, acc, dbuf
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "shade_reference.npz"))
