"""Generates tests/golden/ld_reference.npz from the REFERENCE's own low-discrepancy / random helpers
(include/neural-graphics-primitives/random_val.cuh and dependencies/my_tcnn/dependencies/pcg32/pcg32.h, compiled as host code from where
they lie into oracle/_ref/libref_random_host.so by oracle/Makefile.ref).  CPU only; needs /root/reference:

    make -f oracle/Makefile.ref all && python tests/golden/make_ld_golden.py tests/golden/ld_reference.npz
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_random_host.so")


def main(out_path):
    L = C.CDLL(LIB)
    L.refr_sobol.restype = C.c_uint32
    L.refr_scramble.restype = C.c_uint32
    L.refr_ld_random_val.restype = C.c_float
    L.refr_ld_random_val.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    rs = np.random.RandomState(4242)
    idx = np.concatenate([np.arange(64, dtype=np.uint32), rs.randint(0, 2 ** 32, 448, dtype=np.uint64).astype(np.uint32)])
    seeds = np.array([0, 1, 0xdeadbeef, 0x1234567], dtype=np.uint32)
    sob = np.array([[L.refr_sobol(int(i), d) for d in (0, 1)] for i in idx], dtype=np.uint32)
    scr = np.array([[L.refr_scramble(int(i), int(s)) for s in seeds] for i in idx], dtype=np.uint32)
    ld = np.array([[L.refr_ld_random_val(int(i), int(s), 0) for s in seeds] for i in idx], dtype=np.float32)
    out2 = (C.c_float * 2)()
    spps = np.concatenate([np.arange(256, dtype=np.uint32), rs.randint(0, 2 ** 31, 64).astype(np.uint32)])
    off = []
    for spp in spps:
        L.refr_ld_random_pixel_offset(C.c_uint32(int(spp)), out2)
        off.append((out2[0], out2[1]))
    n = 96
    streams = []
    for seed, adv in [(1337, 0), (1337, 8 * 4095), (1337, 1 << 32), (7, -5), (0xffffffffffff, 123456789)]:
        f, u = np.zeros(n, np.float32), np.zeros(n, np.uint32)
        L.refr_pcg32_stream(C.c_uint64(seed), C.c_int64(adv), C.c_uint32(n), f.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p))
        streams.append((seed, adv, f, u))
    np.savez_compressed(out_path, idx=idx, seeds=seeds, sobol=sob, scramble=scr, ld_random_val=ld, spps=spps, pixel_offset=np.array(off, np.float32),
                        pcg_seed=np.array([s[0] for s in streams], np.uint64), pcg_advance=np.array([s[1] for s in streams], np.int64),
                        pcg_floats=np.stack([s[2] for s in streams]), pcg_uints=np.stack([s[3] for s in streams]))
    print("wrote", out_path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "ld_reference.npz"))
