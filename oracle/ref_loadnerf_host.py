"""TEST INFRASTRUCTURE ONLY: ctypes access to oracle/_ref/libref_loadnerf_host.so -- the REFERENCE's own dataset loader (src/nerf_loader.cu:
load_nerf + NerfDataset::set_training_image with its convert_rgba32 kernel, src/thread_pool.cpp, and the reference's json / filesystem / stb_image
libraries) compiled as host code by oracle/Makefile.ref.  Used by tests/golden/make_loader_golden.py and tests/test_golden_loader.py."""
import contextlib
import ctypes as C
import os
import sys

import numpy as np

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_loadnerf_host.so")
_lib = None


def available():
    return os.path.exists(LIB) and os.path.isdir("/root/reference/src")


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        _lib.refld_load.restype = C.c_void_p
        _lib.refld_last_error.restype = C.c_char_p
    return _lib


@contextlib.contextmanager
def _quiet():
    """The reference's tinylogger draws a progress bar on stdout: send the C-level stdout to /dev/null for the duration of the call."""
    sys.stdout.flush()
    saved, null = os.dup(1), os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(null, 1)
        yield
    finally:
        os.dup2(saved, 1)
        os.close(saved); os.close(null)


def load_nerf(json_paths, sharpen_amount=0.0):
    """Runs the reference's load_nerf on the given transforms json files.  Returns the same dict layout as neus2_b200.loader.load_nerf."""
    paths = "\n".join(os.path.abspath(p) for p in json_paths).encode()
    with _quiet():
        h = lib().refld_load(paths, C.c_float(sharpen_amount))
    if not h:
        raise RuntimeError(lib().refld_last_error().decode())
    h = C.c_void_p(h)
    try:
        n, sc, off, ab, na, mi = C.c_uint32(), C.c_float(), (C.c_float * 3)(), C.c_int(), C.c_int(), C.c_int()
        lib().refld_info(h, C.byref(n), C.byref(sc), off, C.byref(ab), C.byref(na), C.byref(mi))
        focal, principal, xforms, pixels, res_all = [], [], [], [], []
        for i in range(n.value):
            res, f, pp, x, t = (C.c_int * 2)(), (C.c_float * 2)(), (C.c_float * 2)(), (C.c_float * 12)(), C.c_int()
            assert lib().refld_image(h, C.c_uint32(i), res, f, pp, x, C.byref(t), None, C.c_uint64(0)) == 0
            px = np.zeros((res[1], res[0], 4), np.uint8)
            assert t.value == 1, "EImageDataType::Byte expected"          # common.h: None 0, Byte 1, Half 2, Float 3
            assert lib().refld_image(h, C.c_uint32(i), res, f, pp, x, C.byref(t), px.ctypes.data_as(C.c_void_p), C.c_uint64(px.nbytes)) == 0
            focal.append(list(f)); principal.append(list(pp)); xforms.append(np.array(x, np.float32).reshape(3, 4)); pixels.append(px); res_all.append((res[0], res[1]))
        return dict(pixels=pixels, focal=np.array(focal, np.float32), principal=np.array(principal, np.float32), xforms=np.stack(xforms), scale=float(sc.value),
                    offset=np.array(off, np.float32), aabb_scale=int(ab.value), from_na=bool(na.value), from_mitsuba=bool(mi.value), resolution=res_all)
    finally:
        lib().refld_free(h)
