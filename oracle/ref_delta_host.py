"""TEST INFRASTRUCTURE ONLY: ctypes access to oracle/_ref/libref_delta_host.so -- the REFERENCE's own rotation-6d / DeltaNetwork
kernels (include/neural-graphics-primitives/common_operation.cuh) compiled as host code by oracle/Makefile.ref against the stand-in
headers of oracle/ref_host_shim/.  Used by tests/test_ref_delta_host.py and tests/golden/make_delta_golden.py."""
import ctypes as C
import os

import numpy as np

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_delta_host.so")
_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
    return _lib


def _hb(a):
    return np.ascontiguousarray(np.asarray(a, np.float32).astype(np.float16).view(np.uint16))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _half_to_f32(bits):
    return bits.view(np.float16).astype(np.float32)


def rotation_matrix(rot6):
    R = np.zeros(9, np.float32)
    lib().refd_rotation_matrix(_p(_hb(rot6)), _p(R))
    return R.reshape(3, 3)


def apply(coords, rot6, trans, first_frame_offset=None):
    """[n][7] coordinates -> deformed (position [n][3], warped direction [n][3])."""
    coords = np.ascontiguousarray(coords, np.float32)
    in6 = np.ascontiguousarray(coords[:, [0, 1, 2, 4, 5, 6]])
    out6 = np.zeros_like(in6)
    off = None if first_frame_offset is None else np.ascontiguousarray(first_frame_offset, np.float32)
    lib().refd_apply(C.c_uint32(len(in6)), _p(in6), _p(_hb(rot6)), _p(_hb(trans)), _p(off), _p(out6))
    return out6[:, :3].copy(), out6[:, 3:].copy()


def backward_each(coords, dL_dpos, rot6, trans, first_frame_offset=None):
    """Per-sample gradients as (float)half: [n][9] = [d transition 3 | d rotation-6d 6], and d input [n][3]."""
    pos = np.ascontiguousarray(np.asarray(coords, np.float32)[:, :3])
    dL = np.ascontiguousarray(dL_dpos, np.float32)
    n = len(pos)
    rg, tg, ig = np.zeros((n, 6), np.uint16), np.zeros((n, 3), np.uint16), np.zeros((n, 3), np.uint16)
    off = None if first_frame_offset is None else np.ascontiguousarray(first_frame_offset, np.float32)
    lib().refd_backward_each(C.c_uint32(n), _p(dL), _p(pos), _p(_hb(rot6)), _p(_hb(trans)), _p(off), _p(rg), _p(tg), _p(ig))
    return np.concatenate([_half_to_f32(tg), _half_to_f32(rg)], axis=1), _half_to_f32(ig)


def accumulate(rot6, trans, acc_rot, acc_trans, save_only=False):
    """accumulate_global_movement_rotation_6d_kernel (or the read-only save_... kernel): (R [3][3], t [3]) as the stored halves."""
    ar = _hb(list(np.asarray(acc_rot, np.float32).reshape(9)) + [0, 0, 0])
    at = _hb(list(np.asarray(acc_trans, np.float32).reshape(3)) + [0])
    sr, st = np.zeros(12, np.uint16), np.zeros(4, np.uint16)
    lib().refd_accumulate(_p(_hb(rot6)), _p(_hb(trans)), _p(ar), _p(at), C.c_int(1 if save_only else 0), _p(sr), _p(st))
    r, t = (sr, st) if save_only else (ar, at)
    return _half_to_f32(r[:9]).reshape(3, 3), _half_to_f32(t[:3])
