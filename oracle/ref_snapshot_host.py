"""TEST INFRASTRUCTURE ONLY: ctypes access to oracle/_ref/libref_snapshot_host.so -- the REFERENCE's optimizer stack (Adam / ExponentialDecay / Ema
with their step kernels and serialize / deserialize), its GPUMemory <-> json binary conversion and its json / msgpack library compiled as host
code by oracle/Makefile.ref (oracle/ref_snapshot_host.cpp).  Used by tests/golden/make_snapshot_golden.py and tests/test_golden_snapshot.py."""
import ctypes as C
import json
import os

import numpy as np

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_snapshot_host.so")
_lib = None


def available():
    return os.path.exists(LIB) and os.path.isdir("/root/reference/dependencies/my_tcnn/include")


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        _lib.refs_create.restype = C.c_void_p
        _lib.refs_last_error.restype = C.c_char_p
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _bits(a, n):
    if a is None:
        return None
    b = np.zeros(n, np.float16)
    v = np.asarray(a, np.float32).reshape(-1)
    b[:v.size] = v
    return b.view(np.uint16)


class Stack:
    """One Trainer-like bundle: Ema{ExponentialDecay{Adam}} on a NerfNetwork-shaped parameter vector [density MLP | rgb MLP | grid | variance]."""

    def __init__(self, optimizer_config, n_density, n_rgb, n_grid, n_variance, layers):
        self.n = n_density + n_rgb + n_grid + n_variance
        lay = np.ascontiguousarray(np.asarray(layers, np.uint32).reshape(-1, 2))
        self.h = C.c_void_p(lib().refs_create(json.dumps(optimizer_config).encode(), C.c_uint32(n_density), C.c_uint32(n_rgb), C.c_uint32(n_grid), C.c_uint32(n_variance),
                                              _p(lay), C.c_uint32(len(lay))))
        if not self.h:
            raise RuntimeError(lib().refs_last_error().decode())

    def _ok(self, rc):
        if rc != 0:
            raise RuntimeError(lib().refs_last_error().decode())

    def close(self):
        if self.h:
            lib().refs_destroy(self.h)
            self.h = None

    def set_params_full_precision(self, master):
        m = np.ascontiguousarray(master, np.float32)
        self._ok(lib().refs_set_params_full_precision(self.h, _p(m), C.c_uint64(m.size)))

    def optimizer_step(self, loss_scale, grads_half):
        g = np.ascontiguousarray(grads_half, np.float16).view(np.uint16)
        assert g.size == self.n
        self._ok(lib().refs_optimizer_step(self.h, C.c_float(loss_scale), _p(g)))

    def state(self):
        n = self.n
        out = dict(master=np.zeros(n, np.float32), params=np.zeros(n, np.uint16), params_inference=np.zeros(n, np.uint16), m1=np.zeros(n, np.float32),
                   m2=np.zeros(n, np.float32), steps=np.zeros(n, np.uint32), ema=np.zeros(n, np.uint16))
        cs, lr, lf = C.c_uint32(), C.c_float(), C.c_float()
        self._ok(lib().refs_get_state(self.h, _p(out["master"]), _p(out["params"]), _p(out["params_inference"]), _p(out["m1"]), _p(out["m2"]), _p(out["steps"]),
                                      _p(out["ema"]), C.byref(cs), C.byref(lr), C.byref(lf)))
        out.update(current_step=int(cs.value), learning_rate=float(lr.value), learning_rate_factor=float(lf.value))
        return out

    def set_testbed_state(self, training_step, loss, rays_per_batch, measured, measured_before, aabb_scale=1, density_grid=None, rotation=None, transition=None,
                          local_rotation=None, local_transition=None):
        g = None if density_grid is None else np.ascontiguousarray(density_grid, np.float16).view(np.uint16)
        self._ok(lib().refs_set_testbed_state(self.h, C.c_uint32(training_step), C.c_float(loss), C.c_uint32(rays_per_batch), C.c_uint32(measured), C.c_uint32(measured_before),
                                              C.c_int(aabb_scale), _p(g), C.c_uint64(0 if g is None else g.size), _p(_bits(rotation, 12)), _p(_bits(transition, 4)),
                                              _p(_bits(local_rotation, 8)), _p(_bits(local_transition, 4))))

    def testbed_state(self, with_density_grid=False):
        u5 = np.zeros(5, np.uint32); loss = C.c_float(); aabb = C.c_int(); sizes = np.zeros(5, np.uint64)
        self._ok(lib().refs_get_testbed_state(self.h, _p(u5), C.byref(loss), C.byref(aabb), _p(sizes), None, None, None, None, None))
        bufs = [np.zeros(int(s), np.uint16) for s in sizes]
        self._ok(lib().refs_get_testbed_state(self.h, _p(u5), C.byref(loss), C.byref(aabb), _p(sizes), _p(bufs[0]) if with_density_grid else None, _p(bufs[1]), _p(bufs[2]),
                                              _p(bufs[3]), _p(bufs[4])))
        return dict(training_step=int(u5[0]), rays_per_batch=int(u5[1]), measured_batch_size=int(u5[2]), measured_batch_size_before_compaction=int(u5[3]), loss=float(loss.value),
                    aabb_scale=int(aabb.value), density_grid_size=int(sizes[0]), density_grid=bufs[0].view(np.float16) if with_density_grid else None,
                    rotation=bufs[1].view(np.float16), transition=bufs[2].view(np.float16), local_rotation=bufs[3].view(np.float16), local_transition=bufs[4].view(np.float16))

    def save_snapshot(self, network_config, path, include_optimizer_state=True):
        self._ok(lib().refs_save_snapshot(self.h, json.dumps(network_config).encode(), path.encode(), C.c_int(1 if include_optimizer_state else 0)))

    def load_snapshot(self, path):
        self._ok(lib().refs_load_snapshot(self.h, path.encode()))
