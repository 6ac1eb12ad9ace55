"""ctypes binding of oracle/_ref/libngp_network_ref.so -- the REFERENCE's own ngp::NerfNetwork<__half> (nerf_network.h) on its own
tiny-cuda-nn, built from /root/reference in place by `make -f oracle/Makefile.ref network` (oracle/ref_network_shim.cu).
TEST INFRASTRUCTURE ONLY; needs a GPU.  Never imported by the product.  Torch-free: device memory through libcudart (class Cuda)."""
import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libngp_network_ref.so")
_lib = None

# the four blocks of configs/nerf/base.json that reach the NerfNetwork constructor; per_level_scale as Testbed::reset_network derives it
# from top_resolution (src/testbed.cu:2160-2170)
BASE_NETWORK = {
    "encoding": {"otype": "HashGrid", "n_levels": 14, "n_features_per_level": 2, "log2_hashmap_size": 19, "base_resolution": 16,
                 "valid_level_scale": 0.02, "base_valid_level_scale": 0.2, "base_training_step": 100},
    "network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 1},
    "dir_encoding": {"otype": "Composite", "nested": [{"n_dims_to_encode": 3, "otype": "SphericalHarmonics", "degree": 4},
                                                      {"otype": "Identity", "n_bins": 4, "degree": 4}]},
    "rgb_network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 2},
}


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.refnet_create.restype = C.c_void_p
        _lib.refnet_create.argtypes = [C.c_char_p]
        _lib.refnet_last_error.restype = C.c_char_p
        _lib.refnet_n_params.restype = C.c_uint64
        _lib.refnet_n_params.argtypes = [C.c_void_p]
        _lib.refnet_destroy.argtypes = [C.c_void_p]
        _lib.refnet_set_params.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        _lib.refnet_set_training_step.argtypes = [C.c_void_p, C.c_int32]
        _lib.refnet_set_delta.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.refnet_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        _lib.refnet_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        if hasattr(_lib, "refnet_sdf"):
            _lib.refnet_sdf.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
        if hasattr(_lib, "refnet_time_step"):
            _lib.refnet_time_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_float)]
    return _lib


class Cuda:
    """The few libcudart calls the checker needs (device buffers as plain integers)."""

    def __init__(self):
        for name in ("libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so.12", "libcudart.so"):
            try:
                self.rt = C.CDLL(name)
                break
            except OSError:
                continue
        else:
            raise RuntimeError("libcudart not found")
        self.rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        self.rt.cudaFree.argtypes = [C.c_void_p]
        self.rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        self.rt.cudaMemset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
        self.rt.cudaGetErrorString.restype = C.c_char_p
        self._owned = []

    def check(self, rc):
        if rc != 0:
            raise RuntimeError("CUDA error %d: %s" % (rc, self.rt.cudaGetErrorString(C.c_int(rc)).decode()))

    def malloc(self, nbytes):
        p = C.c_void_p()
        self.check(self.rt.cudaMalloc(C.byref(p), max(int(nbytes), 1)))
        self._owned.append(p.value)
        return p.value

    def upload(self, a):
        a = np.ascontiguousarray(a)
        p = self.malloc(a.nbytes)
        self.check(self.rt.cudaMemcpy(p, a.ctypes.data_as(C.c_void_p), a.nbytes, 1))
        return p

    def zeros(self, nbytes):
        p = self.malloc(nbytes)
        self.check(self.rt.cudaMemset(p, 0, int(nbytes)))
        return p

    def memset0(self, p, nbytes):
        self.check(self.rt.cudaMemset(p, 0, int(nbytes)))

    def download(self, p, shape, dtype):
        a = np.empty(shape, dtype=dtype)
        self.check(self.rt.cudaMemcpy(a.ctypes.data_as(C.c_void_p), p, a.nbytes, 2))
        return a

    def sync(self):
        self.check(self.rt.cudaDeviceSynchronize())

    def free_all(self):
        for p in self._owned:
            self.rt.cudaFree(p)
        self._owned = []


class RefNetwork:
    def __init__(self, per_level_scale, network=None, anneal_end=0):
        cfg = json.loads(json.dumps(network or BASE_NETWORK))
        cfg["encoding"]["per_level_scale"] = float(per_level_scale)
        cfg["anneal_end"] = int(anneal_end)
        self.h = lib().refnet_create(json.dumps(cfg).encode())
        if not self.h:
            raise RuntimeError(lib().refnet_last_error().decode())
        self.n_params = int(lib().refnet_n_params(self.h))

    def _ok(self, rc):
        if rc != 0:
            raise RuntimeError(lib().refnet_last_error().decode())

    def close(self):
        if self.h:
            lib().refnet_destroy(self.h)
            self.h = None

    def set_params(self, params_half_dev, n):
        self._ok(lib().refnet_set_params(self.h, params_half_dev, n))

    def set_training_step(self, step):
        self._ok(lib().refnet_set_training_step(self.h, step))

    def set_delta(self, use_delta, train_delta, rotation_6d=None, transition=None):
        buf = None
        if rotation_6d is not None:
            d = np.zeros(12, dtype=np.float16)
            d[0:3] = np.asarray(transition, np.float32)
            d[4:10] = np.asarray(rotation_6d, np.float32)
            buf = d.ctypes.data_as(C.c_void_p)
        self._ok(lib().refnet_set_delta(self.h, int(use_delta), int(train_delta), buf))

    def forward(self, coords_dev, n, out_dev):
        self._ok(lib().refnet_forward(self.h, coords_dev, n, out_dev))

    def backward(self, coords_dev, dL_dev, n, indeed_batch_size, out_dev, grads_half_dev, want_delta=False):
        d = np.zeros(12, dtype=np.float16)
        self._ok(lib().refnet_backward(self.h, coords_dev, dL_dev, n, indeed_batch_size, out_dev, grads_half_dev,
                                       d.ctypes.data_as(C.c_void_p) if want_delta else None))
        return d.astype(np.float32)

    def time_step(self, coords_dev, dL_dev, n_fwd, n_bwd, indeed_batch_size, warmup=3, iters=20):
        """(ms of the inference pass over n_fwd samples, ms of forward + backward over n_bwd samples): the network part of one training step."""
        ms = (C.c_float * 2)()
        self._ok(lib().refnet_time_step(self.h, coords_dev, dL_dev, n_fwd, n_bwd, indeed_batch_size, warmup, iters, ms))
        return float(ms[0]), float(ms[1])

    def sdf(self, pos_dev, n, out_half_dev, as_density=False):
        """NerfNetwork::sdf / ::density with the training parameters: pos_dev [n][3] float -> out_half_dev [n] half."""
        self._ok(lib().refnet_sdf(self.h, pos_dev, n, out_half_dev, int(bool(as_density))))
