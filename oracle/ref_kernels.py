"""ctypes binding of oracle/_ref/libtcnn_ref.so -- the REFERENCE's own tiny-cuda-nn kernels (hash grid + fused MLP),
built from /root/reference in place by oracle/Makefile.ref.  TEST INFRASTRUCTURE ONLY (parity pin + reference-CUDA
timing baseline); needs a GPU.  Never imported by the product."""
import ctypes as C
import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libtcnn_ref.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_grid_create.restype = C.c_void_p
        _lib.ref_mlp_create.restype = C.c_void_p
        _lib.ref_grid_n_params.restype = C.c_uint64
        _lib.ref_mlp_n_params.restype = C.c_uint64
        _lib.ref_last_error.restype = C.c_char_p
        for f in ("ref_grid_n_params", "ref_grid_out_width", "ref_grid_destroy", "ref_mlp_n_params", "ref_mlp_destroy"):
            getattr(_lib, f).argtypes = [C.c_void_p]
        _lib.ref_grid_set_training_step.argtypes = [C.c_void_p, C.c_int]
    return _lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class RefGrid:
    def __init__(self, n_levels=14, log2_hashmap_size=19, base_resolution=16, per_level_scale=1.4524229):
        cfg = dict(otype="HashGrid", n_levels=n_levels, n_features_per_level=2, log2_hashmap_size=log2_hashmap_size,
                   base_resolution=base_resolution, per_level_scale=per_level_scale, valid_level_scale=0.02,
                   base_valid_level_scale=0.2, base_training_step=100)
        self.h = C.c_void_p(lib().ref_grid_create(json.dumps(cfg).encode()))
        if not self.h:
            raise RuntimeError(lib().ref_last_error().decode())
        self.n_params = int(lib().ref_grid_n_params(self.h))
        self.E = 2 * n_levels

    def set_training_step(self, step):
        lib().ref_grid_set_training_step(self.h, step)

    def eval(self, coords, params_h, dL_dy_soa=None, v=None, dL_dy2_soa=None, want_grad=True, iters=1, which=7):
        """coords: cuda float32 [n,stride]; params_h: cuda half [n_params]. Returns dict of cuda tensors + ms."""
        import torch
        n, stride = coords.shape
        dev = coords.device
        enc = torch.zeros((self.E, n), dtype=torch.float16, device=dev)
        dL_dx = torch.zeros((n, 3), dtype=torch.float32, device=dev) if dL_dy_soa is not None else None
        grad = torch.zeros(self.n_params, dtype=torch.float16, device=dev) if want_grad else None
        ddy = torch.zeros((self.E, n), dtype=torch.float16, device=dev) if v is not None else None
        ms = C.c_float(0)
        rc = lib().ref_grid_eval(self.h, C.c_uint32(n), _p(coords), C.c_uint32(stride), _p(params_h), _p(enc), _p(dL_dy_soa), _p(dL_dx), _p(grad),
                                 _p(v), _p(dL_dy2_soa), _p(ddy), C.c_int(iters), C.c_int(which), C.byref(ms))
        if rc != 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return dict(enc=enc, dL_dx=dL_dx, grad=grad, dL_ddLdy=ddy, ms=ms.value)


class RefMlp:
    def __init__(self, n_in, n_hidden_layers, n_out=16):
        cfg = dict(otype="FullyFusedMLP", activation="ReLU", output_activation="None", n_neurons=64, n_hidden_layers=n_hidden_layers,
                   n_input_dims=n_in, n_output_dims=n_out)
        self.h = C.c_void_p(lib().ref_mlp_create(json.dumps(cfg).encode()))
        if not self.h:
            raise RuntimeError(lib().ref_last_error().decode())
        self.n_params = int(lib().ref_mlp_n_params(self.h))
        self.n_in = n_in

    def eval(self, x, params_h, in_soa=False, dL_dout=None, want_din=True, want_wgrad=True, d_in2=None, seed=None, iters=1, which=7):
        """x: cuda half, [n,in] (or [in,n] if in_soa). Returns dict(out [n,16], dL_din (layout of x), wgrad half [n_params])."""
        import torch
        n = x.shape[1] if in_soa else x.shape[0]
        dev = x.device
        out = torch.zeros((n, 16), dtype=torch.float16, device=dev)
        din = torch.zeros_like(x) if (dL_dout is not None and want_din) else None
        wgrad = torch.zeros(self.n_params, dtype=torch.float16, device=dev) if want_wgrad else None
        ms = C.c_float(0)
        rc = lib().ref_mlp_eval(self.h, C.c_uint32(n), C.c_uint32(self.n_in), C.c_int(int(in_soa)), _p(x), _p(params_h), _p(out), _p(dL_dout), _p(din),
                                _p(wgrad), _p(d_in2), _p(seed), C.c_int(iters), C.c_int(which), C.byref(ms))
        if rc != 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return dict(out=out, dL_din=din, wgrad=wgrad, ms=ms.value)


class RefOptimizer:
    """The reference's Ema{ExponentialDecay{Adam}} as configs/nerf/base.json builds it (tcnn optimizers/{ema,exponential_decay,adam}.h),
    over `n_params` parameters of which the first `n_matrix` are matrix weights."""

    def __init__(self, n_params, n_matrix, learning_rate=1e-3, beta1=0.9, beta2=0.99, epsilon=1e-15, l2_reg=1e-6, ema_decay=0.95):
        cfg = dict(otype="Ema", decay=ema_decay, nested=dict(otype="ExponentialDecay", decay_start=20000, decay_interval=10000, decay_base=0.33,
                   nested=dict(otype="Adam", learning_rate=learning_rate, beta1=beta1, beta2=beta2, epsilon=epsilon, l2_reg=l2_reg)))
        lib().ref_opt_create.restype = C.c_void_p
        self.h = C.c_void_p(lib().ref_opt_create(json.dumps(cfg).encode(), C.c_uint64(n_params), C.c_uint64(n_matrix)))
        if not self.h:
            raise RuntimeError(lib().ref_last_error().decode())
        self.n = n_params

    def step(self, loss_scale, master, params_h, grads_h):
        """master: cuda float32, params_h / grads_h: cuda half (updated in place).  Returns the EMA weights as a cuda half tensor copy."""
        import torch
        ema_ptr = C.c_void_p()
        rc = lib().ref_opt_step(self.h, C.c_float(loss_scale), _p(master), _p(params_h), _p(grads_h), C.byref(ema_ptr))
        if rc != 0:
            raise RuntimeError(lib().ref_last_error().decode())
        out = torch.empty(self.n, dtype=torch.float16, device=master.device)
        torch.cuda.synchronize()
        # device-to-device copy of the optimizer-owned EMA buffer through torch's raw-pointer view
        class _Dev:
            def __init__(s, ptr, n):
                s.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f2", "data": (int(ptr), False), "version": 3, "strides": None}
        out.copy_(torch.as_tensor(_Dev(ema_ptr.value, self.n), device=master.device))
        return out

    def close(self):
        if self.h:
            lib().ref_opt_destroy.argtypes = [C.c_void_p]
            lib().ref_opt_destroy(self.h)
            self.h = None


def morton3D(x, y, z):
    lib().ref_morton3D.restype = C.c_uint32
    return int(lib().ref_morton3D(C.c_uint32(x), C.c_uint32(y), C.c_uint32(z)))


def morton3D_invert(x):
    lib().ref_morton3D_invert.restype = C.c_uint32
    return int(lib().ref_morton3D_invert(C.c_uint32(x)))


def pcg32_sequence(seed, advance, n):
    import numpy as np
    u = np.zeros(n, dtype=np.uint32)
    f = np.zeros(n, dtype=np.float32)
    lib().ref_pcg32_sequence(C.c_uint64(seed), C.c_int64(advance), C.c_uint32(n), u.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p))
    return u, f


def generate_random_uniform(seed, n, lower, upper):
    """Returns (values float32 [n] as numpy, rng state after as (state, inc))."""
    import numpy as np
    import torch
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    st = np.zeros(2, dtype=np.uint64)
    rc = lib().ref_generate_random_uniform(C.c_uint64(seed), C.c_uint64(n), _p(out), C.c_float(lower), C.c_float(upper), st.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError(lib().ref_last_error().decode())
    return out.cpu().numpy(), (int(st[0]), int(st[1]))


def fill_rollover(n_elements, n_input, dloss_h=None, coords=None):
    """In place on cuda tensors: dloss_h half [n_elements,16] (wrapped copies rescaled by n_input/n_elements), coords float32 [n_elements,7]."""
    import torch
    n_in = torch.tensor([n_input], dtype=torch.int32, device="cuda")
    rc = lib().ref_fill_rollover(C.c_uint32(n_elements), _p(n_in), _p(dloss_h), _p(coords))
    if rc != 0:
        raise RuntimeError(lib().ref_last_error().decode())
