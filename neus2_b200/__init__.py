"""neus2_b200 -- B200-native (sm_100a) NeuS2 training / rendering hot path.

Python is a thin host mirror of the C ABI (include/neus2_b200.h); the compute lives in
neus2_b200/libneus2_b200.so (hand-written CUDA).  PyTorch is used for device memory, streams and
torch.distributed only.  There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from ._lib import Config, Neus2Error, RenderParams, StepStats, check, load, LIB_PATH  # noqa: F401

__all__ = ["Config", "Context", "Neus2Error", "RenderParams", "StepStats", "default_config"]


def default_config(**overrides):
    cfg = Config()
    load().neus2_default_config(C.byref(cfg))
    for k, v in overrides.items():
        if k == "background_color":
            for i in range(3):
                cfg.background_color[i] = float(v[i])
        else:
            setattr(cfg, k, v)
    return cfg


class _DevArray:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can view it without a copy."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}
        self._owner = owner


def _torch():
    import torch
    return torch


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(stream):
    if stream is None:
        return C.c_void_p(_torch().cuda.current_stream().cuda_stream)
    return C.c_void_p(stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream))


class Context:
    """One NeuS2 model + trainer state on one GPU (the reference's Testbed in Nerf mode + Trainer + NerfNetwork)."""

    def __init__(self, cfg=None, device=0):
        self._lib = load()
        self.cfg = cfg if cfg is not None else default_config()
        self.device = device
        h = C.c_void_p()
        check(self._lib.neus2_ctx_create(C.byref(self.cfg), C.c_int(device), C.byref(h)))
        self._h = h
        self.peers_attached = False
        self.global_movement_active = False   # m_use_delta (mirrors the context's flag for the data-parallel step)
        self.n_params = int(self._lib.neus2_n_params(self._h))
        lay = (C.c_uint32 * 8)()
        check(self._lib.neus2_param_layout(self._h, lay))
        self.layout = dict(zip(["sdf_in", "W0", "W1", "C0", "C1", "C2", "grid", "variance"], [int(x) for x in lay]))

    def close(self):
        if getattr(self, "_h", None):
            self._buffers = None
            self.__dict__.pop("_counts_views", None)
            self.__dict__.pop("_loss_views", None)
            self._lib.neus2_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- tables / buffers --------------------------------------------------------------------
    def grid_tables(self):
        L = self.cfg.n_levels
        off = np.zeros(L + 1, dtype=np.uint32)
        res = np.zeros(L, dtype=np.uint32)
        sc = np.zeros(L, dtype=np.float32)
        check(self._lib.neus2_grid_tables(self._h, off.ctypes.data_as(C.c_void_p), res.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p)))
        return off, res, sc

    def buffers(self):
        """(master fp32, params fp16, ema fp16, grads fp32) as zero-copy torch views of the context's device buffers."""
        if getattr(self, "_buffers", None) is not None:
            return self._buffers
        torch = _torch()
        m, p, e, g = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(self._lib.neus2_buffers(self._h, C.byref(m), C.byref(p), C.byref(e), C.byref(g)))
        dev = f"cuda:{self.device}"
        n = self.n_params
        self._buffers = (torch.as_tensor(_DevArray(m.value, (n,), "<f4", None), device=dev), torch.as_tensor(_DevArray(p.value, (n,), "<f2", None), device=dev),
                         torch.as_tensor(_DevArray(e.value, (n,), "<f2", None), device=dev), torch.as_tensor(_DevArray(g.value, (n,), "<f4", None), device=dev))
        return self._buffers

    def debug_buffer(self, name, typestr, shape):
        torch = _torch()
        p, nb = C.c_void_p(), C.c_uint64()
        check(self._lib.neus2_debug_buffer(self._h, name.encode(), C.byref(p), C.byref(nb)))
        count = int(np.prod(shape))
        assert count * int(typestr[-1]) <= nb.value, (name, shape, nb.value)
        return torch.as_tensor(_DevArray(p.value, shape, typestr, self), device=f"cuda:{self.device}")

    # ---- parameters --------------------------------------------------------------------------
    def init_params(self, sdf_mlp_weights=None):
        w = None
        if sdf_mlp_weights is not None:
            w = np.ascontiguousarray(sdf_mlp_weights, dtype=np.float32)
            assert w.size == 64 * 32 + 16 * 64
        check(self._lib.neus2_init_params(self._h, w.ctypes.data_as(C.c_void_p) if w is not None else None))

    def set_params(self, master, reset_optimizer=True):
        m = np.ascontiguousarray(master, dtype=np.float32)
        check(self._lib.neus2_set_params(self._h, m.ctypes.data_as(C.c_void_p), C.c_uint64(m.size), C.c_int(1 if reset_optimizer else 0)))

    def _require_whole_optimizer_state(self, what):
        """With the sharded optimizer a rank's fp32 master copy / moments / step counters are current only for its own slice."""
        if self.peers_attached and getattr(self, "_opt_synced_step", None) != self.get_state()["optimizer_step"]:
            raise Neus2Error(f"{what}: the optimizer state of this rank is current only for its own parameter slice "
                             "(sharded optimizer); call neus2_b200.parallel.sync_optimizer_state on every rank first")

    def get_params(self):
        self._require_whole_optimizer_state("get_params")
        m = np.empty(self.n_params, dtype=np.float32)
        check(self._lib.neus2_get_params(self._h, m.ctypes.data_as(C.c_void_p), C.c_uint64(m.size)))
        return m

    def set_training_step(self, step):
        check(self._lib.neus2_set_training_step(self._h, C.c_int32(step)))

    # ---- operators ---------------------------------------------------------------------------
    def forward(self, coords, out=None, use_inference_params=False, stream=None):
        torch = _torch()
        assert coords.dtype == torch.float32 and coords.is_contiguous() and coords.shape[1] == 7
        n = coords.shape[0]
        if out is None:
            out = torch.empty((n, 16), dtype=torch.float16, device=coords.device)
        check(self._lib.neus2_forward(self._h, _stream(stream), _ptr(coords), C.c_uint32(n), _ptr(out), C.c_int(int(use_inference_params))))
        return out

    def sdf(self, pos, as_density=False, use_inference_params=False, stream=None):
        torch = _torch()
        assert pos.dtype == torch.float32 and pos.is_contiguous()
        n, stride = pos.shape
        out = torch.empty((n,), dtype=torch.float16, device=pos.device)
        check(self._lib.neus2_sdf(self._h, _stream(stream), _ptr(pos), C.c_uint32(stride), C.c_uint32(n), _ptr(out), C.c_int(int(as_density)),
                                  C.c_int(int(use_inference_params))))
        return out

    def backward(self, coords, dL_dout, n_in, n_rows, indeed_batch_size, stream=None):
        torch = _torch()
        assert coords.dtype == torch.float32 and dL_dout.dtype == torch.float16
        n_in_dev = torch.tensor([n_in], dtype=torch.int32, device=coords.device)
        check(self._lib.neus2_backward(self._h, _stream(stream), _ptr(coords), _ptr(dL_dout), _ptr(n_in_dev), C.c_uint32(n_rows), C.c_uint32(indeed_batch_size)))
        return self.buffers()[3]

    def optimizer_step(self, loss_scale=128.0, stream=None):
        check(self._lib.neus2_optimizer_step(self._h, _stream(stream), C.c_float(loss_scale)))

    # ---- training ----------------------------------------------------------------------------
    def set_dataset(self, pixels, focal, principal, xforms):
        px = np.ascontiguousarray(pixels, dtype=np.uint8)
        assert px.ndim == 4 and px.shape[3] == 4
        f = np.ascontiguousarray(focal, dtype=np.float32)
        pp = np.ascontiguousarray(principal, dtype=np.float32)
        x = np.ascontiguousarray(xforms, dtype=np.float32)
        n, h, w, _ = px.shape
        assert f.shape == (n, 2) and pp.shape == (n, 2) and x.shape == (n, 3, 4)
        check(self._lib.neus2_set_dataset(self._h, C.c_uint32(n), C.c_int32(w), C.c_int32(h), px.ctypes.data_as(C.c_void_p),
                                          f.ctypes.data_as(C.c_void_p), pp.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p)))

    def update_density_grid(self, stream=None):
        check(self._lib.neus2_update_density_grid(self._h, _stream(stream)))

    def train_step_begin(self, rank=0, world_size=1, stream=None):
        check(self._lib.neus2_train_step_begin(self._h, _stream(stream), C.c_uint32(rank), C.c_uint32(world_size)))

    def local_counts(self, stream=None):
        c = (C.c_uint32 * 3)()
        check(self._lib.neus2_local_counts(self._h, _stream(stream), c))
        return [int(x) for x in c]

    def counts_device(self):
        """int32[4] device view of {kept rays, samples, compacted samples, 0} of the step in flight (all-reduce it in place)."""
        p = C.c_void_p()
        check(self._lib.neus2_counts_device(self._h, C.byref(p)))     # the counters ping-pong between two buffers (pipelined step)
        views = self.__dict__.setdefault("_counts_views", {})
        if p.value not in views:
            views[p.value] = _torch().as_tensor(_DevArray(p.value, (4,), "<i4", None), device=f"cuda:{self.device}")
        return views[p.value]

    # ---- data-parallel peers (fused gradient reduction + Adam over NVLink peer memory) ------------------------------------
    IPC_BUFFERS = 4     # NEUS2_IPC_BUFFERS: gradients, fp16 parameters, EMA, sync block

    def ipc_export(self):
        h = (C.c_ubyte * (self.IPC_BUFFERS * 64))()
        check(self._lib.neus2_ipc_export(self._h, h))
        return bytes(h)

    def ipc_attach(self, rank, world_size, all_handles):
        """all_handles: list of world_size `bytes` objects as returned by ipc_export on each rank, in rank order."""
        blob = b"".join(all_handles)
        assert len(blob) == world_size * self.IPC_BUFFERS * 64
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        check(self._lib.neus2_ipc_attach(self._h, C.c_uint32(rank), C.c_uint32(world_size), buf))
        self.peers_attached = True

    def optimizer_shard(self, rank, world_size):
        """(block_params, n_owned): the sharded optimizer deals blocks of block_params parameters round-robin -- parameter i belongs to
        rank (i // block_params) % world_size -- which gives `rank` n_owned parameters."""
        b, e = C.c_uint64(), C.c_uint64()
        check(self._lib.neus2_optimizer_shard(self._h, C.c_uint32(rank), C.c_uint32(world_size), C.byref(b), C.byref(e)))
        return int(b.value), int(e.value)

    def optimizer_step_sharded(self, loss_scale=128.0, stream=None):
        check(self._lib.neus2_optimizer_step_sharded(self._h, _stream(stream), C.c_float(loss_scale)))

    def peer_exchange(self, stream=None):
        """All ranks' backward passes are done; counts_device() / the loss sums / delta_gradients() now hold the sums over ranks."""
        check(self._lib.neus2_peer_exchange(self._h, _stream(stream)))

    def peer_barrier(self, stream=None):
        """Stream-ordered barrier over the attached peers (their earlier work, including peer stores, has completed)."""
        check(self._lib.neus2_peer_barrier(self._h, _stream(stream)))

    def loss_scalars_device(self):
        """(float32[4] device view of the step's {rgb, eikonal, mask, -} loss sums over this rank's rays, due) -- due: this step reports losses."""
        p, due = C.c_void_p(), C.c_int()
        check(self._lib.neus2_loss_scalars_device(self._h, C.byref(p), C.byref(due)))
        views = self.__dict__.setdefault("_loss_views", {})
        if p.value not in views:
            views[p.value] = _torch().as_tensor(_DevArray(p.value, (4,), "<f4", None), device=f"cuda:{self.device}")
        return views[p.value], bool(due.value)

    def train_step_counts_ready(self, stream=None):
        """Data-parallel: the all-reduce of counts_device() is enqueued; take the step's host read-back here (pipelined step)."""
        check(self._lib.neus2_train_step_counts_ready(self._h, _stream(stream)))

    def train_step_end(self, counts_global=None, stream=None):
        st = StepStats()
        cg = None
        if counts_global is not None:
            cg = (C.c_uint32 * 3)(*[int(x) for x in counts_global])
        check(self._lib.neus2_train_step_end(self._h, _stream(stream), cg, C.byref(st)))
        return st

    def train_step(self, stream=None):
        st = StepStats()
        check(self._lib.neus2_train_step(self._h, _stream(stream), C.byref(st)))
        return st

    def get_state(self):
        ts, rpb = C.c_uint32(), C.c_uint32()
        rng = (C.c_uint64 * 2)()
        check(self._lib.neus2_get_state(self._h, C.byref(ts), C.byref(rpb), rng))
        c = (C.c_uint32 * 6)()
        check(self._lib.neus2_get_counters(self._h, c))
        return dict(training_step=ts.value, rays_per_batch=rpb.value, rng_state=(int(rng[0]), int(rng[1])), n_rays_total=int(c[0]),
                    measured_before=int(c[1]), density_grid_ema_step=int(c[2]), optimizer_step=int(c[3]), measured_batch_size=int(c[4]),
                    canonical_training_step=int(c[5]))

    def set_state(self, training_step, rays_per_batch, n_rays_total=0, measured_before=0):
        check(self._lib.neus2_set_state(self._h, C.c_uint32(training_step), C.c_uint32(rays_per_batch), C.c_uint32(n_rays_total), C.c_uint32(measured_before)))

    def set_bitfield(self, bitfield):
        b = np.ascontiguousarray(bitfield, dtype=np.uint8)
        check(self._lib.neus2_set_bitfield(self._h, b.ctypes.data_as(C.c_void_p), C.c_uint64(b.size)))

    def mesh_attributes(self, n_verts, use_inference_params=True, stream=None):
        """(vertex normals [n,3], vertex colours [n,3]) of the mesh of the last marching_cubes call."""
        torch = _torch()
        dev = f"cuda:{self.device}"
        nrm = torch.zeros((n_verts, 3), dtype=torch.float32, device=dev)
        col = torch.zeros((n_verts, 3), dtype=torch.float32, device=dev)
        check(self._lib.neus2_mesh_attributes(self._h, _stream(stream), _ptr(nrm), _ptr(col), C.c_int(int(use_inference_params))))
        return nrm, col

    def next_frame(self, after_learning_rate=0.0, frozen_steps=0, stream=None):
        """Testbed::training_network_next_frame without global-movement prediction (call after set_dataset with the new frame)."""
        check(self._lib.neus2_next_frame(self._h, _stream(stream), C.c_float(after_learning_rate), C.c_uint32(frozen_steps)))

    # ---- global rigid motion of dynamic scenes (hyperparams.predict_global_movement) ---------------------------------
    def begin_global_movement(self, stream=None):
        """accumulated <- delta o accumulated, delta <- identity with a fresh optimizer, m_use_delta = m_train_delta = True
        (the global-movement part of Testbed::training_network_next_frame)."""
        check(self._lib.neus2_begin_global_movement(self._h, _stream(stream)))
        self.global_movement_active = True

    def set_delta_mode(self, use_delta):
        check(self._lib.neus2_set_delta_mode(self._h, C.c_int(int(bool(use_delta)))))
        self.global_movement_active = bool(use_delta)

    def global_movement(self, stream=None):
        """dict(rotation [3,3], transition [3] (accumulated, as stored in half), delta_rotation_6d [6], delta_transition [3], use_delta)."""
        R, t, r6, dt, use = (C.c_float * 9)(), (C.c_float * 3)(), (C.c_float * 6)(), (C.c_float * 3)(), C.c_int()
        check(self._lib.neus2_get_global_movement(self._h, _stream(stream), R, t, r6, dt, C.byref(use)))
        return dict(rotation=np.array(R, dtype=np.float32).reshape(3, 3), transition=np.array(t, dtype=np.float32),
                    delta_rotation_6d=np.array(r6, dtype=np.float32), delta_transition=np.array(dt, dtype=np.float32), use_delta=bool(use.value))

    def set_global_movement(self, rotation=None, transition=None, delta_rotation_6d=None, delta_transition=None, first_frame_offset=None, stream=None):
        def arr(v, n):
            if v is None:
                return None
            a = np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(-1))
            assert a.size == n
            return (C.c_float * n)(*a.tolist())
        check(self._lib.neus2_set_global_movement(self._h, _stream(stream), arr(rotation, 9), arr(transition, 3), arr(delta_rotation_6d, 6),
                                                  arr(delta_transition, 3), arr(first_frame_offset, 3)))

    def delta_gradients(self):
        """Zero-copy device view of the 12-float delta-gradient accumulator ([transition 3 | rotation-6d 6 | 0 0 0])."""
        torch = _torch()
        a = C.c_void_p()
        check(self._lib.neus2_delta_gradients_device(self._h, C.byref(a)))
        return torch.as_tensor(_DevArray(a.value, (12,), "<f4", None), device=f"cuda:{self.device}")

    def reinit_sdf_mlp(self, sdf_mlp_weights=None, stream=None):
        """hyperparams.incremental_reinit_sdf_mlp: the fp32 master copy of the SDF MLP goes back to its geometric initialisation."""
        w = None
        if sdf_mlp_weights is not None:
            w = np.ascontiguousarray(sdf_mlp_weights, dtype=np.float32)
            assert w.size == 64 * 32 + 16 * 64
        check(self._lib.neus2_reinit_sdf_mlp(self._h, _stream(stream), w.ctypes.data_as(C.c_void_p) if w is not None else None))

    def time_frame(self):
        tf, lr = C.c_uint32(), C.c_float()
        check(self._lib.neus2_get_time_frame(self._h, C.byref(tf), C.byref(lr)))
        return int(tf.value), float(lr.value)

    def optimizer_buffers(self):
        """(first moments fp32, second moments fp32, per-parameter steps int32, current_step) -- zero-copy device views."""
        torch = _torch()
        a, b, c, st = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint32()
        check(self._lib.neus2_optimizer_buffers(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(st)))
        dev, n = f"cuda:{self.device}", self.n_params
        return (torch.as_tensor(_DevArray(a.value, (n,), "<f4", None), device=dev), torch.as_tensor(_DevArray(b.value, (n,), "<f4", None), device=dev),
                torch.as_tensor(_DevArray(c.value, (n,), "<i4", None), device=dev), int(st.value))

    def set_optimizer_step(self, step):
        check(self._lib.neus2_set_optimizer_step(self._h, C.c_uint32(step)))

    def set_density_grid(self, grid, ema_step=0, stream=None):
        g = np.ascontiguousarray(grid, dtype=np.float32)
        check(self._lib.neus2_set_density_grid(self._h, _stream(stream), g.ctypes.data_as(C.c_void_p), C.c_uint64(g.size), C.c_uint32(ema_step)))

    STAGES = ["density_grid", "gen_samples", "forward", "loss", "fwdbwd", "wgrad", "optimizer", "step"]

    def set_profiling(self, enable=True):
        check(self._lib.neus2_set_profiling(self._h, C.c_int(int(enable))))

    STAGES_EXT = STAGES + ["peer_exchange_counts", "peer_exchange_grads", "peer_barrier"]

    def stage_times(self):
        """{stage: (mean ms, samples)} measured with CUDA events on the step's stream (data-parallel sync points included)."""
        ms = (C.c_float * 12)()
        cnt = (C.c_uint32 * 12)()
        check(self._lib.neus2_get_stage_times_ext(self._h, ms, cnt))
        return {s: ((ms[i] / cnt[i]) if cnt[i] else 0.0, int(cnt[i])) for i, s in enumerate(self.STAGES_EXT)}

    def launch_count(self):
        return int(self._lib.neus2_launch_count(self._h))

    # ---- inference rendering (Testbed::render_to_cpu, python_api.cu:123-170) -------------------------------------------
    def render_params(self, width, height, camera_matrix=None, focal_length=None, screen_center=None, **kw):
        """neus2_render_params with the reference's defaults (reset_camera, testbed.cu:272-287); keyword arguments set fields."""
        p = RenderParams()
        self._lib.neus2_default_render_params(C.byref(p), C.c_int32(width), C.c_int32(height))
        if camera_matrix is not None:
            p.camera_matrix[:] = [float(v) for v in np.asarray(camera_matrix, dtype=np.float32).reshape(-1)[:12]]
        if focal_length is not None:
            f = np.broadcast_to(np.asarray(focal_length, dtype=np.float32), (2,))
            p.focal_length[:] = [float(f[0]), float(f[1])]
        if screen_center is not None:
            p.screen_center[:] = [float(screen_center[0]), float(screen_center[1])]
        for k, v in kw.items():
            if k == "background_color":
                p.background_color[:] = [float(x) for x in v]
            else:
                setattr(p, k, v)
        return p

    def training_view_params(self, xform, focal, principal, train_res, width=None, height=None, zoom=1.0, **kw):
        """Testbed::set_camera_to_training_view (testbed.cu:264-270) followed by calc_focal_length / render_screen_center
        (testbed.cu:2738-2746): the view of training image (xform, focal [px], principal [relative], train_res) rendered at
        width x height (default: the training resolution)."""
        fov_axis = 1                                                             # m_fov_axis, testbed.h:500
        width = int(train_res[0]) if width is None else int(width)
        height = int(train_res[1]) if height is None else int(height)
        rel_focal = np.asarray(focal, dtype=np.float32) / np.float32(train_res[fov_axis])
        f = rel_focal * np.float32((width, height)[fov_axis]) * np.float32(zoom)
        m_sc = np.float32(1.0) - np.asarray(principal, dtype=np.float32)          # m_screen_center
        sc = (np.float32((np.float32(0.5) - m_sc[0]) * np.float32(zoom) + np.float32(0.5)),   # x in float, y in double (:2745)
              np.float32((0.5 - float(m_sc[1])) * float(np.float32(zoom)) + 0.5))
        return self.render_params(width, height, camera_matrix=xform, focal_length=f, screen_center=sc, **kw)

    def render(self, params, spp=1, out=None, depth=None, stream=None):
        """Renders params' tile for samples 0..spp-1 and returns the [tile_h, tile_w, 4] float32 device tensor."""
        torch = _torch()
        dev = f"cuda:{self.device}"
        if out is None:
            out = torch.empty((params.tile_h, params.tile_w, 4), dtype=torch.float32, device=dev)
        for s in range(spp):
            params.sample_index = s
            check(self._lib.neus2_render(self._h, _stream(stream), C.byref(params), _ptr(out), _ptr(depth)))
        return out

    def render_stats(self, stream=None):
        st = (C.c_uint32 * 2)()
        check(self._lib.neus2_render_stats(self._h, _stream(stream), st))
        return dict(rays_alive=int(st[0]), samples=int(st[1]))

    # ---- SDF lattice + marching cubes (Testbed::marching_cubes, testbed_nerf.cu:4175-4226) ------------------------------
    def sdf_lattice(self, res, aabb=(0, 0, 0, 1, 1, 1), z0=0, z1=None, use_inference_params=True, out=None, stream=None):
        torch = _torch()
        r = (C.c_uint32 * 3)(*[int(v) for v in res])
        bb = (C.c_float * 6)(*[float(v) for v in aabb])
        z1 = int(res[2]) if z1 is None else int(z1)
        if out is None:
            out = torch.empty((z1 - z0, int(res[1]), int(res[0])), dtype=torch.float32, device=f"cuda:{self.device}")
        check(self._lib.neus2_sdf_lattice(self._h, _stream(stream), r, bb, C.c_uint32(z0), C.c_uint32(z1), _ptr(out), C.c_int(int(use_inference_params))))
        return out

    def marching_cubes(self, density, aabb=(0, 0, 0, 1, 1, 1), thresh=0.0, stream=None):
        """density: float32 device tensor [rz, ry, rx].  Returns (verts [nv, 3] float32, faces [nt, 3] int64) device tensors."""
        torch = _torch()
        assert density.dtype == torch.float32 and density.is_contiguous() and density.dim() == 3
        r = (C.c_uint32 * 3)(density.shape[2], density.shape[1], density.shape[0])
        bb = (C.c_float * 6)(*[float(v) for v in aabb])
        nv, ni = C.c_uint32(), C.c_uint32()
        check(self._lib.neus2_marching_cubes(self._h, _stream(stream), r, bb, C.c_float(thresh), _ptr(density), C.byref(nv), C.byref(ni)))
        pv, pi = C.c_void_p(), C.c_void_p()
        check(self._lib.neus2_mesh_buffers(self._h, C.byref(pv), C.byref(pi), None, None))
        dev = f"cuda:{self.device}"
        if nv.value == 0 or ni.value == 0:
            return torch.zeros((0, 3), dtype=torch.float32, device=dev), torch.zeros((0, 3), dtype=torch.int64, device=dev)
        verts = torch.as_tensor(_DevArray(pv.value, (nv.value, 3), "<f4", self), device=dev).clone()
        faces = torch.as_tensor(_DevArray(pi.value, (ni.value // 3, 3), "<i4", self), device=dev).to(torch.int64)
        return verts, faces
