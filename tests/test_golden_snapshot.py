"""CPU pin of the snapshot file format and of the optimizer arithmetic against the REFERENCE's own code run on the CPU: its
Ema{ExponentialDecay{Adam}} classes (step kernels, serialize / deserialize), its GPUMemory <-> json binary conversion and its json / msgpack library,
compiled as host code from where they lie (oracle/Makefile.ref -> libref_snapshot_host.so; tests/golden/make_snapshot_golden.py).

  * a file the reference's code wrote (tests/golden/snapshot_reference.msgpack) is read and restored by neus2_b200.snapshot;
  * a file neus2_b200.snapshot wrote is loaded by the reference's code (from_msgpack, from_json(GPUMemory), the three deserialize()) -- where the tree is mounted;
  * the oracle's Adam / EMA restatement follows the reference's optimizer objects over 45 steps with a decaying learning rate.
"""
import importlib.util
import os
import types

import numpy as np
import pytest

from neus2_b200 import snapshot
from oracle import oracle as O
from oracle import ref_snapshot_host as RS

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD_FILE = os.path.join(HERE, "golden", "snapshot_reference.msgpack")
GOLD = os.path.join(HERE, "golden", "snapshot_reference.npz")
needs_reference = pytest.mark.skipif(not RS.available(), reason="needs the reference tree (oracle/_ref/libref_snapshot_host.so)")


def _mod():
    spec = importlib.util.spec_from_file_location("make_snapshot_golden", os.path.join(HERE, "golden", "make_snapshot_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class FakeContext:
    """The slice of neus2_b200.Context that snapshot.save / restore touch, over CPU tensors (the file format is host logic)."""

    def __init__(self, n_params, learning_rate=1e-2, batch_size=1 << 18):
        import torch
        self.n_params = n_params
        self.cfg = types.SimpleNamespace(learning_rate=learning_rate, batch_size=batch_size, aabb_scale=1)
        self.master = torch.zeros(n_params); self.params = torch.zeros(n_params, dtype=torch.float16); self.ema = torch.zeros(n_params, dtype=torch.float16)
        self.grads = torch.zeros(n_params)
        self.m1 = torch.zeros(n_params); self.m2 = torch.zeros(n_params); self.steps = torch.zeros(n_params, dtype=torch.int32)
        self.current_step = 0
        self.state = dict(training_step=0, rays_per_batch=4096, n_rays_total=0, measured_before=0, density_grid_ema_step=0)
        self.grid = torch.zeros(128 ** 3)
        self.gm = dict(rotation=np.eye(3, dtype=np.float32), transition=np.zeros(3, np.float32), delta_rotation_6d=np.array([1, 0, 0, 0, 1, 0], np.float32),
                       delta_transition=np.zeros(3, np.float32), use_delta=False)

    def buffers(self):
        return self.master, self.params, self.ema, self.grads

    def optimizer_buffers(self):
        return self.m1, self.m2, self.steps, self.current_step

    def get_state(self):
        return dict(self.state)

    def debug_buffer(self, name, typestr, shape):
        assert name == "density_grid"
        return self.grid

    def global_movement(self):
        return dict(self.gm)

    def set_global_movement(self, rotation=None, transition=None, delta_rotation_6d=None, delta_transition=None):
        self.gm.update(rotation=np.asarray(rotation, np.float32).reshape(3, 3), transition=np.asarray(transition, np.float32),
                       delta_rotation_6d=np.asarray(delta_rotation_6d, np.float32), delta_transition=np.asarray(delta_transition, np.float32))

    def set_delta_mode(self, use_delta):
        self.gm["use_delta"] = bool(use_delta)

    def set_params(self, master, reset_optimizer=True):
        import torch
        self.master.copy_(torch.from_numpy(np.asarray(master, np.float32)))
        self.params.copy_(self.master.half())
        if reset_optimizer:
            self.m1.zero_(); self.m2.zero_(); self.steps.zero_(); self.current_step = 0

    def set_optimizer_step(self, step):
        self.current_step = int(step)

    def set_density_grid(self, grid, ema_step=0):
        import torch
        self.grid.copy_(torch.from_numpy(np.asarray(grid, np.float32)))
        self.state["density_grid_ema_step"] = ema_step

    def set_state(self, training_step, rays_per_batch, n_rays_total=0, measured_before=0):
        self.state.update(training_step=training_step, rays_per_batch=rays_per_batch, n_rays_total=n_rays_total, measured_before=measured_before)


def test_file_written_by_the_reference_code_restores():
    g, mod = np.load(GOLD), _mod()
    cfg = snapshot.read(GOLD_FILE)
    assert cfg["optimizer"] == mod.OPTIMIZER and cfg["encoding"]["n_levels"] == 14            # the network config rides along at the top level
    snap = cfg["snapshot"]
    n = mod.N_DENSITY + mod.N_RGB + mod.N_GRID + mod.N_VARIANCE
    assert snap["n_params"] == n and snap["density_grid_size"] == 128 and len(snap["density_grid_binary"]) == 0
    assert np.array_equal(np.frombuffer(snap["params_binary"], np.uint16), g["params_inference"])
    assert np.array_equal(g["params_inference"], g["ema"]) and not np.array_equal(g["params"], g["ema"])    # inference weights = the EMA (trainer.h:90-93)
    adam = snap["optimizer"]["nested"]["nested"]
    assert adam["current_step"] == mod.STEPS and abs(snap["optimizer"]["nested"]["learning_rate_factor"] - 0.33 ** 3) < 1e-7
    ctx = FakeContext(n)
    loss = snapshot.restore(ctx, snap)
    assert abs(loss - mod.TESTBED["loss"]) < 1e-7
    assert np.array_equal(ctx.params.numpy().view(np.uint16), g["params_inference"])           # Trainer::deserialize: the blob is both parameter sets ...
    assert np.array_equal(ctx.master.numpy(), g["params_inference"].view(np.float16).astype(np.float32))     # ... and the master copy is its widening
    assert np.array_equal(ctx.ema.numpy().view(np.uint16), g["ema"])
    assert np.array_equal(ctx.m1.numpy(), g["m1"]) and np.array_equal(ctx.m2.numpy(), g["m2"]) and np.array_equal(ctx.steps.numpy().astype(np.uint32), g["steps"])
    assert ctx.current_step == mod.STEPS
    t = mod.TESTBED
    assert ctx.state["training_step"] == t["training_step"] and ctx.state["rays_per_batch"] == t["rays_per_batch"] and ctx.state["measured_before"] == t["measured_before"]
    assert np.array_equal(ctx.gm["rotation"], np.eye(3)) and np.array_equal(ctx.gm["transition"], t["transition"])
    assert np.array_equal(ctx.gm["delta_rotation_6d"], t["local_rotation"]) and np.array_equal(ctx.gm["delta_transition"], t["local_transition"])
    with pytest.raises(RuntimeError, match="wrong size"):
        snapshot.restore(FakeContext(n + 1), snap)


def test_oracle_optimizer_follows_the_reference_objects():
    g, mod = np.load(GOLD), _mod()
    master, grads = mod.inputs()
    n, n_matrix = master.size, mod.N_DENSITY + mod.N_RGB
    ph = master.astype(np.float16).view(np.uint16).copy()
    m1 = np.zeros(n, np.float32); m2 = np.zeros(n, np.float32); steps = np.zeros(n, np.uint32); ema = np.zeros(n, np.uint16)
    a = mod.OPTIMIZER["nested"]["nested"]; d = mod.OPTIMIZER["nested"]
    factor = np.float32(1.0)
    for k in range(mod.STEPS):
        if k >= d["decay_start"] and (k - d["decay_start"]) % d["decay_interval"] == 0:         # exponential_decay.h:60-71, evaluated before the nested step
            factor = np.float32(factor * np.float32(d["decay_base"]))
        O.adam_ema_step(n_matrix, mod.LOSS_SCALE, np.float32(a["learning_rate"]) * factor, a["beta1"], a["beta2"], a["epsilon"], a["l2_reg"], mod.OPTIMIZER["decay"], k + 1,
                        grads[k].astype(np.float32), master, ph, m1, m2, steps, ema)
    assert abs(float(factor) - float(g["learning_rate_factor"])) < 1e-8
    assert np.array_equal(steps, g["steps"])                                     # per-parameter counters: grid / variance entries skip zero gradients
    assert steps[:n_matrix].min() == mod.STEPS and steps[n_matrix:-4].max() < mod.STEPS and steps[-4] == mod.STEPS and steps[-1] == 0
    # the oracle writes the fused multiply-adds nvcc puts into the reference's kernels (bit-exact against them on the GPU, tests/test_ref_kernels.py);
    # the host build of the same source rounds every product: agreement to float rounding
    assert np.allclose(m1, g["m1"], rtol=2e-6, atol=1e-9) and np.allclose(m2, g["m2"], rtol=2e-6, atol=1e-12)
    assert np.allclose(master, g["master"], rtol=0, atol=2e-6)
    from tests import util as U
    assert U.half_ulp_diff(ph, g["params"]).max() <= 1 and U.half_ulp_diff(ema, g["ema"]).max() <= 1


@needs_reference
def test_file_written_by_this_repo_loads_through_the_reference_code(tmp_path):
    import torch
    mod = _mod()
    n = mod.N_DENSITY + mod.N_RGB + mod.N_GRID + mod.N_VARIANCE
    rs = np.random.RandomState(3)
    ctx = FakeContext(n, learning_rate=1e-2, batch_size=262144)
    ctx.master.copy_(torch.from_numpy(rs.normal(0, 0.1, n).astype(np.float32))); ctx.params.copy_(ctx.master.half())
    ctx.ema.copy_(torch.from_numpy(rs.normal(0, 0.1, n).astype(np.float16)))
    ctx.m1.copy_(torch.from_numpy(rs.normal(0, 1, n).astype(np.float32))); ctx.m2.copy_(torch.from_numpy(rs.uniform(0, 1, n).astype(np.float32)))
    ctx.steps.copy_(torch.from_numpy(rs.randint(0, 30, n).astype(np.int32))); ctx.current_step = 30
    ctx.state.update(training_step=30, rays_per_batch=5120, measured_before=600000, measured_batch_size=250000)   # the MEASURED compacted count, not the target
    ctx.grid.copy_(torch.from_numpy(rs.uniform(0, 2, 128 ** 3).astype(np.float32)))
    ctx.gm.update(rotation=np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], np.float32), transition=np.array([0.5, 0.25, -0.125], np.float32),
                  delta_rotation_6d=np.array([1, 0.125, 0, 0, 1, 0.25], np.float32), delta_transition=np.array([0, 0.0625, 0], np.float32))
    path = str(tmp_path / "ours.msgpack")
    snapshot.save(ctx, path, mod.NETWORK_CONFIG, include_optimizer_state=True, loss=0.5, aabb_scale=2)
    st = RS.Stack(mod.OPTIMIZER, mod.N_DENSITY, mod.N_RGB, mod.N_GRID, mod.N_VARIANCE, mod.LAYERS)
    try:
        st.load_snapshot(path)                                                  # json::from_msgpack + the reference's deserialize() chain
        s, t = st.state(), st.testbed_state(with_density_grid=True)
    finally:
        st.close()
    ema_bits = ctx.ema.numpy().view(np.uint16)
    assert np.array_equal(s["ema"], ema_bits) and np.array_equal(s["params_inference"], ema_bits) and np.array_equal(s["params"], ema_bits)   # params_binary = inference weights
    assert np.array_equal(s["master"], ctx.ema.numpy().astype(np.float32))
    assert np.array_equal(s["m1"], ctx.m1.numpy()) and np.array_equal(s["m2"], ctx.m2.numpy()) and np.array_equal(s["steps"], ctx.steps.numpy().astype(np.uint32))
    assert s["current_step"] == 30 and abs(s["learning_rate"] - 1e-2) < 1e-9 and s["learning_rate_factor"] == 1.0
    assert (t["training_step"], t["rays_per_batch"], t["measured_batch_size"], t["measured_batch_size_before_compaction"], t["aabb_scale"]) == (30, 5120, 250000, 600000, 2)
    assert t["loss"] == 0.5 and t["density_grid_size"] == 128 ** 3
    assert np.array_equal(t["density_grid"], ctx.grid.numpy().astype(np.float16))
    assert np.array_equal(t["rotation"][:9].astype(np.float32).reshape(3, 3), ctx.gm["rotation"]) and t["rotation"].size == 12
    assert np.array_equal(t["transition"][:3].astype(np.float32), ctx.gm["transition"]) and t["transition"].size == 4
    assert np.array_equal(t["local_rotation"][:6].astype(np.float32), ctx.gm["delta_rotation_6d"]) and t["local_rotation"].size == 8
    assert np.array_equal(t["local_transition"][:3].astype(np.float32), ctx.gm["delta_transition"]) and t["local_transition"].size == 4
    # a file without a snapshot is refused with the reference's message
    bad = str(tmp_path / "bad.msgpack")
    import msgpack
    open(bad, "wb").write(msgpack.packb({"encoding": {}}, use_bin_type=True))
    st = RS.Stack(mod.OPTIMIZER, mod.N_DENSITY, mod.N_RGB, mod.N_GRID, mod.N_VARIANCE, mod.LAYERS)
    with pytest.raises(RuntimeError, match="does not contain a snapshot"):
        st.load_snapshot(bad)
    st.close()


@needs_reference
def test_committed_fixture_is_what_the_reference_source_produces(tmp_path):
    mod, g = _mod(), np.load(GOLD)
    mod.main(str(tmp_path))
    assert open(os.path.join(str(tmp_path), "snapshot_reference.msgpack"), "rb").read() == open(GOLD_FILE, "rb").read()
    fresh = np.load(os.path.join(str(tmp_path), "snapshot_reference.npz"))
    assert sorted(fresh.files) == sorted(g.files)
    for name in g.files:
        assert np.array_equal(fresh[name], g[name]), name


def test_learning_rate_factor_of_the_library_is_the_reference_schedule():
    """neus2_learning_rate_factor (C ABI, host arithmetic) == m_learning_rate_factor of the reference's ExponentialDecayOptimizer."""
    import ctypes as C
    import neus2_b200 as N
    from neus2_b200 import _lib
    g, mod = np.load(GOLD), _mod()
    d = mod.OPTIMIZER["nested"]
    cfg = N.default_config(decay_start=d["decay_start"], decay_interval=d["decay_interval"], decay_base=d["decay_base"])
    f = lambda c, s: float(_lib.load().neus2_learning_rate_factor(C.byref(c), C.c_uint32(s)))
    assert np.float32(f(cfg, mod.STEPS - 1)) == np.float32(g["learning_rate_factor"])           # after 45 steps: three decays, float products
    assert f(cfg, 19) == 1.0 and f(cfg, 20) == f(cfg, 29) == float(np.float32(0.33)) and f(cfg, 30) == float(np.float32(0.33) * np.float32(0.33))
    base = N.default_config()                                                                  # base.json: 20000 / 10000 / 0.33, no decay_end
    assert (base.decay_start, base.decay_interval, base.decay_end) == (20000, 10000, 10000000) and abs(base.decay_base - 0.33) < 1e-7
    assert f(base, 19999) == 1.0 and f(base, 20000) == float(np.float32(0.33)) and f(base, 29999) == float(np.float32(0.33))
    ended = N.default_config(decay_start=5, decay_interval=5, decay_end=12, decay_base=0.5)
    assert [f(ended, s) for s in (4, 5, 9, 10, 14, 15, 100)] == [1.0, 0.5, 0.5, 0.25, 0.25, 0.25, 0.25]    # no trigger past decay_end
    assert f(N.default_config(decay_interval=0), 10 ** 6) == 1.0
    if RS.available():      # step by step against the live reference objects
        master, grads = mod.inputs()
        st = RS.Stack(mod.OPTIMIZER, mod.N_DENSITY, mod.N_RGB, mod.N_GRID, mod.N_VARIANCE, mod.LAYERS)
        try:
            st.set_params_full_precision(master)
            for k in range(mod.STEPS):
                st.optimizer_step(mod.LOSS_SCALE, grads[k])
                assert np.float32(st.state()["learning_rate_factor"]) == np.float32(f(cfg, k)), k
        finally:
            st.close()
