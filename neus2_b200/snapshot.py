"""Snapshot files: the msgpack layout of Testbed::save_snapshot / load_snapshot (src/testbed.cu:3144-3266)
with the Trainer / Ema / ExponentialDecay / Adam serialisers of the reference's tiny-cuda-nn fork (trainer.h:281-304,
optimizers/ema.h:182-194, exponential_decay.h:128-140, adam.h:424-445): the network config json at the top level and

  snapshot: { n_params, params_binary (fp16 INFERENCE = EMA weights), optimizer?: { nested: { nested: { current_step,
              base_learning_rate, first_moments_binary, second_moments_binary, param_steps_binary }, learning_rate,
              learning_rate_factor }, weights_ema_binary }, density_grid_size: 128, density_grid_binary (fp16),
              nerf: { aabb_scale, rgb: { rays_per_batch, measured_batch_size, measured_batch_size_before_compaction } },
              training_step, loss, rotation (12 halves), transition (4), local_rotation (8), local_transition (4): the accumulated
              global rigid motion and the current DeltaNetwork (nerf_network.h:1180-1278) }

Loading follows Trainer::deserialize: the fp16 blob becomes BOTH the training and the inference parameters and the fp32
master copy is its widening (trainer.h:263-275).  Host logic only.
"""
import ctypes as C

import numpy as np


def read(path):
    import msgpack
    with open(path, "rb") as f:
        return msgpack.unpackb(f.read(), raw=False, strict_map_key=False)


def _n_cascades(aabb_scale):
    """m_nerf.max_cascade + 1 (testbed_nerf.cu:3069-3073): the smallest c with 2^c >= aabb_scale, plus one."""
    c = 0
    while (1 << c) < int(aabb_scale):
        c += 1
    return c + 1


def save(ctx, path, network_config, include_optimizer_state=False, loss=0.0, aabb_scale=None):
    """aabb_scale: defaults to the context's.  Not written: `nerf.dataset` (the reference stores the camera metadata so that a snapshot
    can be rendered without its dataset, testbed.cu:3170; here a snapshot is always loaded next to the dataset it was trained on)."""
    import msgpack
    import torch
    master, params, ema, _ = ctx.buffers()
    m1, m2, steps, current_step = ctx.optimizer_buffers()
    state = ctx.get_state()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    inference = ema if current_step > 0 else params          # before the first optimizer step the inference weights are the initial ones
    snap = {"n_params": int(ctx.n_params), "params_binary": inference.cpu().numpy().tobytes()}
    if include_optimizer_state:
        require_whole = getattr(ctx, "_require_whole_optimizer_state", None)    # sharded optimizer: parallel.sync_optimizer_state first
        if require_whole is not None:
            require_whole("snapshot.save(include_optimizer_state=True)")
        adam = {"current_step": int(current_step), "base_learning_rate": float(ctx.cfg.learning_rate),
                "first_moments_binary": m1.cpu().numpy().tobytes(), "second_moments_binary": m2.cpu().numpy().tobytes(),
                "param_steps_binary": steps.cpu().numpy().tobytes()}
        from ._lib import load
        # m_learning_rate_factor after current_step steps; a duck-typed config without the decay schedule (tests) has not decayed
        factor = float(load().neus2_learning_rate_factor(C.byref(ctx.cfg), C.c_uint32(max(int(current_step), 1) - 1))) if isinstance(ctx.cfg, C.Structure) else 1.0
        snap["optimizer"] = {"nested": {"nested": adam, "learning_rate": float(ctx.cfg.learning_rate), "learning_rate_factor": factor},
                             "weights_ema_binary": ema.cpu().numpy().tobytes()}
    n_cascades = _n_cascades(ctx.cfg.aabb_scale)                              # the grid the context holds
    grid = ctx.debug_buffer("density_grid", "<f4", (128 ** 3 * n_cascades,)).cpu().numpy()   # all cascades (testbed.cu:3163: m_nerf.density_grid)
    snap["density_grid_size"] = 128
    snap["density_grid_ema_step"] = int(state["density_grid_ema_step"])    # extension: the reference restarts its decay counter
    snap["density_grid_binary"] = grid.astype(np.float16).tobytes()
    snap["nerf"] = {"aabb_scale": int(aabb_scale if aabb_scale else ctx.cfg.aabb_scale),
                    "rgb": {"rays_per_batch": int(state["rays_per_batch"]), "measured_batch_size": int(state.get("measured_batch_size", 0)),
                            "measured_batch_size_before_compaction": int(state.get("measured_before", 0))}}
    snap["training_step"] = int(state["training_step"])
    snap["n_rays_total"] = int(state["n_rays_total"])                       # extension (image round-robin position)
    snap["loss"] = float(loss)
    # save_global_movement / save_local_movement (nerf_network.h:1180-1247): the accumulated motion (12 halves: 3x3 row-major + pad;
    # 4 halves) and the current DeltaNetwork (8 halves: 6d rotation + pad; 4 halves)
    gm = ctx.global_movement()
    snap["rotation"] = np.concatenate([gm["rotation"].reshape(9), np.zeros(3, np.float32)]).astype(np.float16).tobytes()
    snap["transition"] = np.concatenate([gm["transition"], np.zeros(1, np.float32)]).astype(np.float16).tobytes()
    snap["local_rotation"] = np.concatenate([gm["delta_rotation_6d"], np.zeros(2, np.float32)]).astype(np.float16).tobytes()
    snap["local_transition"] = np.concatenate([gm["delta_transition"], np.zeros(1, np.float32)]).astype(np.float16).tobytes()
    snap["use_delta"] = bool(gm["use_delta"])                               # extension: the reference re-derives it from the frame index
    out = dict(network_config)
    out["snapshot"] = snap
    with open(path, "wb") as f:
        f.write(msgpack.packb(out, use_bin_type=True))


def restore(ctx, snap):
    """Returns the stored loss.  Raises RuntimeError on incompatible files like the reference."""
    if int(snap.get("density_grid_size", 128)) != 128:
        raise RuntimeError("Incompatible grid size.")
    p16 = np.frombuffer(snap["params_binary"], dtype=np.float16)
    if p16.size != ctx.n_params:
        raise RuntimeError("Can't set params because CPU buffer has the wrong size.")
    # load_global_movement / load_local_movement (nerf_network.h:1207-1278)
    eye12 = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], np.float16)
    rot = np.frombuffer(snap.get("rotation", eye12.tobytes()), dtype=np.float16).astype(np.float32)
    if rot.size < 9:
        rot = eye12.astype(np.float32)                                   # files of earlier builds stored a 6-value identity placeholder
    tr = np.frombuffer(snap.get("transition", np.zeros(4, np.float16).tobytes()), dtype=np.float16).astype(np.float32)
    lrot = np.frombuffer(snap.get("local_rotation", np.array([1, 0, 0, 0, 1, 0, 0, 0], np.float16).tobytes()), dtype=np.float16).astype(np.float32)
    ltr = np.frombuffer(snap.get("local_transition", np.zeros(4, np.float16).tobytes()), dtype=np.float16).astype(np.float32)
    ctx.set_global_movement(rotation=rot[:9], transition=tr[:3], delta_rotation_6d=lrot[:6], delta_transition=ltr[:3])
    ctx.set_delta_mode(bool(snap.get("use_delta", False)))
    import torch
    ctx.set_params(p16.astype(np.float32), reset_optimizer=True)     # master = widened fp16, params = fp16, optimizer cleared
    master, params, ema, _ = ctx.buffers()
    ema.copy_(params)                                                # m_params_inference == m_params after Trainer::set_params
    opt = snap.get("optimizer")
    if opt is not None:
        adam = opt["nested"]["nested"]
        m1, m2, steps, _ = ctx.optimizer_buffers()
        dev = m1.device
        m1.copy_(torch.from_numpy(np.frombuffer(adam["first_moments_binary"], dtype=np.float32).copy()).to(dev))
        m2.copy_(torch.from_numpy(np.frombuffer(adam["second_moments_binary"], dtype=np.float32).copy()).to(dev))
        if "param_steps_binary" in adam:
            steps.copy_(torch.from_numpy(np.frombuffer(adam["param_steps_binary"], dtype=np.int32).copy()).to(dev))
        ema.copy_(torch.from_numpy(np.frombuffer(opt["weights_ema_binary"], dtype=np.float16).copy()).to(dev))
        ctx.set_optimizer_step(int(adam["current_step"]))
    grid16 = np.frombuffer(snap["density_grid_binary"], dtype=np.float16)
    if grid16.size == 128 ** 3 * _n_cascades(ctx.cfg.aabb_scale):                 # testbed.cu:3226-3230
        ctx.set_density_grid(grid16.astype(np.float32), ema_step=int(snap.get("density_grid_ema_step", 0)))
    elif grid16.size != 0:
        raise RuntimeError("Incompatible number of grid cascades.")
    rgb = snap["nerf"]["rgb"]
    ctx.set_state(int(snap["training_step"]), int(rgb["rays_per_batch"]), int(snap.get("n_rays_total", 0)), int(rgb["measured_batch_size_before_compaction"]))
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return float(snap.get("loss", 0.0))
