"""2-GPU checks of the data-parallel training step (skipped on a single-GPU box): the optimizer fused with its collective
(gradient slices summed through NVLink peer loads inside the Adam kernel, parameters published by peer stores) must leave
both ranks with IDENTICAL parameters and sample counts, and must agree with the NCCL all-reduce + replicated-Adam path to
rounding (the backward's fp32 atomics are not order-deterministic, so no two runs are bit-equal)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, mode, steps, total_world=None):
    import torch
    import torch.distributed as dist
    import neus2_b200 as N
    from neus2_b200 import parallel, synthetic
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    B = 1 << 15
    total_world = total_world or world                   # "single*" modes: one rank stepping the SAME global batch the multi-rank runs split
    # "*capped": rays_per_batch runs into max_rays_per_batch, i.e. every rank traces max_rays_per_batch / W rays -- exactly the size of its
    # per-ray arrays (a reduction or zero-fill sized by the GLOBAL ray count would run off their end)
    # (few enough rays that the first pass's sample budget, 16 x the rank's batch, never binds: which rays a full buffer drops depends
    #  on where the scan over rays starts, i.e. on the sharding; with the budget out of the way one GPU and W GPUs trace the same rays)
    capped = mode.endswith("capped")
    max_rays = 128 * total_world if capped else (1 << 15) * total_world
    if capped:
        B *= 4                                           # the batch wants ~500 rays per rank pair, the cap allows 128 per rank
    ctx = N.Context(N.default_config(batch_size=B * (total_world // 2), world_size=world, max_rays_per_batch=max_rays, mask_loss_weight=0.1,
                                     prefetch_next_samples=0 if mode.endswith("_inorder") else 1), device=rank)
    data = synthetic.render_sphere_views(n_views=6, width=128, height=96, seed=2)
    ctx.set_dataset(data["pixels"], data["focal"], data["principal"], data["xforms"])
    ctx.init_params()
    if mode.startswith("p2p"):
        parallel.attach_peers(ctx, rank, world)
    stats, losses = [], []
    master, params, ema, _ = ctx.buffers()
    block, n_owned = ctx.optimizer_shard(rank, world)
    owned = (torch.arange(ctx.n_params, device=master.device) // block) % world == rank        # the sharded optimizer's ownership map
    saved = {}
    for i in range(steps):
        st = parallel.train_step(ctx, rank, world)
        stats.append((st.rays_per_batch, st.n_rays_kept, st.measured_batch_size_before_compaction, st.measured_batch_size))
        if st.loss_valid:
            losses.append((i, st.loss, st.ek_loss, st.mask_loss))
        if i == 23:      # the parameter comparison between the two optimizer paths is made here, while their trajectories are within rounding
            torch.cuda.synchronize()
            dist.barrier()
            saved = dict(params=params.cpu().numpy().view(np.uint16), ema=ema.cpu().numpy().view(np.uint16),
                         master_shard=master[owned].cpu().numpy(), shard=np.array([block, n_owned]))
            dist.barrier()
    torch.cuda.synchronize()
    dist.barrier()
    saved["stats"] = np.array(stats, dtype=np.int64)
    saved["losses"] = np.array(losses, dtype=np.float64)
    saved["params_end"] = params.cpu().numpy().view(np.uint16)
    if mode.startswith("single") or mode.endswith("capped"):
        np.savez(os.path.join(out_dir, f"{mode}_r{rank}.npz"), **saved)
        dist.barrier()
        ctx.close()
        dist.destroy_process_group()
        return
    if mode == "p2p":
        # sharded optimizer: this rank's fp32 master / moments / step counters are current for ITS slice only.  Whole-state readers must
        # refuse until the slices have been gathered (parallel.sync_optimizer_state), after which every rank holds the same, complete state.
        try:
            ctx.get_params()
            saved["refused"] = np.array(0)
        except N.Neus2Error:
            saved["refused"] = np.array(1)
        m1, m2, psteps, _ = ctx.optimizer_buffers()
        stale = (master.half() != params).float().mean().item()
        parallel.sync_optimizer_state(ctx, rank, world)
        whole = ctx.get_params()
        saved["stale_before_sync"] = np.array(stale)
        saved["master_rounds_to_params"] = np.array(bool(np.array_equal(whole.astype(np.float16).view(np.uint16), saved["params_end"])))
        saved["opt_digest"] = np.array([float(m1.double().sum().item()), float(m2.double().sum().item()), float(psteps.double().sum().item())])
        from neus2_b200 import snapshot
        path = os.path.join(out_dir, f"snap_r{rank}.msgpack")
        snapshot.save(ctx, path, {}, include_optimizer_state=True)
        adam = snapshot.read(path)["snapshot"]["optimizer"]["nested"]["nested"]
        saved["snap_steps_sum"] = np.array(float(np.frombuffer(adam["param_steps_binary"], np.int32).astype(np.float64).sum()))
    # global-movement mode: every rank back-propagates its rows into its own 9 delta-gradient sums; they are all-reduced (48 bytes)
    # and each rank steps its replica of the DeltaNetwork's parameters
    ctx.begin_global_movement()
    for _ in range(6):
        parallel.train_step(ctx, rank, world)
    gm = ctx.global_movement()
    saved["delta"] = np.concatenate([gm["delta_transition"], gm["delta_rotation_6d"]])
    np.savez(os.path.join(out_dir, f"{mode}_r{rank}.npz"), **saved)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.fixture(scope="module")
def runs(tmp_path_factory):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(torch.cuda.device_count(), 8)            # every GPU of the box: 2 on the small lease, 8 on the full one
    out = tmp_path_factory.mktemp("dp")
    for mode in ("nccl", "p2p", "p2p_inorder"):
        mp.spawn(_worker, args=(world, _free_port(), str(out), mode, 42), nprocs=world, join=True)   # past step 32: every other step runs on prefetched samples
    for mode in ("p2p_capped", "nccl_capped"):
        mp.spawn(_worker, args=(world, _free_port(), str(out), mode, 34), nprocs=world, join=True)
    mp.spawn(_worker, args=(1, _free_port(), str(out), "single_capped", 34, world), nprocs=1, join=True)
    runs = {f"{m}_r{r}": np.load(out / f"{m}_r{r}.npz") for m in ("nccl", "p2p", "p2p_inorder", "p2p_capped", "nccl_capped") for r in range(world)}
    runs["single_capped_r0"] = np.load(out / "single_capped_r0.npz")
    runs["world"] = world
    return runs


def test_ranks_hold_identical_parameters(runs):
    for m in ("nccl", "p2p"):
        for r in range(1, runs["world"]):
            assert np.array_equal(runs[f"{m}_r0"]["params"], runs[f"{m}_r{r}"]["params"]), (m, r)
            assert np.array_equal(runs[f"{m}_r0"]["ema"], runs[f"{m}_r{r}"]["ema"]), (m, r)
            assert np.array_equal(runs[f"{m}_r0"]["stats"], runs[f"{m}_r{r}"]["stats"]), (m, r)
            assert np.array_equal(runs[f"{m}_r0"]["params_end"], runs[f"{m}_r{r}"]["params_end"]), (m, r)     # still identical after the pipelined steps


def test_pipelined_steps_size_the_batch_like_the_in_order_step(runs):
    """From step 33 on every other step runs on samples generated during the previous step (config.prefetch_next_samples), sized from the
    ALL-REDUCED counts read back mid-step.  Rays traced, rays kept and sample counts must follow the in-order run (a rank sizing its rays
    from local counts would be off by the world size); the two runs differ only by the order of fp32 atomics."""
    a, b = runs["p2p_r0"]["stats"], runs["p2p_inorder_r0"]["stats"]
    assert len(a) == len(b) == 42
    assert np.array_equal(a[:2], b[:2])
    assert np.all(np.abs(a - b) <= 0.05 * np.maximum(b, 1) + 2), np.abs(a - b).max(0)


def test_fused_peer_optimizer_equals_allreduce_then_adam(runs):
    a, b = runs["nccl_r0"], runs["p2p_r0"]
    # integer step statistics (rays traced, rays kept, samples, compacted samples): identical while the two trajectories are still
    # within rounding of each other, then within a percent (one flipped occupancy bit or saturation test moves a few samples)
    assert np.array_equal(a["stats"][:2], b["stats"][:2])
    assert np.all(np.abs(a["stats"][:24] - b["stats"][:24]) <= 0.02 * np.maximum(a["stats"][:24], 1) + 2), np.abs(a["stats"][:24] - b["stats"][:24]).max(0)
    # the backward pass accumulates with fp32 atomics whose order differs run to run, so two runs agree to rounding, not to
    # the bit; the two optimizer paths must stay as close as two runs of the same path do
    from tests import util as U
    for key in ("params", "ema"):
        d = U.half_ulp_diff(a[key], b[key])
        same = (d == 0).mean()
        assert same > 0.9 and np.percentile(d, 99.9) <= 16, (key, same, np.percentile(d, [99, 99.9, 100]))
    changed = (a["params"] != 0).mean()
    assert changed > 0.5 and a["stats"][-1][3] > 0


def test_shards_partition_the_parameters(runs):
    W = runs["world"]
    shards = [runs[f"p2p_r{r}"]["shard"] for r in range(W)]
    block = int(shards[0][0])
    assert block == 1024 and all(int(s[0]) == block for s in shards) and sum(int(s[1]) for s in shards) == 10559396
    # each rank's fp32 master copy of the blocks it owns (b % W == r) rounds to the fp16 parameters every rank holds
    idx = np.arange(10559396) // block
    for r in range(W):
        m = runs[f"p2p_r{r}"]["master_shard"]
        assert m.size == int(shards[r][1])
        assert np.array_equal(m.astype(np.float16).view(np.uint16), runs["p2p_r0"]["params"][idx % W == r])


def test_delta_network_replicas_agree(runs):
    for m in ("nccl", "p2p"):
        d0 = runs[f"{m}_r0"]["delta"]
        assert np.abs(d0 - np.array([0, 0, 0, 1, 0, 0, 0, 1, 0])).max() > 1e-5        # it moved
        for r in range(1, runs["world"]):
            assert np.array_equal(d0.view(np.uint32), runs[f"{m}_r{r}"]["delta"].view(np.uint32)), (m, r)
    assert np.abs(runs["nccl_r0"]["delta"] - runs["p2p_r0"]["delta"]).max() < 2e-4    # 6 Adam steps of 1e-4: same direction on both paths


def test_reported_losses_are_those_of_the_global_batch(runs):
    """Every 16th step reports rgb / eikonal / mask losses.  Data-parallel they must be the losses of the GLOBAL batch, the same bits on
    every rank (partial sums exchanged before the read-back, summed in rank order), i.e. what one GPU stepping the same global batch
    reports: the same rays at step 0 (same parameters, only the summation order differs), within the run-to-run drift of the fp32
    atomics later.  The runs compared sit at max_rays_per_batch: every rank traces exactly as many rays as its per-ray arrays hold
    (max_rays_per_batch / W) -- a reduction or zero-fill sized by the GLOBAL ray count would run off their end."""
    W = runs["world"]
    one = runs["single_capped_r0"]
    assert one["losses"][:, 0].tolist() == [0, 16, 32] and np.all(one["losses"][:, 1] > 0) and np.all(one["losses"][:, 3] > 0)
    assert one["stats"][-1][0] == 128 * W                                              # the cap binds
    for m in ("p2p_capped", "nccl_capped"):
        a = runs[f"{m}_r0"]
        assert a["stats"][-1][0] == 128 * W
        assert np.array_equal(a["stats"][0], one["stats"][0])                           # step 0: same rays traced / kept, same sample counts
        # later steps: the first pass's sample budget (the previous step's count; a rank's share of it here) truncates the ray list per
        # rank instead of globally, and the fp32 atomics reorder -- the counts agree to a few percent
        assert np.all(np.abs(a["stats"] - one["stats"]) <= 0.08 * np.maximum(one["stats"], 1) + 4), np.abs(a["stats"] - one["stats"]).max(0)
        l0 = a["losses"]
        assert l0[:, 0].tolist() == [0, 16, 32] and np.all(np.isfinite(l0))
        assert np.allclose(l0[0, 1:], one["losses"][0, 1:], rtol=1e-4), (m, l0[0], one["losses"][0])
        assert np.allclose(l0[:, 1:], one["losses"][:, 1:], rtol=0.1), (m, l0, one["losses"])
        for r in range(1, W):
            assert np.array_equal(l0, runs[f"{m}_r{r}"]["losses"]), (m, r)
            assert np.array_equal(a["params_end"], runs[f"{m}_r{r}"]["params_end"]), (m, r)
    for m in ("nccl", "p2p"):                                                           # uncapped runs: same bits on every rank
        for r in range(1, W):
            assert np.array_equal(runs[f"{m}_r0"]["losses"], runs[f"{m}_r{r}"]["losses"]), (m, r)


def test_sharded_optimizer_state_is_gathered_before_it_is_serialised(runs):
    W = runs["world"]
    for r in range(W):
        a = runs[f"p2p_r{r}"]
        assert a["refused"] == 1                                   # whole-state read of a sharded optimizer refuses ...
        assert a["stale_before_sync"] > 0.005                      # ... for a reason: its copy of the other ranks' blocks is stale (only the MLPs + the four coarsest grid levels train during the first 42 steps)
        assert a["master_rounds_to_params"]                        # after the gather the fp32 master rounds to the fp16 parameters everywhere
        assert np.array_equal(a["opt_digest"], runs["p2p_r0"]["opt_digest"])     # and every rank holds the same moments / step counters
        assert a["snap_steps_sum"] == a["opt_digest"][2] and a["snap_steps_sum"] > 0
