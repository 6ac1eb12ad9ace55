"""world_size-2 gloo test of the data-parallel scheme (CPU): ray sharding by global index + all-reduce(sum) of gradients
reproduces the single-rank result.  Compute is the CPU oracle (this is host-logic coverage; the GPU path is in bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from tests import util as U
from neus2_b200 import parallel, synthetic


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene():
    data = synthetic.render_sphere_views(n_views=3, width=48, height=40, seed=3)
    grid = np.zeros(128 ** 3, dtype=np.float32)
    idx = np.arange(128 ** 3, dtype=np.uint32)

    def dec(x):
        x = x & 0x49249249
        x = (x | (x >> 2)) & 0xc30c30c3
        x = (x | (x >> 4)) & 0x0f00f00f
        x = (x | (x >> 8)) & 0xff0000ff
        x = (x | (x >> 16)) & 0x0000ffff
        return x
    pts = (np.stack([dec(idx), dec(idx >> 1), dec(idx >> 2)], 1) + 0.5) / 128
    grid[np.abs(synthetic.sphere_sdf(pts, data["spheres"])) < 0.02] = 1.0
    bitfield, _ = O.update_bitfield(grid, 1)
    return data, bitfield


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = O.make_cfg(n_levels=4, base_resolution=8, log2_hashmap_size=12, per_level_scale=1.6)
    lay = O.param_layout(cfg)
    p = U.random_params(cfg, seed=5)
    ph = U.to_half_bits(p)
    data, bitfield = _scene()
    ds = O.make_dataset(data["pixels"], data["focal"], data["principal"], data["xforms"])
    R, B = 256, 1 << 14
    rng = O.Pcg32(seed=1337)
    # every rank can generate the whole batch (deterministic); it keeps only its shard of the GLOBAL ray indices
    gen = O.generate_training_samples(R, 1 << 16, 0, rng, ds, bitfield)
    lo, hi = parallel.shard_range(R, rank, world)
    keep = [i for i in range(gen["n_kept"]) if lo <= gen["ray_indices"][i] < hi]
    rows = np.concatenate([np.arange(gen["numsteps"][i][1], gen["numsteps"][i][1] + gen["numsteps"][i][0]) for i in keep]) if keep else np.zeros(0, int)
    coords = np.ascontiguousarray(gen["coords"][rows])
    rs = np.random.RandomState(0)
    dL_all = np.zeros((gen["n_samples"], 16), dtype=np.float16)
    dL_all[:, :11] = rs.normal(0, 1e-2, (gen["n_samples"], 11)).astype(np.float16)
    dL = np.ascontiguousarray(dL_all[rows]).view(np.uint16)
    g = O.net_backward(cfg, coords, ph, dL, B) if len(rows) else np.zeros(lay["total"], np.float32)
    var_local = g[lay["variance"]]
    t = torch.from_numpy(g.copy())
    parallel.allreduce_sum_(t)
    counts = torch.tensor([len(keep), len(rows)], dtype=torch.int64)
    parallel.allreduce_sum_(counts)
    if rank == 0:
        np.savez(os.path.join(out_dir, "dp.npz"), grads=t.numpy(), counts=counts.numpy(), var_local=var_local)
    dist.destroy_process_group()


def test_dp2_gradients_equal_single_rank(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "dp.npz")
    # single rank reference
    cfg = O.make_cfg(n_levels=4, base_resolution=8, log2_hashmap_size=12, per_level_scale=1.6)
    lay = O.param_layout(cfg)
    p = U.random_params(cfg, seed=5)
    data, bitfield = _scene()
    ds = O.make_dataset(data["pixels"], data["focal"], data["principal"], data["xforms"])
    gen = O.generate_training_samples(256, 1 << 16, 0, O.Pcg32(seed=1337), ds, bitfield)
    assert gen["n_kept"] > 10
    rs = np.random.RandomState(0)
    n = gen["n_samples"]
    dL = np.zeros((n, 16), dtype=np.float16)
    dL[:, :11] = rs.normal(0, 1e-2, (n, 11)).astype(np.float16)
    ref = O.net_backward(cfg, np.ascontiguousarray(gen["coords"][:n]), U.to_half_bits(p), dL.view(np.uint16), 1 << 14)
    assert list(got["counts"]) == [gen["n_kept"], n]                          # shards partition the rays exactly
    v = lay["variance"]
    a, b = got["grads"][:v], ref[:v]
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-5                   # linear in the batch: only fp32 summation order differs
    assert abs(got["grads"][v] - ref[v]) <= 2e-3 * abs(ref[v]) + 1e-6         # per-rank half rounding of the variance gradient


def test_shard_range_partitions():
    for world in (1, 2, 4, 8):
        R = 128 * world * 3
        spans = [parallel.shard_range(R, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == R
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def test_interleaved_rows_cover_everything_once():
    for n in (5, 16, 1080):
        for world in (1, 2, 4, 8):
            rows = sorted(sum((list(range(a, b)) for r in range(world) for a, b in parallel.interleaved_rows(n, r, world, 4)), []))
            assert rows == list(range(n))


def test_split_rows_covers_everything_once():
    for n in (1, 7, 16, 1080, 513):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.split_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


class _FakeCtx:
    """Stands in for neus2_b200.Context on CPU: 'renders' a tile / 'evaluates' a slab as a pure function of the global
    pixel / lattice index, which is the property the real kernels have (tests/test_gpu_render_mesh.py proves it on the GPU)."""

    def render(self, p, spp=1):
        ry = torch.arange(p.tile_h)
        rows = p.tile_y0 + ((ry // p.row_block) * p.row_block * p.row_stride + ry % p.row_block if getattr(p, "row_block", 0) else ry)
        ys = rows.to(torch.float32)[:, None, None]
        xs = torch.arange(p.tile_x0, p.tile_x0 + p.tile_w, dtype=torch.float32)[None, :, None]
        return (ys * 1000 + xs + torch.arange(4, dtype=torch.float32)[None, None, :] * 0.25) * spp

    def sdf_lattice(self, res, aabb, z0=0, z1=None, use_inference_params=True):
        z = torch.arange(z0, z1, dtype=torch.float32)[:, None, None]
        y = torch.arange(res[1], dtype=torch.float32)[None, :, None]
        x = torch.arange(res[0], dtype=torch.float32)[None, None, :]
        return z * 10000 + y * 100 + x


class _P:
    width, height = 13, 11
    tile_x0 = tile_y0 = tile_w = tile_h = 0
    row_block = row_stride = 0


def _infer_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = _FakeCtx()
    img = parallel.render(ctx, _P(), spp=2, rank=rank, world_size=world, chunk=2)
    lat = parallel.sdf_lattice(ctx, (5, 4, 7), rank=rank, world_size=world)
    torch.save(dict(img=img, lat=lat), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_render_bands_and_lattice_slabs_reassemble(tmp_path, world):
    port = _free_port()
    mp.spawn(_infer_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    p = _P()
    p.tile_w, p.tile_h = p.width, p.height
    whole_img = _FakeCtx().render(p, spp=2)
    whole_lat = _FakeCtx().sdf_lattice((5, 4, 7), None, 0, 7)
    for r in range(world):
        got = torch.load(tmp_path / f"r{r}.pt")
        assert torch.equal(got["img"], whole_img) and torch.equal(got["lat"], whole_lat)


class _FakeTrainCtx:
    """Records the order in which parallel.train_step drives a context, with CPU tensors standing in for the device buffers."""

    def __init__(self, rank, peers, delta, due=True):
        self.rank, self.peers_attached, self.global_movement_active = rank, peers, delta
        self.calls = []
        self.counts = torch.tensor([10 + rank, 100 + rank, 50 + rank, 0], dtype=torch.int32)
        self.grads = torch.full((8,), float(rank + 1))
        self.dgrads = torch.arange(12, dtype=torch.float32) * (rank + 1)
        self.loss_sums = torch.tensor([1.0, 2.0, 3.0, 0.0]) * (rank + 1)
        self.due = due

        class _Cfg:
            loss_scale = 128.0
        self.cfg = _Cfg()

    def train_step_begin(self, rank, world):
        self.calls.append(("begin", rank, world))

    def counts_device(self):
        return self.counts

    def train_step_counts_ready(self):
        self.calls.append(("counts_ready", self.counts.tolist(), self.loss_sums.tolist()))

    def loss_scalars_device(self):
        return self.loss_sums, self.due

    def peer_exchange(self):
        self.calls.append(("exchange",))

    def peer_barrier(self):
        self.calls.append(("barrier",))

    def delta_gradients(self):
        return self.dgrads

    def optimizer_step_sharded(self, loss_scale):
        self.calls.append(("sharded", self.dgrads.tolist()))

    def buffers(self):
        return None, None, None, self.grads

    def train_step_end(self):
        self.calls.append(("end", self.grads.tolist()))
        return "stats"


def _train_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = {}
    for peers in (True, False):
        ctx = _FakeTrainCtx(rank, peers, delta=True)
        assert parallel.train_step(ctx, rank, world) == "stats"
        out[peers] = ctx.calls
    ctx = _FakeTrainCtx(rank, False, delta=False, due=False)
    assert parallel.train_step(ctx, rank, world) == "stats"
    out["not_due"] = ctx.calls
    torch.save(out, os.path.join(out_dir, f"t{rank}.pt"))
    dist.destroy_process_group()


def test_data_parallel_step_choreography(tmp_path):
    """world_size 2 over gloo.  Peer-memory path: no collective library call at all -- exchange (counts / loss sums / delta gradients
    summed on the device) -> host read-back point -> sharded optimizer -> barrier -> end.  Collective path: the sample counts and, on
    reporting steps, the loss sums are all-reduced BEFORE the host read-back point (counts_ready), the 12 delta gradients and the 42 MB
    gradient buffer before the replicated optimizer inside train_step_end."""
    port = _free_port()
    mp.spawn(_train_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = torch.load(tmp_path / f"t{r}.pt")
        p2p, nccl, not_due = got[True], got[False], got["not_due"]
        assert [c[0] for c in p2p] == ["begin", "exchange", "counts_ready", "sharded", "barrier", "end"]
        assert [c[0] for c in nccl] == ["begin", "counts_ready", "end"]
        assert p2p[0] == ("begin", r, 2)
        assert p2p[2][1] == [10 + r, 100 + r, 50 + r, 0]                                 # peer path: the sums are taken by the exchange kernel, not here
        assert p2p[5][1] == [float(r + 1)] * 8                                           # ... and no gradient all-reduce
        assert nccl[1][1] == [21, 201, 101, 0]                                           # global counts at the read-back point
        assert nccl[1][2] == [3.0, 6.0, 9.0, 0.0]                                        # global loss sums at the read-back point (reporting step)
        assert nccl[2][1] == [3.0] * 8                                                   # gradients summed before Adam
        assert not_due[1][1] == [21, 201, 101, 0] and not_due[1][2] == [1.0 * (r + 1), 2.0 * (r + 1), 3.0 * (r + 1), 0.0]   # no loss all-reduce off the reporting steps
    # single rank: one call, no collectives
    class _One:
        def train_step(self):
            return "one"
    assert parallel.train_step(_One(), 0, 1) == "one"


class _FakeShardedCtx:
    """CPU stand-in for a context whose sharded optimizer keeps only ITS parameter blocks current (blocks of `block` parameters dealt
    round-robin over the ranks): everything this rank does not own holds a stale marker."""
    peers_attached = True

    def __init__(self, rank, world, n, block):
        self.rank, self.world, self.block = rank, world, block
        idx = torch.arange(n)
        own = (idx // block) % world == rank
        truth = idx.float() * 0.5 + 1.0
        self.master = torch.where(own, truth, torch.full((n,), -1.0))
        self.m1 = torch.where(own, truth * 2, torch.full((n,), -2.0))
        self.m2 = torch.where(own, truth * 3, torch.full((n,), -3.0))
        self.steps = torch.where(own, idx.int() + 7, torch.full((n,), -4, dtype=torch.int32))

    def buffers(self):
        return self.master, None, None, None

    def optimizer_buffers(self):
        return self.m1, self.m2, self.steps, 5

    def optimizer_shard(self, rank, world):
        return self.block, 0

    def get_state(self):
        return {"optimizer_step": 5}


def _sync_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import unittest.mock as mock
    n, block = 10 * 16 + 5, 16                       # ten full blocks and a partial one
    ctx = _FakeShardedCtx(rank, world, n, block)
    with mock.patch("torch.cuda.synchronize", lambda *a, **k: None):
        parallel.sync_optimizer_state(ctx, rank, world)
    torch.save(dict(master=ctx.master, m1=ctx.m1, m2=ctx.m2, steps=ctx.steps, synced=ctx._opt_synced_step), os.path.join(out_dir, f"s{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sync_optimizer_state_gathers_the_round_robin_blocks(tmp_path, world):
    """parallel.sync_optimizer_state over gloo: every rank ends up with every block's owner's values (fp32 master, both moments, step
    counters), including the last, partial block -- what snapshot.save / get_params need after a run with the sharded optimizer."""
    mp.spawn(_sync_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    n = 10 * 16 + 5
    idx = torch.arange(n)
    truth = idx.float() * 0.5 + 1.0
    for r in range(world):
        got = torch.load(tmp_path / f"s{r}.pt")
        assert torch.equal(got["master"], truth) and torch.equal(got["m1"], truth * 2) and torch.equal(got["m2"], truth * 3)
        assert torch.equal(got["steps"], idx.int() + 7) and got["synced"] == 5
