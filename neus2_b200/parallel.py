"""Data-parallel training step: rays sharded across ranks, NCCL (or gloo) all-reduce of the gradient buffer.

The reference is single-GPU (no collective anywhere); this is the build-out BASELINE.json asks for.  A ray's
image / pixel / jitter / background colour are pure functions of its GLOBAL index and the step's RNG state
(src/testbed_nerf.cu:1241-1261,1304,1632), so rank r simply traces the global rays r, r + W, r + 2W, ... (interleaved: consecutive
ray indices walk through the images, and every rank should see every image) and back-propagates its own B/W compacted rows; every loss normaliser uses the global ray count and batch size, hence the SUM of the per-rank
gradients is the gradient of the global batch, and one all-reduce(sum) before the fused Adam keeps all replicas identical.
"""


def shard_range(n_total, rank, world_size):
    """Contiguous shard [begin, end) of n_total items for `rank`; n_total must divide evenly."""
    assert n_total % world_size == 0, (n_total, world_size)
    per = n_total // world_size
    return rank * per, (rank + 1) * per


def allreduce_sum_(tensor, group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def attach_peers(ctx, rank, world_size, group=None):
    """Exchange CUDA IPC handles of every rank's gradient / fp16 parameter / EMA buffers and map them (once, at set-up), so
    that the optimizer step can reduce gradients and publish parameters through NVLink peer memory."""
    import torch.distributed as dist
    handles = [None] * world_size
    dist.all_gather_object(handles, ctx.ipc_export(), group=group)
    ctx.ipc_attach(rank, world_size, handles)


def sync_optimizer_state(ctx, rank, world_size, group=None):
    """With the sharded optimizer (attach_peers) rank r keeps the fp32 master weights, both Adam moments and the per-parameter step
    counters current only for ITS slice of the parameters.  Before anything reads them whole (snapshot.save with optimizer state,
    Context.get_params) every rank calls this: each slice is broadcast from its owner, after which all ranks hold the complete,
    identical optimizer state.  Collective; checkpoint-time cost (4 buffers x world_size broadcasts)."""
    import torch
    import torch.distributed as dist
    if world_size > 1 and ctx.peers_attached:
        master = ctx.buffers()[0]
        m1, m2, steps, _ = ctx.optimizer_buffers()
        torch.cuda.synchronize(master.device)
        block, _ = ctx.optimizer_shard(rank, world_size)          # blocks of `block` parameters are dealt round-robin over the ranks
        n = master.numel()
        full = n // block
        for buf in (master, m1, m2, steps):
            body = buf[: full * block].view(full, block)
            for r in range(world_size):
                src = dist.get_global_rank(group, r) if group is not None else r
                part = body[r::world_size].contiguous()
                dist.broadcast(part, src=src, group=group)
                body[r::world_size].copy_(part)
            if n > full * block:                                  # the last, partial block
                r = full % world_size
                dist.broadcast(buf[full * block:], src=dist.get_global_rank(group, r) if group is not None else r, group=group)
        torch.cuda.synchronize(master.device)
    ctx._opt_synced_step = ctx.get_state()["optimizer_step"]


def train_step(ctx, rank=0, world_size=1, group=None):
    """One data-parallel Testbed::train() step on this rank's Context.  Returns the step stats (global counts, global losses).

    With peers attached (attach_peers) no collective library call is left on the step's path: the two cross-rank ordering points are
    flag exchanges over mapped peer memory (a few microseconds; the first one also carries the sums of the step's counts, loss sums and
    DeltaNetwork gradient sums), and the optimizer is fused with its collective: every rank sums its 1/W parameter slice of all ranks'
    gradients through NVLink peer loads inside the Adam kernel, stores the new fp16 parameters into every rank's buffers and clears the
    gradients it consumed -- no all-reduce of the 42 MB gradient buffer, 1/W of the optimizer-state traffic per GPU, no memsets.
    Without peers it falls back to NCCL (or gloo) all-reduces followed by the replicated Adam (same arithmetic)."""
    if world_size == 1:
        return ctx.train_step()
    ctx.train_step_begin(rank, world_size)
    # everything below is enqueued behind the backward pass on the step's stream: no host round trip between begin and end
    if ctx.peers_attached:
        ctx.peer_exchange()                      # all backward passes done; {rays kept, samples, compacted}, loss sums, delta sums summed over ranks
        ctx.train_step_counts_ready()            # pipelined step: the host read-back is taken here; the next step's samples are generated beside the optimizer
        ctx.optimizer_step_sharded(ctx.cfg.loss_scale)
        ctx.peer_barrier()                       # peers' stores have landed / peers are done reading before anyone starts the next step
    else:
        allreduce_sum_(ctx.counts_device(), group)
        sums, due = ctx.loss_scalars_device()
        if due:
            allreduce_sum_(sums, group)          # every 16th step: the reported losses are those of the global batch
        ctx.train_step_counts_ready()
        if ctx.global_movement_active:
            # the DeltaNetwork's 9 gradient sums (48 bytes): every rank then steps its replica of the 12 parameters identically
            allreduce_sum_(ctx.delta_gradients(), group)
        allreduce_sum_(ctx.buffers()[3], group)  # 10,559,396 fp32 gradients over NVLink
    return ctx.train_step_end()


# ---- inference: independent units, no data-path collective ---------------------------------------------------
def split_rows(n_rows, rank, world_size):
    """Rows [begin, end) of rank `rank` when n_rows image rows (or lattice z-layers) are dealt out in contiguous bands
    whose sizes differ by at most one (the first n_rows % world_size ranks take the extra row)."""
    base, extra = divmod(int(n_rows), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def _gather_bands(local, n_rows, rank, world_size, group=None):
    """All ranks contribute their band ([rows_r, ...], rows as in split_rows); every rank gets the stacked whole.
    Bands are padded to the largest band so that one all_gather moves everything (results only: no collective sits on the
    per-sample data path)."""
    import torch
    import torch.distributed as dist
    if world_size == 1:
        return local
    max_rows = split_rows(n_rows, 0, world_size)[1]
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world_size)]
    dist.all_gather(parts, pad, group=group)
    out = []
    for r in range(world_size):
        b, e = split_rows(n_rows, r, world_size)
        out.append(parts[r][: e - b])
    return torch.cat(out, dim=0)


def interleaved_rows(n_rows, rank, world_size, chunk=16):
    """Row chunks [(y0, y1), ...] of rank `rank`: the image is cut into chunks of `chunk` rows that are dealt round-robin, so
    that every GPU gets a similar share of the rows that actually hit the object (contiguous bands would give the GPUs
    holding the image centre most of the work)."""
    return [(y, min(y + chunk, n_rows)) for i, y in enumerate(range(0, int(n_rows), chunk)) if i % world_size == rank]


def render(ctx, params, spp=1, rank=0, world_size=1, group=None, chunk=16):
    """Testbed::render_to_cpu with the image rows dealt across GPUs in blocks of `chunk` rows, round-robin (tiles of a frame are
    independent: every pixel's ray, Sobol jitter and colour depend only on its own index).  A rank renders its whole share in ONE
    call (interleaved-row tile): the persistent march kernel needs thousands of rays in flight, a 16-row strip alone would be
    latency-bound.  Returns the whole [H, W, 4] image on every rank."""
    import torch
    import torch.distributed as dist
    if world_size == 1:
        params.tile_x0, params.tile_y0, params.tile_w, params.tile_h = 0, 0, params.width, params.height
        params.row_block, params.row_stride = 0, 0
        return ctx.render(params, spp=spp)
    mine = interleaved_rows(params.height, rank, world_size, chunk)
    if not mine:
        raise ValueError("more ranks than row chunks: lower `chunk`")
    params.tile_x0, params.tile_w = 0, params.width
    params.tile_y0, params.tile_h = mine[0][0], sum(y1 - y0 for y0, y1 in mine)
    params.row_block, params.row_stride = chunk, world_size
    local = ctx.render(params, spp=spp)
    rows_max = max(sum(y1 - y0 for y0, y1 in interleaved_rows(params.height, r, world_size, chunk)) for r in range(world_size))
    pad = torch.zeros((rows_max, params.width, 4), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world_size)]
    dist.all_gather(parts, pad, group=group)
    out = torch.empty((params.height, params.width, 4), dtype=local.dtype, device=local.device)
    for r in range(world_size):
        off = 0
        for (y0, y1) in interleaved_rows(params.height, r, world_size, chunk):
            out[y0:y1] = parts[r][off:off + (y1 - y0)]
            off += y1 - y0
    return out


def sdf_lattice(ctx, res, aabb=(0, 0, 0, 1, 1, 1), rank=0, world_size=1, group=None, use_inference_params=True):
    """Testbed::get_density_on_grid with the lattice split into z-slabs, one per GPU; every rank gets the whole lattice
    ([rz, ry, rx] float32), so marching cubes (milliseconds) can run anywhere without a seam."""
    z0, z1 = split_rows(res[2], rank, world_size)
    slab = ctx.sdf_lattice(res, aabb, z0=z0, z1=z1, use_inference_params=use_inference_params)
    return _gather_bands(slab, res[2], rank, world_size, group)
