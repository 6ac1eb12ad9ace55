"""Data-parallel training step: rays sharded across ranks, NCCL (or gloo) all-reduce of the gradient buffer.

The reference is single-GPU (no collective anywhere); this is the build-out BASELINE.json asks for.  A ray's
image / pixel / jitter / background colour are pure functions of its GLOBAL index and the step's RNG state
(src/testbed_nerf.cu:1241-1261,1304,1632), so rank r simply traces global rays [r*R/W, (r+1)*R/W) and back-propagates its
own B/W compacted rows; every loss normaliser uses the global ray count and batch size, hence the SUM of the per-rank
gradients is the gradient of the global batch, and one all-reduce(sum) before the fused Adam keeps all replicas identical.
"""


def shard_range(n_total, rank, world_size):
    """Contiguous shard [begin, end) of n_total items for `rank`; n_total must divide evenly (rays_per_batch is kept a
    multiple of 128*world_size by the context)."""
    assert n_total % world_size == 0, (n_total, world_size)
    per = n_total // world_size
    return rank * per, (rank + 1) * per


def allreduce_sum_(tensor, group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def train_step(ctx, rank=0, world_size=1, group=None):
    """One data-parallel Testbed::train() step on this rank's Context.  Returns the step stats (global counts)."""
    import torch
    if world_size == 1:
        return ctx.train_step()
    ctx.train_step_begin(rank, world_size)
    grads = ctx.buffers()[3]
    counts = torch.tensor(ctx.local_counts(), dtype=torch.int64, device=grads.device)
    allreduce_sum_(grads, group)      # 10,559,396 fp32 gradients over NVLink
    allreduce_sum_(counts, group)     # {rays kept, samples, compacted samples}: drives the shared rays_per_batch update
    return ctx.train_step_end([int(x) for x in counts.tolist()])
