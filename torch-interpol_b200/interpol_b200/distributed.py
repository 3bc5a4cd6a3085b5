"""Multi-GPU use of the hot path: one process per GPU (`torchrun`), NCCL (or
gloo on CPU for tests) through `torch.distributed`.

Every op is independent per batch element (reference: nd.py:95-106 only
broadcasts over batch), so the batch is sharded contiguously over ranks with
NO collective on the data path.  Two optional collectives exist:

* `gather_batch`   -- all-gather of the per-rank outputs when every rank wants
                      the full batch;
* `push_to_shared` -- several ranks splat different point sets into ONE target
                      volume: every rank scatters into a private volume, then a
                      single all-reduce(SUM) over NVLink/NVSwitch combines them.
"""
import torch


def shard_bounds(batch, world_size, rank):
    """Contiguous split of `batch` items: ranks < batch % world get one more."""
    base, rem = divmod(int(batch), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t, world_size=None, rank=None, dim=0):
    """This rank's contiguous slice of `t` along `dim` (a view, no copy)."""
    import torch.distributed as dist
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(t.shape[dim], world_size, rank)
    return t.narrow(dim, lo, hi - lo)


def gather_batch(local, batch, group=None):
    """All-gather per-rank batch shards (possibly of unequal length) into the
    full (batch, ...) tensor on every rank."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(batch, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    padded = local
    if local.shape[0] != width:
        pad = local.new_zeros((width - local.shape[0],) + tuple(local.shape[1:]))
        padded = torch.cat([local, pad], 0)
    out = local.new_empty((world * width,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    parts = [out[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, 0)


def sharded(op, *tensors, batch=None, gather=False, group=None, **kwargs):
    """Run `op(*shards, **kwargs)` on this rank's slice of the batch axis of
    every tensor argument (tensors whose leading size is 1 are broadcast)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if batch is None:
        batch = max(t.shape[0] for t in tensors)
    lo, hi = shard_bounds(batch, world, rank)
    shards = [t if t.shape[0] == 1 else t[lo:hi] for t in tensors]
    local = op(*shards, **kwargs)
    return gather_batch(local, batch, group) if gather else local


def push_to_shared(push_fn, input, grid, shape, group=None, **kwargs):
    """Splat this rank's points (`input`, `grid`) into a volume shared by all
    ranks: private scatter + one all-reduce(SUM).  Adjoint of every rank pulling
    from the same volume."""
    import torch.distributed as dist
    out = push_fn(input, grid, shape, **kwargs)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out
