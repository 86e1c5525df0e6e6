"""Multi-GPU plumbing: clip batches shard across ranks (one process per GPU), no collective inside the model,
one gather of the solved poses at the end.  Works with the `nccl` backend on GPUs and
with `gloo` on CPU (the host logic is backend-agnostic and is what tests/test_distributed_cpu.py exercises)."""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank`; the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_clips(x, rank=None, world=None):
    """The slice of a (N, T, V, C) clip batch this rank solves."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_poses(local, n_total, group=None):
    """All ranks receive the (n_total, T', D) poses, in clip order.  Ragged shards are padded to the largest
    shard for the collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]
    m = max(sizes)
    pad = local
    if local.shape[0] < m:
        pad = torch.cat([local, local.new_zeros((m - local.shape[0],) + tuple(local.shape[1:]))])
    out = local.new_empty((world * m,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    assert sizes[rank] == local.shape[0]
    return torch.cat([out[r * m:r * m + sizes[r]] for r in range(world)])


def solve_sharded(model, x_all, group=None):
    """Each rank runs `model` on its shard of x_all (already on this rank's device) and all ranks get every pose."""
    xs = shard_clips(x_all, dist.get_rank(group), dist.get_world_size(group))
    poses = model(xs)["poses"]
    return gather_poses(poses, x_all.shape[0], group)
