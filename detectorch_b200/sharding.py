"""Multi-GPU: images are independent, so the batch is sharded across ranks (one process per GPU) with no
data-path collective; torch.distributed is only used to (optionally) gather the padded per-image results.
Works with nccl (GPU tensors) and gloo (CPU tensors, used by the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(total, world, rank):
    """Contiguous, balanced [begin, end) slice of `total` images for `rank` (first `total % world` ranks get one more)."""
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_results(local, group=None):
    """local: dict of tensors whose dim 0 is this rank's image shard.  Returns (on every rank) the dict of tensors
    concatenated in global image order.  Shards may have different sizes (padded all_gather)."""
    world = dist.get_world_size(group)
    out = {}
    n_local = torch.tensor([next(iter(local.values())).size(0)], dtype=torch.int64, device=next(iter(local.values())).device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    for k, t in local.items():
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.size(0)] = t
        parts = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out[k] = torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)
    return out
