"""Multi-GPU plumbing (one process per GPU, torch.distributed): contiguous input slices per rank and
the gather-then-combine of the per-rank partials.  GT products and EC sums are not NCCL
reduction ops, so partials are all-gathered and combined locally in rank order."""
import numpy as np


def shard_bounds(n, rank, world):
    """[lo, hi) of rank's contiguous slice; slice sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_pairing_ip(ctx, g1_dev, g2_dev, n_local, out_ptr, scratch):
    """Product of pairings over vectors sharded across ranks.  g1_dev / g2_dev: this rank's slice (device);
    scratch = (partial, gathered) torch int32 tensors of 144 and world*144 elements on the same device."""
    import torch.distributed as dist

    partial, gathered = scratch
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        ctx.pairing_ip_dev(g1_dev, g2_dev, n_local, out_ptr)
        return
    ctx.miller_partial_dev(g1_dev, g2_dev, n_local, partial.data_ptr())
    dist.all_gather_into_tensor(gathered, partial)
    ctx.gt_combine_dev(gathered.data_ptr(), world, out_ptr)
