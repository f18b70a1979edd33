"""Multi-GPU host logic: reads are independent, so a batch is dealt to ranks in contiguous, base-balanced shards and the
records come back in input order.  No data-path collective; torch.distributed is used for the barrier, the max-over-ranks
timing and (optionally) gathering per-rank record counts.  Mirrors the reference's only parallelism, data parallelism
over reads (lra.cpp:678-714), with ranks in place of pthreads."""
import numpy as np


def shard_bounds(weights, world):
    """Split items with the given weights (bases per read) into `world` contiguous shards of near-equal total weight.
    Returns an int64 array of world+1 boundaries."""
    w = np.asarray(weights, dtype=np.int64)
    c = np.concatenate([[0], np.cumsum(w)])
    targets = c[-1] * np.arange(1, world, dtype=np.float64) / world
    cuts = np.searchsorted(c, targets, side="left")
    b = np.concatenate([[0], cuts, [len(w)]]).astype(np.int64)
    return np.maximum.accumulate(b)


def my_shard(weights, rank, world):
    b = shard_bounds(weights, world)
    return int(b[rank]), int(b[rank + 1])


def max_over_ranks(value, dist=None, device=None):
    """Max of a python float over all ranks (time-like metrics are reported as the slowest rank's)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_counts(n, dist=None, device=None):
    """All ranks' record counts (used to place variable-length results in input order)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(n)]
    import torch
    t = torch.tensor([int(n)], dtype=torch.int64, device=device or "cpu")
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(x[0]) for x in out]
