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


# ---- scatter of read batches / gather of alignment records (the only exchange steps of the path: north_star) --------------
def _world(dist):
    if dist is None or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(), dist.get_world_size()


def scatter_batch(ascii_t, read_len_t, dist=None, device="cpu", src=0):
    """Rank `src` holds a batch of reads (uint8 tensor of all bases back to back on `device`, int64 tensor of read lengths); the others pass
    None.  Reads are dealt in contiguous base-balanced shards.  Every rank returns (ascii of its shard, int64 lengths of its shard, (lo, hi),
    bounds) -- tensors on `device`, received with point-to-point sends (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    import torch
    rank, world = _world(dist)
    if world == 1:
        return ascii_t, read_len_t, (0, int(read_len_t.numel())), np.array([0, int(read_len_t.numel())], np.int64)
    n = torch.tensor([int(read_len_t.numel()) if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src)
    lens = read_len_t.to(device=device, dtype=torch.int64) if rank == src else torch.empty(int(n[0]), dtype=torch.int64, device=device)
    dist.broadcast(lens, src)
    lens_np = lens.cpu().numpy()
    b = shard_bounds(lens_np, world)
    byte_off = np.concatenate([[0], np.cumsum(lens_np)]).astype(np.int64)
    lo, hi = int(b[rank]), int(b[rank + 1])
    if rank == src:
        ops = [dist.P2POp(dist.isend, ascii_t[int(byte_off[b[r]]):int(byte_off[b[r + 1]])].contiguous(), r) for r in range(world) if r != src and b[r + 1] > b[r]]
        mine = ascii_t[int(byte_off[lo]):int(byte_off[hi])]
    else:
        mine = torch.empty(int(byte_off[hi] - byte_off[lo]), dtype=torch.uint8, device=device)
        ops = [dist.P2POp(dist.irecv, mine, src)] if hi > lo else []
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return mine, lens[lo:hi], (lo, hi), b


def gather_parts(parts, dist=None, device="cpu", dst=0):
    """Every rank passes a list of 1-D tensors (its alignment records: per-read arrays, record bytes, cigar words).  Rank `dst` gets
    [parts of rank 0, parts of rank 1, ...] (rank order == input order of the shards); the others get None."""
    import torch
    rank, world = _world(dist)
    if world == 1:
        return [parts]
    sizes = torch.tensor([int(p.numel()) for p in parts], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    if rank == dst:
        out, ops = [], []
        for r in range(world):
            if r == dst:
                out.append(parts); continue
            bufs = [torch.empty(int(all_sizes[r][i]), dtype=parts[i].dtype, device=device) for i in range(len(parts))]
            out.append(bufs)
            ops += [dist.P2POp(dist.irecv, t, r) for t in bufs if t.numel()]
    else:
        out = None
        ops = [dist.P2POp(dist.isend, p.contiguous(), dst) for p in parts if p.numel()]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out
