"""World-size-2 gloo test of the multi-GPU host logic (lra_b200/shard.py): shards are contiguous, cover every job once,
per-rank results concatenated in rank order equal the unsharded result, and the time reduction is a max."""
import os
import socket
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lra_b200 import shard


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        sys.path.insert(0, p)
    from oracle import pyoracle as po
    import jobgen
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(77)           # every rank sees the same batch
    qa, ta, qo, to, ql, tl, k = jobgen.batch(rng, 240)
    lo, hi = shard.my_shard(ql.astype(np.int64) + tl, rank, world)
    s, nb, off, blk, st = po.aog_batch_port(qa, ta, qo[lo:hi].copy(), to[lo:hi].copy(), ql[lo:hi].copy(), tl[lo:hi].copy(),
                                            k[lo:hi].copy(), 4, -3, -4)
    counts = shard.gather_counts(hi - lo, dist)
    tmax = shard.max_over_ranks(1.0 + rank, dist)
    flat = np.concatenate([blk[off[j]:off[j] + nb[j]] for j in range(hi - lo)] or [np.zeros((0, 3), np.uint32)])
    q.put((rank, lo, hi, s, nb, flat, counts, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps: p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for p in ps: p.join(60)
    import jobgen
    from oracle import pyoracle as po
    rng = np.random.default_rng(77)
    qa, ta, qo, to, ql, tl, k = jobgen.batch(rng, 240)
    s, nb, off, blk, st = po.aog_batch_port(qa, ta, qo, to, ql, tl, k, 4, -3, -4)
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 240        # contiguous cover
    assert res[0][6] == res[1][6] == [res[0][2] - res[0][1], res[1][2] - res[1][1]]
    assert res[0][7] == res[1][7] == 2.0                                          # max over ranks
    assert (np.concatenate([res[0][3], res[1][3]]) == s).all()
    assert (np.concatenate([res[0][4], res[1][4]]) == nb).all()
    flat = np.concatenate([blk[off[j]:off[j] + nb[j]] for j in range(240)])
    assert (np.concatenate([res[0][5], res[1][5]]) == flat).all()


def test_shard_bounds_balance():
    w = np.array([10, 10, 10, 1000, 10, 10, 10, 10], dtype=np.int64)
    b = shard.shard_bounds(w, 4)
    assert b[0] == 0 and b[-1] == len(w) and (np.diff(b) >= 0).all()
    assert shard.shard_bounds(np.ones(8, np.int64), 8).tolist() == list(range(9))


def _worker_sg(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    lens = rng.integers(1, 400, size=97).astype(np.int64); lens[40] = 9000
    ascii_all = rng.integers(65, 90, size=int(lens.sum())).astype(np.uint8)
    a, l, (lo, hi), b = shard.scatter_batch(torch.from_numpy(ascii_all) if rank == 0 else None, torch.from_numpy(lens) if rank == 0 else None, dist)
    # "map" a shard: per read (first base, length) and a variable-length record (the read reversed)
    off = np.concatenate([[0], np.cumsum(l.numpy())])
    first = torch.tensor([int(a[off[i]]) for i in range(hi - lo)], dtype=torch.int32)
    rev = torch.cat([a[off[i]:off[i + 1]].flip(0) for i in range(hi - lo)]) if hi > lo else torch.zeros(0, dtype=torch.uint8)
    got = shard.gather_parts([first, l.to(torch.int64), rev], dist)
    q.put((rank, lo, hi, [[p.numpy() for p in g] for g in got] if got is not None else None))
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_reads_gather_records_two_ranks():
    """The exchange steps of the N > 1 path: rank 0's batch is dealt in base-balanced shards, every rank 'maps' its shard, rank 0 gets the
    records back in input order."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker_sg, args=(r, world, port, q)) for r in range(world)]
    for p in ps: p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for p in ps: p.join(60)
    rng = np.random.default_rng(5)
    lens = rng.integers(1, 400, size=97).astype(np.int64); lens[40] = 9000
    ascii_all = rng.integers(65, 90, size=int(lens.sum())).astype(np.uint8)
    off = np.concatenate([[0], np.cumsum(lens)])
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 97 and res[1][3] is None
    g = res[0][3]
    first = np.concatenate([g[0][0], g[1][0]]); ln = np.concatenate([g[0][1], g[1][1]]); rev = np.concatenate([g[0][2], g[1][2]])
    assert (ln == lens).all() and (first == ascii_all[off[:-1]]).all()
    assert (rev == np.concatenate([ascii_all[off[i]:off[i + 1]][::-1] for i in range(97)])).all()
    # base balance: the heavy read does not leave one rank with most of the bases
    assert abs(int(lens[:res[0][2]].sum()) - int(lens[res[0][2]:].sum())) <= 9000
