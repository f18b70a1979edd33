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
