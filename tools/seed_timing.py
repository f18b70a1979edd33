#!/usr/bin/env python
"""Per-kernel timing of lra_b200_seed_batch (a2-a5) on synthetic reads vs a synthetic genome (development aid, GPU box).
The index is a plain sorted minimizer list of the genome (oracle StoreMinimizers, w = 10), a valid `genomemm`."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import synth, lra_b200
from oracle import pyoracle as po

glen = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
genome = synth.gen_ref(glen, 1, 78)[0][1]
t0 = time.time()
gt, gp = po.store_minimizers(genome, 17, 10)
order = np.argsort(gt & np.uint64(0x7FFFFFFFFFFFFFFF), kind="stable")
gt, gp = gt[order], gp[order]
print("index: %d minimizers in %.1f s" % (len(gt), time.time() - t0))
reads = synth.gen_reads([("chr1", genome)], n, "ont", 2)
read_len = np.array([len(r[1]) for r in reads], np.uint32)
read_off = np.zeros(n, np.uint64); read_off[1:] = np.cumsum(read_len[:-1])
arena = np.concatenate([r[1] for r in reads])
ctx = lra_b200.Context(0)
R = ctx.seq_upload(arena); G = ctx.seq_upload(genome); I = ctx.index_upload(gt, gp)
for it in range(3):
    t0 = time.time()
    o = ctx.seed_batch(R, G, I, read_off, read_len, 17, 10, 150)
    dt = time.time() - t0
    print("iter", it, "wall ms", round(dt * 1e3, 1), "reads/s", round(n / dt), "matches/read", o["n_matches"] / n, "mm/read", o["n_minimizers"].mean())
for s in ctx.kernel_stats():
    print("  %-24s %9.3f ms" % (s["name"], s["ms"]))
