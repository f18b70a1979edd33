#!/usr/bin/env python
"""Stand-alone timing of the small MapRead stages that are not in bench.py yet (development aid; run on the GPU box):
a6 anchor sorts, a7 CleanOffDiagonal, a16 chain filters, a20 RefineBreakpoint, a24 GlobalChain, on ONT-shaped synthetic batches
(16384 reads: two anchor lists of ~190 seeds per read, one chain of ~400 anchors per read, 2048 breakpoints, 2048 chaining problems)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import lra_b200

R = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rng = np.random.default_rng(1)
ctx = lra_b200.Context(0)


def report(tag, n_units, t0):
    dt = time.time() - t0
    ks = ctx.kernel_stats()
    print("%-28s wall %8.2f ms  kernel %8.3f ms  (%d units, %.2f M units/s kernel)" % (tag, dt * 1e3, sum(s["ms"] for s in ks), n_units, n_units / max(sum(s["ms"] for s in ks), 1e-6) / 1e3))


# anchor lists: per read and strand ~190 seeds on the read's diagonal plus noise
n_lists = 2 * R
sizes = rng.integers(100, 280, n_lists)
off = np.zeros(n_lists + 1, np.uint64); off[1:] = np.cumsum(sizes)
N = int(off[-1])
lid = np.repeat(np.arange(n_lists), sizes)
q = rng.integers(0, 20000, N).astype(np.uint32)
base = rng.integers(10_000_000, 2_900_000_000, n_lists)[lid]
noise = rng.random(N) < 0.1
t = (base + q.astype(np.int64) + rng.integers(-20, 20, N)).astype(np.int64)
t[noise] = rng.integers(0, 2_900_000_000, int(noise.sum()))
t = t.astype(np.uint32)
qt = q.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
strand = np.zeros(n_lists, np.uint8)
for it in range(2):
    t0 = time.time(); sq, st, perm = ctx.sort_matches_batch(0, q, t, off); report("a6 DiagonalSort", n_lists, t0)
opts = dict(cleanMaxDiag=200, minDiagCluster=3, bypassClustering=1, cleanClustersize=100, SecondCleanMinDiagCluster=10, punish_anchorfreq=5, anchorPerlength=5,
            SecondCleanMaxDiag=120, ExtractDiagonalFromClean=1, globalK=17)
hdr = np.append(np.arange(0, 3_000_000_000, 125_000_000, dtype=np.uint64), np.uint64(3_000_000_000))
for it in range(2):
    t0 = time.time(); o = ctx.clean_off_diagonal_batch(sq, st, qt[perm], off, strand, opts, hdr); report("a7 CleanOffDiagonal", n_lists, t0)
print("   kept %d of %d anchors, %d clusters" % (int(o["keep"].sum()), N, int(o["n_cl"].sum())))

# chains: ~400 anchors per read along a diagonal with small indels
csz = rng.integers(200, 600, R)
coff = np.zeros(R + 1, np.uint64); coff[1:] = np.cumsum(csz)
M = int(coff[-1]); cid = np.repeat(np.arange(R), csz)
step = rng.integers(20, 80, M)
cq = np.zeros(M, np.int64); cq[:] = np.cumsum(step); cq -= np.repeat(cq[coff[:-1].astype(np.int64)], csz)
ct = rng.integers(10_000_000, 2_900_000_000, R)[cid] + cq + np.cumsum(rng.choice([0, 0, 0, 0, 7, -7, 40, -40], M))
ln = rng.choice([17, 25, 40, 60, 120], M).astype(np.uint32)
for mode in (0, 1, 4, 5):
    t0 = time.time(); keep = ctx.chain_filter_batch(mode, cq.astype(np.uint32), ct.astype(np.uint32), ln, np.zeros(M, np.uint8), coff); report("a16 chain filter mode %d" % mode, R, t0)

import bpgen, chaingen
cs = bpgen.cases(3, n=24) * 86
fwd, rcs, gen, bp = bpgen.pack(cs)
f = ctx.seq_upload(fwd[:-16]); r = ctx.seq_upload(rcs[:-16]); g = ctx.seq_upload(gen[:-16])
for it in range(2):
    t0 = time.time(); o = ctx.refine_breakpoint_batch(f, r, g, bp); report("a20 RefineBreakpoint", len(cs), t0)

probs = [p for s in range(1, 60) for p in chaingen.problems(s, sizes=(40, 200, 400))] * 12
foff = np.zeros(len(probs) + 1, np.uint64); foff[1:] = np.cumsum([len(p[0]) for p in probs])
frag = np.concatenate([p[0] for p in probs]); sc = np.concatenate([p[1] for p in probs])
for it in range(2):
    t0 = time.time(); o = ctx.global_chain_batch(frag, foff, sc); report("a24 GlobalChain", len(probs), t0)
