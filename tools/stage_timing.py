#!/usr/bin/env python
"""Stand-alone timing of the small MapRead stages that are not in bench.py yet (development aid; run on the GPU box):
a6 anchor sorts, a7 CleanOffDiagonal, a11 chain splitting, a15 LinearExtend, a16 chain filters, a20 RefineBreakpoint, a24 GlobalChain, on ONT-shaped synthetic batches
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

# a15 LinearExtend (low-accuracy): 16384 reads, one cluster each of ~3 k K-mer anchors along the read's true alignment (ONT: the exact >= 17-base stretches)
import lextgen, spchaingen, cgluegen
RL = min(R, 4096)
arena, items = lextgen.reads(5, 64, read_len=20000, single=False, sub=0.04, indel=0.04, max_parts=2)
items = (items * (RL // 64 + 1))[:RL]
read_arena, ep = lextgen.to_batch(items)
rs = ctx.seq_upload(read_arena[:-16]); gs = ctx.seq_upload(arena[:-16])
for it in range(2):
    t0 = time.time(); o = ctx.linear_extend_batch(rs, gs, ep, 17, 0, 1); report("a15 LinearExtend+Trim (%d reads)" % RL, len(ep["q"]), t0)
print("   %d anchors -> %d extended" % (len(ep["q"]), len(o["q"])))
carena, citems = lextgen.chain_reads(6, 64, read_len=20000)
citems = (citems * (RL // 64 + 1))[:RL]
cra, cd = lextgen.chains_to_batch(citems)
rs2 = ctx.seq_upload(cra[:-16]); gs2 = ctx.seq_upload(carena[:-16])
for it in range(2):
    t0 = time.time(); o = ctx.linear_extend_chains_batch(rs2, gs2, cd, 17, 1, 1, 100); report("a15 LinearExtend_chain+Merge (%d reads)" % RL, len(cd["q"]), t0)

# a11: one UltimateChain per read
chs = spchaingen.chains(7, 512) * (R // 512)
c_off = np.zeros(len(chs) + 1, np.uint64); c_off[1:] = np.cumsum([len(c["q"]) for c in chs])
cat = lambda k, dt: np.concatenate([np.asarray(c[k], dt) for c in chs])
ac = dict(c_off=c_off, q=cat("q", np.uint32), t=cat("t", np.uint32), len=cat("len", np.int32), strand=cat("strand", np.uint8), cnum=cat("cnum", np.int32),
          link=np.concatenate([np.append(c["link"], 0).astype(np.uint8)[:len(c["q"])] for c in chs]))
for it in range(2):
    t0 = time.time(); o = ctx.split_chains_batch(ac, spchaingen.HDR, 50000, 0); report("a11 SPLITChain (%d anchors)" % int(c_off[-1]), len(chs), t0)
rg = np.random.default_rng(3)
mc = [cgluegen.merge_case(rg) for _ in range(512)] * (R // 512)
sp, chrom, strand_, box, base_ = [], [], [], [], 0
for c in mc:
    sp.append(c["sp"] + base_); chrom.append(c["chrom"]); strand_.append(c["strand"]); box.append(c["box"]); base_ += len(c["chrom"])
so = np.zeros(len(mc) + 1, np.uint64); so[1:] = np.cumsum([len(c["sp"]) for c in mc])
for it in range(2):
    t0 = time.time(); o = ctx.merge_chain_batch(np.concatenate(sp), so, np.concatenate(chrom), np.concatenate(strand_), np.concatenate(box)); report("a11 MergeChain", len(mc), t0)
