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

# a14 RefineSpace (both branches) and a17 RefineByLinearAlignment: spaces / gaps between anchors, a8 SplitRoughClustersWithGaps
import test_refine_space as TRS, test_refine_linear as TRL, roughgen
contig, reads, sp = TRS.large_spaces(7, 256)
rng2 = np.random.default_rng(8)
for x in sp[::3]:
    x["qe"] = min(x["qs"] + int(rng2.integers(0, 900)), len(reads[x["read"]])); x["te"] = x["ts"] + max(0, (x["qe"] - x["qs"]) + int(rng2.integers(-20, 21)) - x["lrlength"])
d = TRS.space_dict(reads, sp)
d = {k: np.tile(np.asarray(v), 16) for k, v in d.items()}          # 4096 spaces (the same 256, sixteen times: the work is what counts)
rs3 = ctx.seq_upload(np.concatenate(reads)); gs3 = ctx.seq_upload(contig)
for it in range(2):
    t0 = time.time(); o = ctx.refine_space_batch(rs3, gs3, d, 17, 4, -3, -4, W=10, local_max_freq=30)
    dt = time.time() - t0
    print("%-28s wall %8.2f ms  (%d spaces, %d pairs)" % ("a14 RefineSpace (mixed)", dt * 1e3, len(d["qs"]), int(o["n_pairs"].sum())))
contig2, reads2, gl = TRL.gaps(2, 1500)
roff = np.zeros(len(reads2), np.int64); roff[1:] = np.cumsum([len(r) for r in reads2[:-1]])
g = dict(cur_read_end=[x[1] for x in gl], next_read_start=[x[2] & 0xFFFFFFFF for x in gl], cur_genome_end=[x[3] for x in gl], next_genome_start=[x[4] & 0xFFFFFFFF for x in gl],
         read_off=[int(roff[x[0]]) for x in gl], chrom_off=np.zeros(len(gl), np.uint32))
g = {k: np.tile(np.asarray(v, np.uint32), 64) for k, v in g.items()}
rs4 = ctx.seq_upload(np.concatenate(reads2)); gs4 = ctx.seq_upload(contig2)
for it in range(2):
    t0 = time.time(); o = ctx.refine_linear_batch(rs4, gs4, g, 4, -3, -4, 15); report("a17 RefineByLinearAlignment", len(g["read_off"]), t0)
rg2 = np.random.default_rng(4)
cs = [roughgen.rough_list(rg2) for _ in range(512)] * (2 * R // 512)
l_off, lr_off = [0], [0]
cols = {k: [] for k in ["q", "t", "r_start", "r_end", "r_box", "r_strand", "r_freq", "r_chrom"]}
for q_, t_, rc_ in cs:
    l_off.append(l_off[-1] + len(q_)); lr_off.append(lr_off[-1] + len(rc_["start"]))
    cols["q"].append(q_); cols["t"].append(t_)
    for k_, kk in [("r_start", "start"), ("r_end", "end"), ("r_box", "box"), ("r_strand", "strand"), ("r_freq", "freq"), ("r_chrom", "chrom")]:
        cols[k_].append(rc_[kk])
rl = {k: np.concatenate(v) for k, v in cols.items()}; rl.update(l_off=np.array(l_off, np.uint64), lr_off=np.array(lr_off, np.uint64))
for it in range(2):
    t0 = time.time(); o = ctx.split_rough_batch(rl, 17, 1000, 2, 500); report("a8 SplitRoughClustersWithGaps (%d anchors)" % l_off[-1], len(cs), t0)
