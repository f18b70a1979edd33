"""Synthetic AffineOneGapAlign job streams for bench.py and the full-size parity tests.

Job SHAPES (qLen, tLen, k) are drawn from the tables captured from the reference on synthetic reads of the named
profile (tests/golden/aog_shapes_<profile>.npy, made by tools/make_golden.py); job CONTENT is synthetic: the target
window is a random window of the synthetic genome and the query is that sequence with i.i.d. errors at the profile's
rate (sub:ins:del = 1:1:1), cut to the drawn query length.
"""
import json
import os
import numpy as np

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
PROFILE_ERR = {"ccs": 0.01, "ont": 0.08, "clr": 0.12}
PROFILE_SCORING = {"ccs": (4, -3, -4), "ont": (4, -1, -2), "clr": (4, -1, -2)}  # localMatch/Mismatch/Indel presets


def meta():
    return json.load(open(os.path.join(GOLD, "aog_shapes_meta.json")))


def draw_shapes(profile, n_jobs, rng):
    tab = np.load(os.path.join(GOLD, "aog_shapes_%s.npy" % profile))
    s = tab[rng.integers(0, len(tab), size=n_jobs)]
    return s[:, 0].astype(np.int32), s[:, 1].astype(np.int32), s[:, 2].astype(np.int32)


def make_jobs(profile, n_jobs, seed, genome_len, fetch_windows):
    """fetch_windows(starts[int64], lengths[int64]) -> (concatenated uint8 ASCII, offsets) of genome windows.
    Returns dict with q_arena (uint8 ASCII, padded), q_off, t_off (GLOBAL genome positions), q_len, t_len, k and
    t_arena_compact/t_off_compact (the same target windows concatenated, for CPU arms that have no genome)."""
    rng = np.random.default_rng(seed)
    ql, tl, k = draw_shapes(profile, n_jobs, rng)
    src_len = (np.maximum(ql, tl).astype(np.int64) * 13) // 10 + 48
    t_pos = rng.integers(0, genome_len - int(src_len.max()) - 1, size=n_jobs, dtype=np.int64)
    src, src_off = fetch_windows(t_pos, src_len)
    # mutate the whole concatenation at once, then find where each job's segment starts in the mutated text
    err = PROFILE_ERR[profile]
    n = len(src)
    r = rng.random(n)
    kind = np.zeros(n, np.uint8)
    kind[r < err] = 1; kind[r < 2 * err / 3] = 2; kind[r < err / 3] = 3
    cnt = np.ones(n, np.int64); cnt[kind == 3] = 0; cnt[kind == 2] = 2
    pos = np.cumsum(cnt) - cnt
    total = int(pos[-1] + cnt[-1])
    code = (np.searchsorted(synth.BASES, src) & 3).astype(np.uint8)
    sub = synth.BASES[(code + rng.integers(1, 4, size=n, dtype=np.uint8)) & 3]
    base = np.where(kind == 1, sub, src)
    mut = np.empty(total + 16, np.uint8); mut[total:] = ord("A")
    keep = kind != 3
    mut[pos[keep]] = base[keep]
    ins = kind == 2
    mut[pos[ins] + 1] = synth.BASES[rng.integers(0, 4, size=int(ins.sum()), dtype=np.uint8)]
    seg_start = pos[src_off]
    seg_end = np.append(seg_start[1:], total)
    avail = seg_end - seg_start
    ql = np.minimum(ql, np.maximum(avail, 0)).astype(np.int32)
    # compact the query arena: job j = mut[seg_start[j] : seg_start[j]+ql[j]]
    q_off = np.zeros(n_jobs, np.int64); np.cumsum(ql[:-1], out=q_off[1:])
    qtot = int(ql.sum())
    gather = np.repeat(seg_start - q_off, ql) + np.arange(qtot, dtype=np.int64)
    q_arena = np.empty(qtot + 16, np.uint8); q_arena[qtot:] = ord("A")
    q_arena[:qtot] = mut[gather]
    # compact target arena
    t_off_c = np.zeros(n_jobs, np.int64); np.cumsum(tl[:-1], out=t_off_c[1:])
    ttot = int(tl.sum())
    gather_t = np.repeat(src_off - t_off_c, tl) + np.arange(ttot, dtype=np.int64)
    t_arena = np.empty(ttot + 16, np.uint8); t_arena[ttot:] = ord("A")
    t_arena[:ttot] = src[gather_t]
    return dict(q_arena=q_arena, q_off=q_off.astype(np.uint32), t_off=t_pos.astype(np.uint32), q_len=ql, t_len=tl, k=k,
                t_arena_compact=t_arena, t_off_compact=t_off_c.astype(np.uint32), scoring=PROFILE_SCORING[profile])


def host_genome_fetcher(genome):
    """fetch_windows over a host uint8 genome array."""
    def fetch(starts, lengths):
        off = np.zeros(len(starts), np.int64); np.cumsum(lengths[:-1], out=off[1:])
        tot = int(lengths.sum())
        idx = np.repeat(starts - off, lengths) + np.arange(tot, dtype=np.int64)
        return genome[idx], off
    return fetch


def torch_genome_fetcher(genome_dev):
    """fetch_windows over a device-resident torch uint8 genome tensor (gathers on the GPU, returns host arrays)."""
    import torch

    def fetch(starts, lengths):
        off = np.zeros(len(starts), np.int64); np.cumsum(lengths[:-1], out=off[1:])
        tot = int(lengths.sum())
        dev = genome_dev.device
        s = torch.from_numpy(starts - off).to(dev)
        idx = torch.repeat_interleave(s, torch.from_numpy(lengths).to(dev)) + torch.arange(tot, device=dev)
        return genome_dev[idx].cpu().numpy(), off
    return fetch


def check_blocks_property(jobs, res, m, mm, indel):
    """Size-independent self-check of a batch result for jobs in the one-sided mode: blocks are ordered, inside the
    windows and non-overlapping, and the score equals the score recomputed from the blocks (matches/mismatches inside
    blocks + indel per gap base along the path from the origin to the traceback start cell).  Returns #jobs checked."""
    ql = jobs["q_len"].astype(np.int64); tl = jobs["t_len"].astype(np.int64); kin = jobs["k"].astype(np.int64)
    diag = np.maximum(1, np.minimum(ql, tl)); k0 = np.minimum(diag, kin)
    one = diag + 2 * k0 >= np.maximum(ql, tl)
    k = 2 * k0
    qB = np.minimum(diag + k, ql + 1); tB = np.minimum(diag + k, tl + 1)
    nb = res["n_blocks"].astype(np.int64); off = res["block_off"].astype(np.int64)
    blk = res["blocks"]
    jid = np.repeat(np.arange(len(nb)), nb)
    bi = np.repeat(off, nb) + (np.arange(int(nb.sum())) - np.repeat(np.cumsum(nb) - nb, nb))
    b = blk[bi].astype(np.int64)
    qp, tp, ln = b[:, 0], b[:, 1], b[:, 2]
    assert (ln > 0).all()
    assert (qp + ln <= ql[jid]).all() and (tp + ln <= tl[jid]).all()
    same = jid[1:] == jid[:-1]
    assert (qp[1:][same] >= (qp + ln)[:-1][same]).all() and (tp[1:][same] >= (tp + ln)[:-1][same]).all()
    # recompute the score of one-sided jobs
    qa = jobs["q_arena"]; ta = jobs["t_arena_compact"]
    tot = int(ln.sum())
    boff = np.cumsum(ln) - ln
    within = np.arange(tot, dtype=np.int64) - np.repeat(boff, ln)
    qi = np.repeat(jobs["q_off"].astype(np.int64)[jid] + qp, ln) + within
    ti = np.repeat(jobs["t_off_compact"].astype(np.int64)[jid] + tp, ln) + within
    eq = qa[qi] == ta[ti]
    per_block = np.add.reduceat(np.where(eq, m, mm), boff) if tot else np.zeros(0, np.int64)
    sc = np.zeros(len(nb), np.int64); np.add.at(sc, jid, per_block)
    cov = np.zeros(len(nb), np.int64); np.add.at(cov, jid, ln)
    expect = sc + indel * ((qB - 1 - cov) + (tB - 1 - cov))
    assert (res["score"].astype(np.int64)[one] == expect[one]).all()
    return int(one.sum())


# ---------------------------------------------------------------------------------------------------------------------
# a19 IndelRefineAlignment: synthetic SEGMENTS.  One segment per read: the target is a random genome window of the read's
# length, the read is that window with i.i.d. errors, and the segment's input blocks are the maximal gap-free runs of the
# true alignment (what the upstream chain refinement hands to IndelRefineAlignment: blocks separated by small indels).
PROFILE_REFINE_BAND = {"ccs": 7, "ont": 7, "clr": 20}
PROFILE_END_ALIGN = {"ccs": 1, "ont": 0, "clr": 0}


def make_segments(profile, n_reads, seed, genome_len, fetch_windows, max_len=None):
    rng = np.random.default_rng(seed)
    lens = synth.read_lengths({"ccs": "ccs10k", "ont": "ont", "clr": "clr"}[profile], n_reads, rng).astype(np.int64)
    if max_len:
        lens = np.minimum(lens, max_len)
    t_pos = rng.integers(0, genome_len - int(lens.max()) - 1, size=n_reads, dtype=np.int64)
    src, src_off = fetch_windows(t_pos, lens)
    n = len(src)
    err = PROFILE_ERR[profile]
    r = rng.random(n)
    kind = np.zeros(n, np.uint8)
    kind[r < err] = 1; kind[r < 2 * err / 3] = 2; kind[r < err / 3] = 3
    first = np.zeros(n, bool); first[src_off] = True
    last = np.zeros(n, bool); last[src_off + lens - 1] = True
    kind[first | last] = 0                      # segments start and end on an aligned base
    cnt = np.ones(n, np.int64); cnt[kind == 3] = 0; cnt[kind == 2] = 2
    pos = np.cumsum(cnt) - cnt
    total = int(pos[-1] + cnt[-1])
    code = (np.searchsorted(synth.BASES, src) & 3).astype(np.uint8)
    sub = synth.BASES[(code + rng.integers(1, 4, size=n, dtype=np.uint8)) & 3]
    q_arena = np.empty(total + 16, np.uint8); q_arena[total:] = ord("A")
    keep = kind != 3
    q_arena[pos[keep]] = np.where(kind == 1, sub, src)[keep]
    ins = kind == 2
    q_arena[pos[ins] + 1] = synth.BASES[rng.integers(0, 4, size=int(ins.sum()), dtype=np.uint8)]
    q_base = pos[src_off]
    read_len = np.append(q_base[1:], total) - q_base
    # blocks: maximal runs of emitted bases not interrupted by an insertion or a deletion
    prev_kind = np.concatenate([[0], kind[:-1]])
    prev_keep = np.concatenate([[False], keep[:-1]])
    start = keep & (first | ~prev_keep | (prev_kind == 2))
    run_id = np.cumsum(start) - 1
    blen = np.bincount(run_id[keep], minlength=int(start.sum()))
    sidx = np.flatnonzero(start)
    seg_of = np.searchsorted(src_off, sidx, side="right") - 1
    blocks = np.stack([pos[sidx] - q_base[seg_of], sidx - src_off[seg_of], blen], 1).astype(np.uint32)
    blk_cnt = np.bincount(seg_of, minlength=n_reads).astype(np.int32)
    blk_off = np.zeros(n_reads, np.uint64); blk_off[1:] = np.cumsum(blk_cnt[:-1])
    t_arena = np.empty(n + 16, np.uint8); t_arena[:n] = src; t_arena[n:] = ord("A")
    m, mm, indel = PROFILE_SCORING[profile]
    return dict(q_arena=q_arena, blocks_in=blocks, blk_off=blk_off, blk_cnt=blk_cnt, q_base=q_base.astype(np.uint32),
                t_base=t_pos.astype(np.uint32), read_len=read_len.astype(np.int32), contig_len=lens.astype(np.int32),
                t_arena_compact=t_arena, t_base_compact=src_off.astype(np.uint32), k=PROFILE_REFINE_BAND[profile], match=m, mismatch=mm,
                indel=indel, end_align=PROFILE_END_ALIGN[profile], bases=int(lens.sum()))


def make_clusters(sg, global_k=17, compact=False, contig=125_000_000, genome_len=None):
    """The cluster batch REFINEclusters / Refine_splitchain would see for the reads of make_segments(): one cluster per read whose anchors
    are the exact stretches of at least global_k bases of the read's alignment (what minimizer seeding + clustering hand over on a
    repeat-free genome: ONT ~380 anchors per 20 kb read).  compact=False: genome positions in the full genome
    (contigs of `contig` bases); compact=True: positions in sg["t_arena_compact"], every read's window its own contig (the CPU arm)."""
    blocks = sg["blocks_in"]; n = len(sg["blk_cnt"])
    seg_of = np.repeat(np.arange(n), sg["blk_cnt"])
    keep = blocks[:, 2] >= global_k
    first = np.zeros(len(blocks), bool); first[sg["blk_off"].astype(np.int64)] = True        # every read keeps at least its first block
    keep |= first
    t_base = (sg["t_base_compact"] if compact else sg["t_base"]).astype(np.int64)
    m_q = blocks[keep, 0].astype(np.uint32)
    m_t = (blocks[keep, 1].astype(np.int64) + t_base[seg_of[keep]]).astype(np.uint32)
    m_len = blocks[keep, 2].astype(np.uint32)
    cnt = np.bincount(seg_of[keep], minlength=n)
    m_off = np.zeros(n + 1, np.uint64); m_off[1:] = np.cumsum(cnt)
    lo = m_off[:-1].astype(np.int64); hi = m_off[1:].astype(np.int64) - 1
    box = np.stack([m_q[lo], m_q[hi] + global_k, m_t[lo], m_t[hi] + global_k], 1).astype(np.uint32)
    if compact:      # one contig per read window, plus the 16 padding bases as a last dummy contig (no chain touches the genome's very last window)
        hdr = np.append(sg["t_base_compact"].astype(np.uint64), [np.uint64(len(sg["t_arena_compact"]) - 16), np.uint64(len(sg["t_arena_compact"]))])
    else:
        hdr = np.append(np.arange(0, genome_len, contig, dtype=np.uint64), np.uint64(genome_len))
    # the same anchors as a split chain of the low-accuracy pipeline (Refine_splitchain): anchor lengths = the exact stretches, boundaries with
    # the lengths, the contig each chain lies on (chains across a contig boundary are dropped by the reference before this stage: empty here)
    m_len[hi] = np.maximum(m_len[hi], 2) - 1          # keep TEnd strictly inside the contig (Header::Find(TEnd) of a chain ending exactly at a contig
    last_len = m_len[hi]                                # end names the NEXT contig, and the reference drops such chains before this stage)
    cbox = np.stack([m_q[lo], m_q[hi] + last_len, m_t[lo], m_t[hi] + last_len], 1).astype(np.uint32)
    chrom = (np.searchsorted(hdr, cbox[:, 2], side="right") - 1).astype(np.int32)
    ok = (np.searchsorted(hdr, cbox[:, 3], side="left") - 1) == chrom
    return dict(m_q=m_q, m_t=m_t, m_off=m_off, box=box, strand=np.zeros(n, np.uint8), read_id=np.arange(n, dtype=np.uint32), hdr_pos=hdr,
                global_k=global_k, small_k=10, window=100, local_max_freq=15,
                m_len=m_len, m_strand=np.zeros(len(m_q), np.uint8), chain_box=cbox, chrom=np.where(ok, chrom, 0).astype(np.int32), chain_ok=ok, limitrefine=1)
