"""a21 Alignment::CalculateStatistics: oracle pinned on the reference, kernel logic through the emulator (CPU) and the real
kernels through the C ABI (GPU).  CIGAR ops, the counters and the float NV value must be bit-identical."""
import os
import numpy as np
import pytest

from oracle import pyoracle as po
import synth
import workload


def segments(profile, n, seed=3, refine=True):
    genome = synth.gen_ref(3_000_000, 1, 5)[0][1]
    sb = workload.make_segments(profile, n, seed, len(genome), workload.host_genome_fetcher(genome))
    if refine:  # statistics run on the REFINED block lists in the reference (Map_highacc.h:718-721)
        outs = po.indel_refine_batch_port(sb, sb["t_arena_compact"], sb["t_base_compact"])
        sb = dict(sb)
        sb["blocks_in"] = np.concatenate(outs).astype(np.uint32)
        sb["blk_cnt"] = np.array([len(o) for o in outs], np.int32)
        off = np.zeros(n, np.uint64); off[1:] = np.cumsum(sb["blk_cnt"][:-1]); sb["blk_off"] = off
    return sb


def expect(sb, which):
    out = []
    for s in range(len(sb["blk_cnt"])):
        qb, tb = int(sb["q_base"][s]), int(sb["t_base_compact"][s])
        read = sb["q_arena"][qb:qb + sb["read_len"][s]].tobytes(); text = sb["t_arena_compact"][tb:tb + sb["contig_len"][s]].tobytes()
        blocks = sb["blocks_in"][int(sb["blk_off"][s]):int(sb["blk_off"][s]) + sb["blk_cnt"][s]]
        if which == "ref":
            st, v, c = po.calc_stats_ref(read, text, blocks)
        else:
            st, v, cig = po.calc_stats_port(read, text, 0, blocks); c = po.cigar_string(cig)
        out.append((st[:15], v, c))
    return out


@pytest.mark.skipif(po.ref() is None, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
@pytest.mark.parametrize("profile", ["ont", "ccs"])
def test_oracle_matches_reference(profile):
    sb = segments(profile, 12)
    for a, b in zip(expect(sb, "ref"), expect(sb, "port")):
        assert (a[0] == b[0]).all() and a[1] == b[1] and a[2] == b[2]


def check(o, sb, exp):
    for s, (st, v, c) in enumerate(exp):
        assert (o["stats"][s][:15] == st).all(), s
        assert o["value"][s] == v, (s, o["value"][s], v)
        a, b = int(o["cigar_off"][s]), int(o["cigar_off"][s + 1])
        assert po.cigar_string(o["cigar"][a:b]) == c, s


@pytest.mark.parametrize("thread_kernels", [0, 1])
def test_emu_stats(thread_kernels):
    import emu_lib
    sb = segments("ont", 4)
    o = emu_lib.calc_stats(sb, sb["t_arena_compact"], sb["t_base_compact"], po.log_lut(), thread_kernels)
    check(o, sb, expect(sb, "port"))


def _handmade():
    """Runs that span many lanes of the warp kernel, long gaps (the float replay of NV), zero-length blocks, a one-block segment,
    a segment shorter than a warp, an empty segment."""
    rng = np.random.default_rng(77)
    B = np.frombuffer(b"ACGT", np.uint8)
    reads, texts, blocks = [], [], []
    # 1: a perfect 3000-base match (one '=' run over all 32 lanes), then a 700-base deletion, then 40 matches with one mismatch
    t = B[rng.integers(0, 4, 3740)].copy(); r = np.concatenate([t[:3000], t[3700:3740]]); r[3020] = B[(np.searchsorted(B, r[3020]) + 1) & 3]
    reads.append(r); texts.append(t); blocks.append(np.array([[0, 0, 3000], [3000, 3700, 40]], np.uint32))
    # 2: insertion of 25 (log table), deletion of 12000 and of 100002 (the two constant penalties), zero-length blocks in between
    t = B[rng.integers(0, 4, 112200)].copy(); r = np.concatenate([t[:50], B[rng.integers(0, 4, 25)], t[50:60], t[12060:12080], t[112082:112100]])
    reads.append(r); texts.append(t)
    blocks.append(np.array([[0, 0, 50], [75, 50, 10], [85, 60, 0], [85, 12060, 20], [105, 12080, 0], [105, 112082, 18]], np.uint32))
    # 3: one block; 4: five columns; 5: no blocks
    t = B[rng.integers(0, 4, 70)].copy(); reads.append(t[10:60].copy()); texts.append(t); blocks.append(np.array([[0, 10, 50]], np.uint32))
    t = B[rng.integers(0, 4, 9)].copy(); reads.append(t[:5].copy()); texts.append(t); blocks.append(np.array([[0, 0, 2], [2, 3, 2]], np.uint32))
    reads.append(t[:5].copy()); texts.append(t); blocks.append(np.zeros((0, 3), np.uint32))
    # 6: both gaps positive (common columns), mismatches inside them
    t = B[rng.integers(0, 4, 400)].copy(); r = t[:380].copy(); r[100:104] = B[(np.searchsorted(B, r[100:104]) + 2) & 3]
    reads.append(r); texts.append(t); blocks.append(np.array([[0, 0, 100], [104, 104, 96], [203, 220, 150]], np.uint32))
    n = len(reads)
    q_base = np.zeros(n, np.uint32); t_base = np.zeros(n, np.uint32)
    q_base[1:] = np.cumsum([len(x) for x in reads[:-1]]); t_base[1:] = np.cumsum([len(x) for x in texts[:-1]])
    cnt = np.array([len(b) for b in blocks], np.int32); off = np.zeros(n, np.uint64); off[1:] = np.cumsum(cnt[:-1])
    pad = np.full(16, ord("A"), np.uint8)
    return dict(q_arena=np.concatenate(reads + [pad]), t_arena_compact=np.concatenate(texts + [pad]), blocks_in=np.concatenate(blocks), blk_off=off, blk_cnt=cnt,
                q_base=q_base, t_base_compact=t_base, read_len=np.array([len(x) for x in reads], np.int32), contig_len=np.array([len(x) for x in texts], np.int32))


@pytest.mark.skipif(po.ref() is None, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
def test_handmade_oracle_matches_reference():
    sb = _handmade()
    for a, b in zip(expect(sb, "ref"), expect(sb, "port")):
        assert (a[0] == b[0]).all() and a[1] == b[1] and a[2] == b[2]


@pytest.mark.parametrize("thread_kernels", [0, 1])
def test_emu_stats_handmade(thread_kernels):
    import emu_lib
    sb = _handmade()
    o = emu_lib.calc_stats(sb, sb["t_arena_compact"], sb["t_base_compact"], po.log_lut(), thread_kernels)
    exp = expect(sb, "port")
    assert "700D" in exp[0][2] and "25I" in exp[1][2] and "12000D" in exp[1][2] and "100002D" in exp[1][2]
    check(o, sb, exp)


def test_long_gaps_use_the_log_table():
    read = b"ACGT" * 30; text = b"ACGT" * 10 + b"G" * 300 + b"ACGT" * 20
    st, v, cig = po.calc_stats_port(read, text, 0, np.array([[0, 0, 40], [40, 340, 80]], np.uint32))
    assert po.cigar_string(cig) == "40=300D80=" and abs(float(v) - 101.928925) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("profile,n", [("ont", 200), ("ccs", 1500)])
def test_gpu_stats(profile, n):
    import lra_b200
    ctx = lra_b200.Context(0)
    sb = segments(profile, n, seed=9)
    q = ctx.seq_upload(sb["q_arena"][:-16]); t = ctx.seq_upload(sb["t_arena_compact"][:-16])
    g = dict(sb); g["t_base"] = sb["t_base_compact"]
    o = ctx.calc_stats_batch(q, t, g, po.log_lut())
    check(o, sb, expect(sb, "ref" if po.ref() is not None else "port"))
    # a long deletion, so that the float log-table branch is exercised on the device
    read = b"ACGT" * 30; text = b"ACGT" * 10 + b"G" * 300 + b"ACGT" * 20
    q2 = ctx.seq_upload(read); t2 = ctx.seq_upload(text)
    g2 = dict(blocks_in=np.array([[0, 0, 40], [40, 340, 80]], np.uint32), blk_off=np.zeros(1, np.uint64), blk_cnt=np.array([2], np.int32),
              q_base=np.zeros(1, np.uint32), t_base=np.zeros(1, np.uint32), read_len=np.array([len(read)], np.int32))
    o2 = ctx.calc_stats_batch(q2, t2, g2, po.log_lut())
    st, v, cig = po.calc_stats_port(read, text, 0, g2["blocks_in"])
    assert o2["value"][0] == v and po.cigar_string(o2["cigar"][:int(o2["cigar_off"][1])]) == "40=300D80="
    sb3 = _handmade()
    q3 = ctx.seq_upload(sb3["q_arena"][:-16]); t3 = ctx.seq_upload(sb3["t_arena_compact"][:-16])
    g3 = dict(sb3); g3["t_base"] = sb3["t_base_compact"]
    check(ctx.calc_stats_batch(q3, t3, g3, po.log_lut()), sb3, expect(sb3, "ref" if po.ref() is not None else "port"))
    for x in (q, t, q2, t2, q3, t3):
        x.free()
    ctx.close()
