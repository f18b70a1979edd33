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


def test_emu_stats():
    import emu_lib
    sb = segments("ont", 4)
    o = emu_lib.calc_stats(sb, sb["t_arena_compact"], sb["t_base_compact"], po.log_lut())
    check(o, sb, expect(sb, "port"))


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
    for x in (q, t, q2, t2):
        x.free()
    ctx.close()
