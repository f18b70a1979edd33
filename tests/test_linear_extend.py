"""a15 LinearExtend (GenomePairs overload) + Checkbp + DecideCoordinates + TrimOverlappedAnchors(vector<Cluster>&): the restatement pinned on the
unmodified reference, the kernels through the emulator (CPU) and through the C ABI (GPU).  Everything is integer work: bit-exact."""
import numpy as np
import pytest

from oracle import pyoracle as po
import lextgen

K = 17
# (seed, one part per group, unsorted input, skipsorting, trim): the two call sites of the low-accuracy pipeline + literal corner modes
MODES = [(1, True, False, 1, 0), (2, False, False, 0, 1), (3, False, True, 0, 1), (4, True, True, 1, 0), (5, False, False, 1, 1)]
HAVE_REF = po.ref() is not None


def expected(arena, items, skip, trim, which):
    return [po.linear_extend(read, arena, rd, K, skip, trim, which=which) for read, rd in items]


def check_batch(o, exp):
    e = 0
    for g_base, x in exp:
        G = len(x["e_off"]) - 1
        n = len(x["q"])
        assert (o["e_off"][g_base:g_base + G + 1].astype(np.int64) - e == x["e_off"]).all()
        for k in ("q", "t", "len"):
            assert np.array_equal(o[k][e:e + n], x[k]), k
        assert np.array_equal(o["box"][g_base:g_base + G], x["box"])
        e += n
    assert e == len(o["q"])


def with_bases(exp):
    out, g = [], 0
    for x in exp:
        out.append((g, x)); g += len(x["e_off"]) - 1
    return out


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
@pytest.mark.parametrize("mode", MODES)
def test_oracle_matches_reference(mode):
    seed, single, unsorted, skip, trim = mode
    arena, items = lextgen.reads(seed, 30, single=single, unsorted=unsorted)
    merged = 0
    for (read, rd), a, b in zip(items, expected(arena, items, skip, trim, "ref"), expected(arena, items, skip, trim, "port")):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        merged += len(rd["q"]) - len(b["q"])
    assert merged > 100          # the inputs do get merged / extended


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
def test_handmade_oracle_matches_reference():
    arena, read, rd, rd2 = lextgen.handmade(K)
    for r in (rd, rd2):
        for trim in (0, 1):
            a = po.linear_extend(read, arena, r, K, 0, trim, which="ref"); b = po.linear_extend(read, arena, r, K, 0, trim, which="port")
            for k in a:
                assert np.array_equal(a[k], b[k]), k


def test_handmade_known_answer():
    """The extension runs through the dinucleotide repeat on both diagonals, stops at the planted mismatch, and the first long anchor is trimmed by
    the 24-base genome overlap + 1."""
    arena, read, rd, rd2 = lextgen.handmade(K)
    a = po.linear_extend(read, arena, rd, K, 0, 0)
    assert list(zip(a["q"].tolist(), a["t"].tolist(), a["len"].tolist())) == [(0, 0, 84), (62, 60, 153), (230, 228, 116)]
    assert a["box"].tolist() == [[0, 346, 0, 344]]
    b = po.linear_extend(read, arena, rd, K, 0, 1)
    assert b["len"].tolist() == [59, 153, 116]
    c = po.linear_extend(read, arena, rd2, K, 0, 1)     # duplicates: which copy is trimmed is decided by the std::sort tie order
    assert sorted(c["len"].tolist()) == [59, 84, 116, 116, 153, 153]


@pytest.mark.parametrize("mode", MODES)
def test_emu_linear_extend(mode):
    import emu_lib
    seed, single, unsorted, skip, trim = mode
    arena, items = lextgen.reads(seed, 12, single=single, unsorted=unsorted)
    read_arena, ep = lextgen.to_batch(items)
    o = emu_lib.linear_extend(read_arena, arena, ep, K, skip, trim)
    check_batch(o, with_bases(expected(arena, items, skip, trim, "port")))


def test_emu_handmade_and_empty():
    import emu_lib
    arena, read, rd, rd2 = lextgen.handmade(K)
    items = [(read, rd), (read, rd2), (read, dict(rd, q=np.zeros(0, np.uint32), t=np.zeros(0, np.uint32), p_off=np.array([0, 0], np.int32)))]
    read_arena, ep = lextgen.to_batch(items)
    for trim in (0, 1):
        o = emu_lib.linear_extend(read_arena, arena, ep, K, 0, trim)
        check_batch(o, with_bases(expected(arena, items, 0, trim, "port")))


def test_emu_long_anchor_lists_reach_introsort():
    """More than 16 long anchors with ties in one group: the replayed introsort leaves the insertion-sort-only regime."""
    import emu_lib
    arena, read, rd, _ = lextgen.handmade(K)
    reps = 12
    n = len(rd["q"])
    big = dict(q=np.tile(rd["q"], reps), t=np.tile(rd["t"], reps), p_off=np.arange(reps + 1, dtype=np.int32) * n, p_strand=np.zeros(reps, np.uint8),
               chrom_off=np.zeros(reps, np.uint64), chrom_len=np.tile(rd["chrom_len"], reps), g_off=np.array([0, reps], np.int32))
    items = [(read, big)]
    read_arena, ep = lextgen.to_batch(items)
    exp = expected(arena, items, 0, 1, "port")
    assert (exp[0]["len"] >= 40).sum() > 16
    if HAVE_REF:
        r = po.linear_extend(read, arena, big, K, 0, 1, which="ref")
        assert np.array_equal(r["len"], exp[0]["len"]) and np.array_equal(r["q"], exp[0]["q"])
    check_batch(emu_lib.linear_extend(read_arena, arena, ep, K, 0, 1), with_bases(exp))


# ---------------------------------------------------------------------------------------------- the high-accuracy overload (chains)

def expected_chains(arena, items, trim, which, skiprep=1):
    out = []
    for read, cd, chains in items:
        for ch in chains:
            out.append(po.linear_extend_chain(read, arena, cd, ch, K, skiprep, trim, 100, which=which))
    return out


def check_chain_batch(o, exp):
    e = u = 0
    for x in exp:
        E = len(x["e_off"]) - 1
        n = len(x["q"])
        assert (o["e_off"][u:u + E + 1].astype(np.int64) - e == x["e_off"]).all()
        for k in ("q", "t", "len", "ovp"):
            assert np.array_equal(o[k][e:e + n], x[k]), k
        assert np.array_equal(o["box"][u:u + E], x["box"])
        assert int(o["overlap"][u:u + E].sum()) == x["overlap"]
        for j in range(E):            # md_head -> (start, end) runs of MergeMatchesSameDiag
            a, b = int(x["e_off"][j]), int(x["e_off"][j + 1])
            heads = np.flatnonzero(o["md_head"][e + a:e + b])
            ms, me = x["md_start"][x["md_off"][j]:x["md_off"][j + 1]], x["md_end"][x["md_off"][j]:x["md_off"][j + 1]]
            if b > a:
                assert np.array_equal(heads, ms) and np.array_equal(np.append(heads[1:], b - a), me)
            else:
                assert ms.tolist() == [0] and me.tolist() == [1]
        e += n; u += E
    assert e == len(o["q"])


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
@pytest.mark.parametrize("trim", [0, 1])
def test_chain_oracle_matches_reference(trim):
    arena, items = lextgen.chain_reads(21 + trim, 40)
    ovl = runs = total = 0
    for a, b in zip(expected_chains(arena, items, trim, "ref"), expected_chains(arena, items, trim, "port")):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        ovl += b["overlap"]; runs += len(b["md_start"]); total += len(b["q"])
    assert ovl > 20 and runs < total       # overlap anchors exist, and some anchors are merged into same-diagonal runs


def test_chain_skiprepetitive_off_has_no_overlap_anchors():
    arena, items = lextgen.chain_reads(23, 10)
    for x in expected_chains(arena, items, 1, "port", skiprep=0):
        assert x["overlap"] == 0 and not x["ovp"].any()


@pytest.mark.parametrize("trim", [0, 1])
def test_emu_linear_extend_chains(trim):
    import emu_lib
    arena, items = lextgen.chain_reads(31 + trim, 12)
    read_arena, cd = lextgen.chains_to_batch(items)
    o = emu_lib.linear_extend_chains(read_arena, arena, cd, K, 1, trim, 100)
    check_chain_batch(o, expected_chains(arena, items, trim, "port"))


def test_pairs_trim_mode_2():
    """TrimOverlappedAnchors(GenomePairs&, vector<int>&): anchors >= 50 only, forward order; the handmade first anchor (84 long) is still trimmed,
    and a 45-long anchor would not be."""
    import emu_lib
    arena, read, rd, rd2 = lextgen.handmade(K)
    b = po.linear_extend(read, arena, rd, K, 0, 2)
    assert b["len"].tolist() == [59, 153, 116]
    items = [(read, rd), (read, rd2)]
    read_arena, ep = lextgen.to_batch(items)
    check_batch(emu_lib.linear_extend(read_arena, arena, ep, K, 0, 2), with_bases(expected(arena, items, 0, 2, "port")))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", MODES)
def test_gpu_linear_extend(mode):
    import lra_b200
    seed, single, unsorted, skip, trim = mode
    ctx = lra_b200.Context(0)
    arena, items = lextgen.reads(seed + 10, 150, single=single, unsorted=unsorted)
    hm_arena, hm_read, rd, rd2 = lextgen.handmade(K)
    # the handmade contig is appended to the genome arena, its reads to the batch
    base = len(arena) - 16
    g2 = np.concatenate([arena[:-16], hm_arena])
    shift = lambda r: dict(r, chrom_off=r["chrom_off"] + np.uint64(base))
    items2 = items + [(hm_read, shift(rd)), (hm_read, shift(rd2))]
    read_arena, ep = lextgen.to_batch(items2)
    rs = ctx.seq_upload(read_arena[:-16]); gs = ctx.seq_upload(g2[:-16])
    o = ctx.linear_extend_batch(rs, gs, ep, K, skip, trim)
    which = "ref" if HAVE_REF else "port"
    exp = expected(arena, items, skip, trim, which) + [po.linear_extend(hm_read, hm_arena, rd, K, skip, trim, which=which),
                                                       po.linear_extend(hm_read, hm_arena, rd2, K, skip, trim, which=which)]
    check_batch(o, with_bases(exp))
    # no groups / no anchors
    e = ctx.linear_extend_batch(rs, gs, dict(g_off=np.zeros(1, np.uint64), p_off=np.zeros(1, np.uint64), p_strand=np.zeros(0, np.uint8), chrom_off=np.zeros(0, np.uint64),
                                             chrom_len=np.zeros(0, np.uint32), read_off=np.zeros(0, np.uint64), read_len=np.zeros(0, np.uint32),
                                             q=np.zeros(0, np.uint32), t=np.zeros(0, np.uint32)), K, skip, trim)
    assert len(e["q"]) == 0
    rs.free(); gs.free(); ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("trim", [0, 1])
def test_gpu_linear_extend_chains(trim):
    import lra_b200
    ctx = lra_b200.Context(0)
    arena, items = lextgen.chain_reads(41 + trim, 200)
    read_arena, cd = lextgen.chains_to_batch(items)
    rs = ctx.seq_upload(read_arena[:-16]); gs = ctx.seq_upload(arena[:-16])
    o = ctx.linear_extend_chains_batch(rs, gs, cd, K, 1, trim, 100)
    check_chain_batch(o, expected_chains(arena, items, trim, "ref" if HAVE_REF else "port"))
    rs.free(); gs.free(); ctx.close()


@pytest.mark.gpu
def test_gpu_pairs_trim_mode_2():
    import lra_b200
    ctx = lra_b200.Context(0)
    arena, items = lextgen.reads(77, 100, single=False)
    read_arena, ep = lextgen.to_batch(items)
    rs = ctx.seq_upload(read_arena[:-16]); gs = ctx.seq_upload(arena[:-16])
    check_batch(ctx.linear_extend_batch(rs, gs, ep, K, 0, 2), with_bases(expected(arena, items, 0, 2, "port")))
    rs.free(); gs.free(); ctx.close()
