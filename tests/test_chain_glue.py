"""a11: MergeChain and switchindex -- restatements pinned on the unmodified reference, kernels through the emulator (CPU) and the C ABI (GPU)."""
import numpy as np
import pytest

from oracle import pyoracle as po
import cgluegen

HAVE_REF = po.ref() is not None


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
def test_oracle_matches_reference():
    rng = np.random.default_rng(1)
    joined = shorter = 0
    for _ in range(1500):
        c = cgluegen.merge_case(rng)
        a = po.merge_chain(c["sp"], c["chrom"], c["strand"], c["box"], which="ref"); b = po.merge_chain(c["sp"], c["chrom"], c["strand"], c["box"])
        assert np.array_equal(a, b)
        joined += len(b) - int(b.sum())
        s = cgluegen.switch_case(rng)
        x = po.switchindex(s["ch"], s["link"], s["coarse"], s["cq"], which="ref"); y = po.switchindex(s["ch"], s["link"], s["coarse"], s["cq"])
        assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1])
        shorter += len(s["ch"]) - len(y[0])
    assert joined > 500 and shorter > 1000


def test_switchindex_known_answer():
    """{22, 125, 19, 125, 16, 17, 125, 57, 125} -> the 125 sandwich collapses (the reference's own example, Mapping_ultility.h:83), then 57 ... stay."""
    coarse = np.arange(200, dtype=np.int32)
    cq = np.stack([np.arange(200) * 100, np.arange(200) * 100 + 50], 1).astype(np.uint32)
    ch, link = po.switchindex([22, 125, 19, 125, 16, 17, 125, 57, 125], [1, 0, 1, 0, 1, 0, 1, 0], coarse, cq)
    assert ch.tolist() == [22, 125] and link.tolist() == [1]
    ch, link = po.switchindex([3, 3, 4, 4, 4, 5], [1, 0, 1, 1, 0], coarse, cq)
    assert ch.tolist() == [3, 4, 5] and link.tolist() == [0, 0]


def merge_batch(cases):
    sp, first, chrom, strand, box = [], [], [], [], []
    base = 0
    for c in cases:
        sp += [base + int(x) for x in c["sp"]]; first += [1] + [0] * (len(c["sp"]) - 1)
        chrom += list(c["chrom"]); strand += list(c["strand"]); box += c["box"].tolist(); base += len(c["chrom"])
    off = np.zeros(len(cases) + 1, np.uint64); off[1:] = np.cumsum([len(c["sp"]) for c in cases])
    return np.array(sp, np.int32), np.array(first, np.uint8), off, np.array(chrom, np.int32), np.array(strand, np.uint8), np.array(box, np.uint32)


def switch_batch(cases):
    ch, link, coarse, cq = [], [], [], []
    bs = bc = 0
    for c in cases:
        ch += [bs + int(x) for x in c["ch"]]; link += list(c["link"]) + [0]
        coarse += [bc + int(x) for x in c["coarse"]]; cq += c["cq"].tolist(); bs += len(c["coarse"]); bc += len(c["cq"])
    off = np.zeros(len(cases) + 1, np.uint64); off[1:] = np.cumsum([len(c["ch"]) for c in cases])
    return np.array(ch, np.int32), np.array(link, np.uint8), off, np.array(coarse, np.int32), np.array(cq, np.uint32), [sum(len(x["cq"]) for x in cases[:i]) for i in range(len(cases))]


def check_switch(res, off, cases, bases, which):
    ch, link, n_out, nl_out = res
    for k, c in enumerate(cases):
        x = po.switchindex(c["ch"], c["link"], c["coarse"], c["cq"], which=which)
        a = int(off[k])
        assert np.array_equal(ch[a:a + n_out[k]] - bases[k], x[0]) and np.array_equal(link[a:a + nl_out[k]], x[1]), k


def test_emu_chain_glue():
    import emu_lib
    rng = np.random.default_rng(5)
    mc = [cgluegen.merge_case(rng) for _ in range(300)]
    sp, first, off, chrom, strand, box = merge_batch(mc)
    head = emu_lib.merge_chain(sp, first, chrom, strand, box)
    exp = np.concatenate([po.merge_chain(c["sp"], c["chrom"], c["strand"], c["box"]) for c in mc])
    assert np.array_equal(head, exp)
    sc = [cgluegen.switch_case(rng) for _ in range(300)]
    ch, link, off, coarse, cq, bases = switch_batch(sc)
    check_switch(emu_lib.switchindex(ch, link, off, coarse, cq), off, sc, bases, "port")


@pytest.mark.gpu
def test_gpu_chain_glue():
    import lra_b200
    ctx = lra_b200.Context(0)
    rng = np.random.default_rng(9)
    which = "ref" if HAVE_REF else "port"
    mc = [cgluegen.merge_case(rng) for _ in range(4000)]
    sp, first, off, chrom, strand, box = merge_batch(mc)
    head = ctx.merge_chain_batch(sp, off, chrom, strand, box)
    exp = np.concatenate([po.merge_chain(c["sp"], c["chrom"], c["strand"], c["box"], which=which) for c in mc])
    assert np.array_equal(head, exp)
    sc = [cgluegen.switch_case(rng) for _ in range(4000)]
    ch, link, off, coarse, cq, bases = switch_batch(sc)
    check_switch(ctx.switchindex_batch(ch, link, off, coarse, cq), off, sc, bases, which)
    ctx.close()


# ---------------------------------------------------------------------------------------------- a17 SwitchToOriginalAnchors

def switch_orig_case(rng):
    n_cl = int(rng.integers(1, 6))
    run_off, start, end = [0], [], []
    for _ in range(n_cl):
        cuts = np.unique(np.concatenate([[0], rng.integers(1, 60, int(rng.integers(0, 8))), [60]]))
        start += cuts[:-1].tolist(); end += cuts[1:].tolist(); run_off.append(len(start))
    n = int(rng.integers(1, 30))
    cnum = rng.integers(0, n_cl, n)
    k = np.array([rng.integers(0, run_off[c + 1] - run_off[c]) for c in cnum])
    return dict(cnum=cnum, k=k, run_off=run_off, start=start, end=end, coarse=rng.integers(0, 50, n_cl))


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
def test_switch_to_original_oracle_matches_reference():
    rng = np.random.default_rng(4)
    for _ in range(300):
        c = switch_orig_case(rng)
        a = po.switch_to_original(**c, which="ref"); b = po.switch_to_original(**c)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.gpu
def test_gpu_switch_to_original():
    import lra_b200
    ctx = lra_b200.Context(0)
    rng = np.random.default_rng(6)
    cases = [switch_orig_case(rng) for _ in range(2000)]
    rs = np.concatenate([[c["start"][c["run_off"][a] + r] for a, r in zip(c["cnum"], c["k"])] for c in cases])
    re_ = np.concatenate([[c["end"][c["run_off"][a] + r] for a, r in zip(c["cnum"], c["k"])] for c in cases])
    co = np.concatenate([[c["coarse"][a] for a in c["cnum"]] for c in cases])
    off, chain, ci = ctx.switch_to_original_batch(rs, re_, co)
    which = "ref" if HAVE_REF else "port"
    exp = [po.switch_to_original(**c, which=which) for c in cases]
    assert np.array_equal(chain, np.concatenate([e[0] for e in exp])) and np.array_equal(ci, np.concatenate([e[1] for e in exp]))
    assert int(off[-1]) == len(chain)
    o2 = ctx.switch_to_original_batch([], [], [])
    assert len(o2[1]) == 0
    ctx.close()
