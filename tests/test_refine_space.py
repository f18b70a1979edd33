"""a14 (core, small spaces): RefineSpace's AffineOneGapAlign branch -- alignment with band 30, exact K-mers at multiples of K inside the blocks, identity,
coordinate shift -- and its minimizer branch for larger spaces (non-canonical minimizers, std::sort, banded CompareLists).  The restatement is pinned on the
unmodified reference; the kernels run through the emulator (CPU) and the C ABI (GPU)."""
import numpy as np
import pytest

from oracle import pyoracle as po

HAVE_REF = po.ref() is not None
B = np.frombuffer(b"ACGT", np.uint8)
SC = (4, -3, -4)
K = 17


def spaces(seed, n):
    rng = np.random.default_rng(seed)
    contig = B[rng.integers(0, 4, 80_000)].copy()
    reads, out = [], []
    for _ in range(n):
        L = int(rng.integers(1200, 4000)); s = int(rng.integers(100, len(contig) - L - 1200))
        read = contig[s:s + L].copy()
        mut = rng.random(L) < float(rng.choice([0.0, 0.02, 0.08]))
        read[mut] = B[rng.integers(0, 4, int(mut.sum()))]
        if rng.random() < 0.5:                       # a small deletion in the read
            cut = int(rng.integers(100, L - 100)); read = np.concatenate([read[:cut], read[cut + int(rng.integers(1, 9)):]])
        for _ in range(int(rng.integers(1, 5))):
            ql = int(rng.choice([0, 5, 40, 200, 600, 999])); ql = min(ql, len(read) - 1)
            qs = int(rng.integers(0, len(read) - ql))
            tl = max(0, min(999, ql + int(rng.integers(-25, 26))))
            lrts = int(rng.choice([0, 0, 20])); lrlen = int(rng.choice([0, 0, 35]))
            tl_core = max(0, tl - lrlen)
            ts = s + qs + int(rng.integers(-10, 11)) + lrts
            st = int(rng.random() < 0.4); cs = int(rng.random() < 0.8)
            out.append(dict(read=len(reads), qs=qs, qe=qs + ql, ts=ts, te=ts + tl_core, st=st, cs=cs, lrts=lrts, lrlength=lrlen))
        reads.append(read)
    return contig, reads, out


def expected(contig, reads, sp, which):
    return [po.refine_space(reads[x["read"]], contig, K, x["qs"], x["qe"], x["ts"], x["te"], x["st"], x["cs"], x["lrts"], x["lrlength"], *SC, which=which) for x in sp]


def same(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.float32(a[2]).tobytes() == np.float32(b[2]).tobytes()


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
def test_oracle_matches_reference():
    contig, reads, sp = spaces(1, 120)
    pairs = 0
    for x, a, b in zip(sp, expected(contig, reads, sp, "ref"), expected(contig, reads, sp, "port")):
        assert same(a, b), (x, a, b)
        pairs += len(b[0])
    assert pairs > 300


def test_emu_refine_space():
    import emu_lib
    contig, reads, sp = spaces(3, 50)
    roff = np.zeros(len(reads), np.int64); roff[1:] = np.cumsum([len(r) for r in reads[:-1]])
    pad = np.full(16, ord("A"), np.uint8)
    col = lambda k: [x[k] for x in sp]
    d = dict(qs=col("qs"), qe=col("qe"), ts=col("ts"), te=col("te"), lrts=col("lrts"), lrlength=col("lrlength"), read_off=[int(roff[x["read"]]) for x in sp],
             read_len=[len(reads[x["read"]]) for x in sp], chrom_off=np.zeros(len(sp), np.uint32), flip=[x["cs"] and x["st"] for x in sp])
    o = emu_lib.refine_space(np.concatenate(reads + [pad]), np.concatenate([contig, pad]), d, K, *SC)
    for i, e in enumerate(expected(contig, reads, sp, "port")):
        a = int(o["pair_off"][i]); n = int(o["n_pairs"][i])
        assert same((o["pq"][a:a + n], o["pt"][a:a + n], o["identity"][i]), e), i


def large_spaces(seed, n):
    """Spaces of 1000 bases or more on at least one axis (the minimizer branch), over a contig with repeats and N runs; K, W, the band and localMaxFreq vary."""
    rng = np.random.default_rng(seed)
    contig = B[rng.integers(0, 4, 200_000)].copy()
    for _ in range(20):
        a = int(rng.integers(0, 190_000)); b = int(rng.integers(0, 190_000)); L = int(rng.integers(50, 400)); contig[b:b + L] = contig[a:a + L]
    contig[50_000:50_030] = ord("N")
    reads, out = [], []
    for _ in range(n):
        L = int(rng.integers(1500, 9000)); s = int(rng.integers(100, len(contig) - L - 3000))
        read = contig[s:s + L].copy()
        mut = rng.random(L) < float(rng.choice([0.0, 0.03, 0.1])); read[mut] = B[rng.integers(0, 4, int(mut.sum()))]
        if rng.random() < 0.3:
            read[int(rng.integers(0, L - 20)):][:10] = ord("N")
        ql = min(int(rng.choice([300, 1000, 1500, 4000])), L - 1); qs = int(rng.integers(0, L - ql))
        tl = ql + int(rng.integers(-200, 201)); tl = max(tl, 1000 if ql < 1000 else 10)
        lrts = int(rng.choice([0, 25])); lrl = int(rng.choice([0, 40])); ts = s + qs + int(rng.integers(-30, 31)) + lrts
        out.append(dict(read=len(reads), qs=qs, qe=qs + ql, ts=ts, te=ts + tl - lrl, st=int(rng.random() < 0.4), cs=int(rng.random() < 0.8), lrts=lrts, lrlength=lrl,
                        diag=int(rng.choice([100, 300, 1000]))))
        reads.append(read)
    return contig, reads, out


def expected_large(contig, reads, sp, which, Kx, W, mf):
    return [po.refine_space(reads[x["read"]], contig, Kx, x["qs"], x["qe"], x["ts"], x["te"], x["st"], x["cs"], x["lrts"], x["lrlength"], *SC, which=which, W=W, diag=x["diag"],
                            local_max_freq=mf) for x in sp]


def space_dict(reads, sp):
    roff = np.zeros(len(reads), np.int64); roff[1:] = np.cumsum([len(r) for r in reads[:-1]])
    col = lambda k: [x[k] for x in sp]
    d = dict(qs=col("qs"), qe=col("qe"), ts=col("ts"), te=col("te"), lrts=col("lrts"), lrlength=col("lrlength"), read_off=[int(roff[x["read"]]) for x in sp],
             read_len=[len(reads[x["read"]]) for x in sp], chrom_off=np.zeros(len(sp), np.uint32), flip=[x["cs"] and x["st"] for x in sp])
    if sp and "diag" in sp[0]:
        d["diag"] = col("diag")
    return d


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
@pytest.mark.parametrize("Kx,W,mf", [(17, 10, 30), (15, 5, 2)])
def test_large_oracle_matches_reference(Kx, W, mf):
    contig, reads, sp = large_spaces(5, 60)
    pairs = 0
    for x, a, b in zip(sp, expected_large(contig, reads, sp, "ref", Kx, W, mf), expected_large(contig, reads, sp, "port", Kx, W, mf)):
        assert same(a, b), x
        pairs += len(b[0])
    assert pairs > 3000


@pytest.mark.parametrize("Kx,W,mf", [(17, 10, 30), (15, 5, 2)])
def test_emu_refine_space_large(Kx, W, mf):
    import emu_lib
    contig, reads, sp = large_spaces(6, 25)
    pad = np.full(16, ord("A"), np.uint8)
    o = emu_lib.refine_space_large(np.concatenate(reads + [pad]), np.concatenate([contig, pad]), space_dict(reads, sp), Kx, W, mf)
    for i, e in enumerate(expected_large(contig, reads, sp, "port", Kx, W, mf)):
        a = int(o["pair_off"][i]); n = int(o["n_pairs"][i])
        assert same((o["pq"][a:a + n], o["pt"][a:a + n], o["identity"][i]), e), i


@pytest.mark.gpu
@pytest.mark.parametrize("Kx,W,mf", [(17, 10, 30), (15, 5, 2)])
def test_gpu_refine_space_mixed(Kx, W, mf):
    """Small and large spaces in one call: both branches, one slot layout."""
    import lra_b200
    ctx = lra_b200.Context(0)
    contig, reads, sp = large_spaces(7, 300)
    rng = np.random.default_rng(8)
    for x in sp[::3]:                       # every third space shrinks below 1000 x 1000: the alignment branch
        x["qe"] = min(x["qs"] + int(rng.integers(0, 900)), len(reads[x["read"]])); x["te"] = x["ts"] + max(0, (x["qe"] - x["qs"]) + int(rng.integers(-20, 21)) - x["lrlength"])
    rs = ctx.seq_upload(np.concatenate(reads)); gs = ctx.seq_upload(contig)
    o = ctx.refine_space_batch(rs, gs, space_dict(reads, sp), Kx, *SC, W=W, local_max_freq=mf)
    nlarge = 0
    for i, e in enumerate(expected_large(contig, reads, sp, "ref" if HAVE_REF else "port", Kx, W, mf)):
        a = int(o["pair_off"][i]); n = int(o["n_pairs"][i])
        assert same((o["pq"][a:a + n], o["pt"][a:a + n], o["identity"][i]), e), (i, sp[i])
        nlarge += e[2] == np.float32(-1.0)
    assert 100 < nlarge < 300
    rs.free(); gs.free(); ctx.close()


@pytest.mark.gpu
def test_gpu_refine_space():
    import lra_b200
    ctx = lra_b200.Context(0)
    contig, reads, sp = spaces(2, 1200)
    roff = np.zeros(len(reads), np.int64); roff[1:] = np.cumsum([len(r) for r in reads[:-1]])
    rs = ctx.seq_upload(np.concatenate(reads)); gs = ctx.seq_upload(contig)
    col = lambda k: [x[k] for x in sp]
    d = dict(qs=col("qs"), qe=col("qe"), ts=col("ts"), te=col("te"), lrts=col("lrts"), lrlength=col("lrlength"), read_off=[int(roff[x["read"]]) for x in sp],
             read_len=[len(reads[x["read"]]) for x in sp], chrom_off=np.zeros(len(sp), np.uint32), flip=[x["cs"] and x["st"] for x in sp])
    o = ctx.refine_space_batch(rs, gs, d, K, *SC)
    for i, e in enumerate(expected(contig, reads, sp, "ref" if HAVE_REF else "port")):
        a = int(o["pair_off"][i]); n = int(o["n_pairs"][i])
        assert same((o["pq"][a:a + n], o["pt"][a:a + n], o["identity"][i]), e), i
    big = dict(d); big["qe"] = list(d["qe"]); big["qe"][0] = d["qs"][0] + 1000
    with pytest.raises(lra_b200.LraB200Error):        # a space of the minimizer branch without its refineSpaceDiag: refused, not guessed
        ctx.refine_space_batch(rs, gs, big, K, *SC)
    rs.free(); gs.free(); ctx.close()
