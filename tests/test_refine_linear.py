"""a17 (leaf): RefineByLinearAlignment = SetMatchAndGaps + AlignSubstrings + RefineSubstrings around a18, batched over gaps.  The restatement is pinned on
the unmodified reference; the GPU path (job kernel, a18 kernels, shift kernel) goes through the C ABI."""
import numpy as np
import pytest

from oracle import pyoracle as po

HAVE_REF = po.ref() is not None
B = np.frombuffer(b"ACGT", np.uint8)
SC = (4, -3, -4, 15)          # localMatch, localMismatch, localIndel, localBand of the reference's defaults... the values only need to agree on both sides


def gaps(seed, n):
    """Reads = mutated copies of contig windows; gaps of 0..400 bases with length differences, empty gaps, touching anchors (qe == qs), reversed coordinates
    (next before cur: Matched() <= 0)."""
    rng = np.random.default_rng(seed)
    contig = B[rng.integers(0, 4, 60_000)].copy()
    out, reads = [], []
    roff = 0
    for _ in range(n):
        L = int(rng.integers(500, 3000)); s = int(rng.integers(0, len(contig) - L - 500))
        read = contig[s:s + L].copy()
        mut = rng.random(L) < 0.05
        read[mut] = B[rng.integers(0, 4, int(mut.sum()))]
        for _ in range(int(rng.integers(1, 6))):
            qs = int(rng.integers(0, L - 450))
            ql = int(rng.choice([0, 1, 5, 30, 120, 400]))
            d = int(rng.choice([0, 0, 1, -1, 3, -7, 20, -20]))
            tl = max(0, ql + d)
            ts = s + qs + int(rng.integers(-3, 4))
            qe, te = qs + ql, ts + tl
            x = rng.random()           # Matched() <= 0: one side ends exactly one before it starts, or both are reversed (a lone reversed side with m > 0
            if x < 0.05:               # makes the reference build a string of negative length)
                qe = qs - 1
            elif x < 0.10:
                te = ts - 1
            elif x < 0.14:
                qe = qs - int(rng.integers(1, 4)); te = ts - int(rng.integers(1, 4))
            if ts < 4:
                continue
            out.append((len(reads), qs, qe, ts, te))
        reads.append(read)
    return contig, reads, out


def expected(contig, reads, gl, which):
    return [po.refine_linear(reads[r], contig, qs, qe, ts, te, *SC, which=which) for r, qs, qe, ts, te in gl]


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
def test_oracle_matches_reference():
    contig, reads, gl = gaps(1, 150)
    nb = empty = 0
    for a, b in zip(expected(contig, reads, gl, "ref"), expected(contig, reads, gl, "port")):
        assert np.array_equal(a, b)
        nb += len(b); empty += len(b) == 0
    assert nb > 500 and empty > 10


def test_emu_refine_linear():
    import emu_lib
    contig, reads, gl = gaps(3, 60)
    roff = np.zeros(len(reads), np.int64); roff[1:] = np.cumsum([len(r) for r in reads[:-1]])
    pad = np.full(16, ord("A"), np.uint8)
    g = dict(cur_read_end=[x[1] for x in gl], next_read_start=[x[2] & 0xFFFFFFFF for x in gl], cur_genome_end=[x[3] for x in gl], next_genome_start=[x[4] & 0xFFFFFFFF for x in gl],
             read_off=[int(roff[x[0]]) for x in gl], chrom_off=np.zeros(len(gl), np.uint32))
    o = emu_lib.refine_linear(np.concatenate(reads + [pad]), np.concatenate([contig, pad]), g, *SC)
    for i, e in enumerate(expected(contig, reads, gl, "port")):
        a = int(o["block_off"][i])
        assert o["n_blocks"][i] == len(e) and np.array_equal(o["blocks"][a:a + len(e)], e), i


@pytest.mark.gpu
def test_gpu_refine_linear():
    import lra_b200
    ctx = lra_b200.Context(0)
    contig, reads, gl = gaps(2, 1500)
    roff = np.zeros(len(reads), np.int64); roff[1:] = np.cumsum([len(r) for r in reads[:-1]])
    rs = ctx.seq_upload(np.concatenate(reads)); gs = ctx.seq_upload(contig)
    g = dict(cur_read_end=[x[1] for x in gl], next_read_start=[x[2] & 0xFFFFFFFF for x in gl], cur_genome_end=[x[3] for x in gl], next_genome_start=[x[4] & 0xFFFFFFFF for x in gl],
             read_off=[int(roff[x[0]]) for x in gl], chrom_off=np.zeros(len(gl), np.uint32))
    o = ctx.refine_linear_batch(rs, gs, g, *SC)
    exp = expected(contig, reads, gl, "ref" if HAVE_REF else "port")
    for i, e in enumerate(exp):
        a = int(o["block_off"][i])
        assert o["n_blocks"][i] == len(e), i
        assert np.array_equal(o["blocks"][a:a + len(e)], e), i
    assert ctx.refine_linear_batch(rs, gs, {k: [] for k in g}, *SC)["n_blocks_total"] == 0
    rs.free(); gs.free(); ctx.close()
