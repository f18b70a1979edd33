"""a12 LocalIndex::IndexSeq and a13 REFINEclusters on the GPU through the C ABI, against the reference (libref_lra.so travels to
the GPU box prebuilt) or, where it is absent, the pinned restatement."""
import numpy as np
import pytest

from oracle import pyoracle as po
import refinegen

pytestmark = pytest.mark.gpu
B = np.frombuffer(b"ACGT", np.uint8)
WHICH = "ref" if po.ref() is not None else "port"


def test_gpu_lindex_genome_and_reads():
    import lra_b200
    ctx = lra_b200.Context(0)
    rng = np.random.default_rng(7)
    contigs = []
    for L in (9, 14, 2048, 300000, 123457, 70001):
        s = B[rng.integers(0, 4, L)].copy()
        if L == 123457:
            s[5000:5100] = ord("N"); s[40000:52000] = np.tile(B[rng.integers(0, 4, 3)], 4000); s[90000:96000] = ord("A"); s[-1] = ord("N")
        if L == 70001:
            s[rng.random(L) < 0.01] = ord("N")
        contigs.append(s)
    lens = np.array([len(c) for c in contigs], np.uint32)
    start = np.zeros(len(contigs), np.uint64); start[1:] = np.cumsum(lens[:-1])
    arena = ctx.seq_upload(np.concatenate(contigs))
    for mf in (5, 15):
        img = ctx.lindex_build(arena, start, lens, max_freq=mf)
        wo, bd, mn = img.download()
        li = po.local_index(contigs, max_freq=mf, which=WHICH)
        assert (wo == li.seq_off).all() and (bd == li.bnd).all() and len(mn) == len(li.mins) and (mn == li.mins).all()
        # an uploaded image (the <ref>.gli path) is the same object
        up = ctx.lindex_upload(start, lens, 2048, wo, bd, mn)
        wo2, bd2, mn2 = up.download()
        assert (wo2 == wo).all() and (bd2 == bd).all() and (mn2 == mn).all()
        img.free(); up.free()
    names = [s["name"] for s in ctx.kernel_stats()]
    assert "lidx_window" in names
    arena.free(); ctx.close()


@pytest.mark.parametrize("seed,n_reads", [(31, 12), (32, 40)])
def test_gpu_refine_clusters(seed, n_reads):
    import lra_b200
    ctx = lra_b200.Context(0)
    case = refinegen.make_case(seed, n_reads=n_reads)
    pk = refinegen.pack_case(case)
    hdr = case["hdr"]
    g = ctx.seq_upload(pk["genome"][:-16]); rd = ctx.seq_upload(pk["arena"][:-16])
    rc = ctx.seq_revcomp(rd, pk["read_off"], pk["read_len"])
    gl = ctx.lindex_build(g, hdr[:-1], np.diff(hdr).astype(np.uint32))
    rf = ctx.lindex_build(rd, pk["read_off"], pk["read_len"]); rr = ctx.lindex_build(rc, pk["read_off"], pk["read_len"])
    o = ctx.refine_clusters_batch(gl, rf, rr, pk["cl"])
    exp = refinegen.expected(case, WHICH)
    assert o["n_anchors"] == sum(len(e["rq"]) for e in exp) and o["n_anchors"] > 1000
    refinegen.check_batch(o, pk["cl"], exp)
    # capacity protocol
    with pytest.raises(lra_b200.LraB200Error) as ei:
        ctx.refine_clusters_batch(gl, rf, rr, pk["cl"], anchor_cap=10)
    assert ei.value.code == lra_b200.capi.EOVERFLOW
    # an empty batch
    e = dict(pk["cl"]); e.update(m_q=np.zeros(0, np.uint32), m_t=np.zeros(0, np.uint32), m_off=np.zeros(1, np.uint64), box=np.zeros((0, 4), np.uint32),
                                 strand=np.zeros(0, np.uint8), read_id=np.zeros(0, np.uint32))
    assert ctx.refine_clusters_batch(gl, rf, rr, e)["n_anchors"] == 0
    for x in (gl, rf, rr, g, rd, rc):
        x.free()
    ctx.close()


@pytest.mark.parametrize("seed,n_reads,limit", [(41, 16, 1), (42, 30, 1), (41, 16, 0)])
def test_gpu_refine_splitchains(seed, n_reads, limit):
    import lra_b200
    ctx = lra_b200.Context(0)
    case = refinegen.make_case(seed, n_reads=n_reads)
    chains = refinegen.make_chains(case, seed)
    pk = refinegen.pack_chains(case, chains, limitrefine=limit)
    hdr = case["hdr"]
    g = ctx.seq_upload(pk["genome"][:-16]); rd = ctx.seq_upload(pk["arena"][:-16])
    rc = ctx.seq_revcomp(rd, pk["read_off"], pk["read_len"])
    gl = ctx.lindex_build(g, hdr[:-1], np.diff(hdr).astype(np.uint32))
    rf = ctx.lindex_build(rd, pk["read_off"], pk["read_len"]); rr = ctx.lindex_build(rc, pk["read_off"], pk["read_len"])
    # images rebuilt in place give the same result (the steady state of a batch loop)
    rf = ctx.lindex_build(rd, pk["read_off"], pk["read_len"], reuse=rf); rc = ctx.seq_revcomp(rd, pk["read_off"], pk["read_len"], reuse=rc)
    rr = ctx.lindex_build(rc, pk["read_off"], pk["read_len"], reuse=rr)
    o = ctx.refine_splitchains_batch(gl, rf, rr, pk["cl"])
    exp = refinegen.expected_chains(case, chains, WHICH, limitrefine=limit)
    assert o["n_anchors"] == sum(len(e["rq"]) for e in exp) and o["n_anchors"] > 1000
    refinegen.check_chain_batch(o, exp)
    for x in (gl, rf, rr, g, rd, rc):
        x.free()
    ctx.close()
