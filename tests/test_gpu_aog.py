"""GPU parity tests for a18 AffineOneGapAlign: the CUDA path, called through the C ABI (include/lra_b200.h), against the
oracle (C restatement, golden vectors captured from the reference, and the reference header itself when its prebuilt
wrapper library travelled to the box).  Bit-exact: scores, block counts, every block triple."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po
import jobgen
import synth
import workload

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ctx():
    import lra_b200
    c = lra_b200.Context(0)
    yield c
    c.close()


def run_and_compare(ctx, batch, m, mm, indel):
    qa, ta, qo, to, ql, tl, k = batch
    es, enb, eoff, eblk, st = po.aog_batch_port(qa, ta, qo, to, ql, tl, k, m, mm, indel)
    assert (st == 0).all()
    q = ctx.seq_upload(qa[:-16]); t = ctx.seq_upload(ta[:-16])
    try:
        r = ctx.aog_batch(q, t, qo, to, ql, tl, k, m, mm, indel)
    finally:
        q.free(); t.free()
    bad = np.flatnonzero((r["score"] != es) | (r["n_blocks"] != enb))
    assert len(bad) == 0, (bad[:5], r["score"][bad[:5]], es[bad[:5]], ql[bad[:5]], tl[bad[:5]], k[bad[:5]])
    assert r["n_blocks_total"] == int(enb.sum())
    for j in range(len(qo)):
        a = r["blocks"][int(r["block_off"][j]):int(r["block_off"][j]) + enb[j]]
        e = eblk[eoff[j]:eoff[j] + enb[j]]
        assert (a == e).all(), (j, ql[j], tl[j], k[j])
    return r


def test_seq_pack_matches_alphabet(ctx):
    rng = np.random.default_rng(0)
    for n in [1, 15, 16, 31, 32, 33, 1000, 100003]:
        s = np.frombuffer(b"ACGTacgtNnXR\x00\x01\x02\x03\x07-", dtype=np.uint8)[rng.integers(0, 18, n)].copy()
        a = ctx.seq_upload(s)
        b2, nm = a.download()
        a.free()
        lut = np.full(256, 4, np.uint8)
        for ch, v in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
            lut[ch] = v
        for v in range(8):
            lut[v] = v & 3
        c = lut[s]
        two = np.where(c == 4, 0, c).astype(np.uint32)
        pad = (-n) % 16
        two_p = np.concatenate([two, np.zeros(pad, np.uint32)]).reshape(-1, 16)
        exp_b2 = (two_p << (2 * np.arange(16, dtype=np.uint32))).sum(1).astype(np.uint32)
        pad = (-n) % 32
        msk = np.concatenate([(c == 4).astype(np.uint64), np.ones(pad, np.uint64)]).reshape(-1, 32)
        exp_nm = (msk << np.arange(32, dtype=np.uint64)).sum(1).astype(np.uint32)
        assert (b2 == exp_b2).all()
        assert (nm == exp_nm).all()


@pytest.mark.parametrize("name", ["aog_kat", "aog_ccs", "aog_ont", "aog_clr"])
def test_golden(ctx, name):
    recs = po.read_aog_capture(os.path.join(GOLD, name + ".bin"))
    groups = {}
    for r in recs:
        groups.setdefault((r["m"], r["mm"], r["indel"]), []).append(r)
    for (m, mm, indel), rs in groups.items():
        qa, ta, qo, to, ql, tl, k = jobgen.pack([r["q"] for r in rs], [r["t"] for r in rs], [r["k"] for r in rs])
        q = ctx.seq_upload(qa[:-16]); t = ctx.seq_upload(ta[:-16])
        res = ctx.aog_batch(q, t, qo, to, ql, tl, k, m, mm, indel)
        q.free(); t.free()
        for j, r in enumerate(rs):
            assert res["score"][j] == r["score"], (name, j, len(r["q"]), len(r["t"]), r["k"])
            assert res["n_blocks"][j] == len(r["blocks"])
            o = int(res["block_off"][j])
            assert (res["blocks"][o:o + len(r["blocks"])] == r["blocks"]).all()


@pytest.mark.parametrize("seed", [21, 22, 23, 24])
def test_random_jobs_vs_oracle(ctx, seed):
    rng = np.random.default_rng(seed)
    m, mm, indel = jobgen.SCORINGS[seed & 1]
    run_and_compare(ctx, jobgen.batch(rng, 1500), m, mm, indel)


def test_single_call_mirror(ctx):
    import lra_b200
    q = b"ACGTTTGACCATTAGGACCAGATTTACCA"; t = b"ACGTTGACCATTAGGCCAGATTTTACCA"
    s, b = lra_b200.AffineOneGapAlign(q, len(q), t, len(t), 4, -3, -4, 5, ctx=ctx)
    es, eb, st = po.aog_port(q, t, 4, -3, -4, 5)
    assert s == es and (b == eb).all()


def test_edge_cases(ctx):
    import lra_b200
    from lra_b200 import capi
    q = ctx.seq_upload(b"ACGTACGTAC"); t = ctx.seq_upload(b"ACGTACGTAC")
    r = ctx.aog_batch(q, t, np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.int32), np.zeros(0, np.int32),
                      np.zeros(0, np.int32), 4, -3, -4)
    assert r["n_blocks_total"] == 0
    r = ctx.aog_batch(q, t, [0, 0, 3], [0, 2, 0], [0, 1, 0], [5, 0, 0], [3, 3, 1], 4, -3, -4)   # empty windows are in the domain
    for j, (qq, tt, kk) in enumerate([(b"", b"ACGTA", 3), (b"A", b"", 3), (b"", b"", 1)]):
        es, eb, st = po.aog_port(qq, tt, 4, -3, -4, kk)
        assert r["score"][j] == es and r["n_blocks"][j] == len(eb) == 0
    with pytest.raises(capi.LraB200Error) as e:   # negative length: outside the domain
        ctx.aog_batch(q, t, [0], [0], [-1], [5], [3], 4, -3, -4)
    assert e.value.code == capi.EINVAL
    with pytest.raises(capi.LraB200Error) as e:   # window runs past the arena
        ctx.aog_batch(q, t, [5], [0], [9], [5], [3], 4, -3, -4)
    assert e.value.code == capi.EINVAL
    with pytest.raises(capi.LraB200Error) as e:   # not enough room for the blocks
        ctx.aog_batch(q, t, [0, 0, 0], [0, 0, 0], [10, 10, 10], [10, 10, 10], [3, 3, 3], 4, -3, -4, block_cap=2)
    assert e.value.code == capi.EOVERFLOW
    r = ctx.aog_batch(q, t, [0, 0, 0], [0, 0, 0], [10, 10, 10], [10, 10, 10], [3, 3, 3], 4, -3, -4, block_cap=3)
    assert (r["score"] == 40).all() and (r["n_blocks"] == 1).all()
    q.free(); t.free()


@pytest.mark.parametrize("profile", ["ont", "ccs"])
def test_full_size_batch_properties_and_reference_sample(ctx, profile):
    """A bench-sized batch: size-independent properties on every job, and a bit-exact comparison of a sample of it with
    the reference header itself (prebuilt oracle/_ref/libref_lra.so) or, if that is absent, the C restatement."""
    n_jobs = 1_000_000 if profile == "ont" else 300_000
    genome = synth.gen_ref(20_000_000, 1, 99)[0][1]
    jobs = workload.make_jobs(profile, n_jobs, 5, len(genome), workload.host_genome_fetcher(genome))
    m, mm, indel = jobs["scoring"]
    q = ctx.seq_upload(jobs["q_arena"][:-16]); t = ctx.seq_upload(genome)
    r = ctx.aog_batch(q, t, jobs["q_off"], jobs["t_off"], jobs["q_len"], jobs["t_len"], jobs["k"], m, mm, indel)
    q.free(); t.free()
    checked = workload.check_blocks_property(jobs, r, m, mm, indel)
    assert checked > 0.9 * n_jobs
    # sample vs the reference
    sel = np.random.default_rng(1).choice(n_jobs, 60000, replace=False)
    sub = {kk: jobs[kk][sel] for kk in ["q_off", "t_off_compact", "q_len", "t_len", "k"]}
    if po.ref() is not None:
        es, enb, eoff, eblk = po.aog_batch_ref(jobs["q_arena"], jobs["t_arena_compact"], sub["q_off"], sub["t_off_compact"],
                                               sub["q_len"], sub["t_len"], sub["k"], m, mm, indel, nthreads=os.cpu_count() or 1)
    else:
        es, enb, eoff, eblk, _ = po.aog_batch_port(jobs["q_arena"], jobs["t_arena_compact"], sub["q_off"], sub["t_off_compact"],
                                                   sub["q_len"], sub["t_len"], sub["k"], m, mm, indel)
    assert (r["score"][sel] == es).all() and (r["n_blocks"][sel] == enb).all()
    go = r["block_off"][sel].astype(np.int64)
    tot = int(enb.sum())
    within = np.arange(tot) - np.repeat(np.cumsum(enb) - enb, enb)
    assert (r["blocks"][np.repeat(go, enb) + within] == eblk[np.repeat(eoff, enb) + within]).all()
