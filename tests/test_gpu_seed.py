"""GPU parity tests for a1-a5 (the seeding prefix of MapRead) through the C ABI against the oracle: per read, the matches in
the reference's allMatches order with their strand flags, and the sorted minimizer counts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po
import seedgen


@pytest.fixture(scope="module")
def ctx():
    import lra_b200
    c = lra_b200.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("seed,k,w,mf,n_reads", [(1, 17, 10, 150, 64), (2, 15, 10, 2, 40), (3, 17, 20, 150, 40), (5, 19, 10, 30, 40)])
def test_seed_batch_vs_oracle(ctx, seed, k, w, mf, n_reads):
    case = seedgen.make_case(seed, glen=200000, n_reads=n_reads, k=k, w=w)
    which = "ref" if po.ref() is not None else "port"
    exp = seedgen.expected(case, mf, which)
    reads = ctx.seq_upload(case["arena"][:-16]); genome = ctx.seq_upload(case["genome"][:-16])
    idx = ctx.index_upload(case["idx_t"], case["idx_pos"])
    o = ctx.seed_batch(reads, genome, idx, case["read_off"], case["read_len"], k, w, mf)
    ctx.index_free(idx); reads.free(); genome.free()
    assert o["n_matches"] == sum(len(e[0]) for e in exp) and o["n_matches"] > 100
    for r, e in enumerate(exp):
        a, b = int(o["match_off"][r]), int(o["match_off"][r + 1])
        assert b - a == len(e[0]), r
        for key, ev in zip(["q_t", "q_pos", "t_t", "t_pos", "strand"], e):
            assert (o[key][a:b] == ev).all(), (r, key)
    names = [s["name"] for s in ctx.kernel_stats()]
    assert names == ["seed_minimizers", "seed_sort", "seed_compare<count>", "seed_scan", "seed_compare<emit>"]


def test_seed_overflow_and_revcomp(ctx):
    from lra_b200 import capi
    case = seedgen.make_case(7, glen=120000, n_reads=12)
    reads = ctx.seq_upload(case["arena"][:-16]); genome = ctx.seq_upload(case["genome"][:-16])
    idx = ctx.index_upload(case["idx_t"], case["idx_pos"])
    with pytest.raises(capi.LraB200Error) as e:
        ctx.seed_batch(reads, genome, idx, case["read_off"], case["read_len"], 17, 10, 150, match_cap=5)
    assert e.value.code == capi.EOVERFLOW
    rc = ctx.seq_revcomp(reads, case["read_off"], case["read_len"])
    b2, nm = rc.download()
    n = len(case["arena"]) - 16
    got2 = (np.repeat(b2, 16)[:n] >> (2 * (np.arange(n) % 16)).astype(np.uint32)) & 3
    gotn = (np.repeat(nm, 32)[:n] >> (np.arange(n) % 32).astype(np.uint32)) & 1
    got = np.where(gotn == 1, 4, got2)
    lut = np.full(256, 4, np.uint8); lut[[65, 67, 71, 84]] = [0, 1, 2, 3]
    for r in range(len(case["read_off"])):
        o, L = int(case["read_off"][r]), int(case["read_len"][r])
        assert (got[o:o + L] == lut[seedgen.COMP[case["arena"][o:o + L][::-1]]]).all()
    ctx.index_free(idx); reads.free(); genome.free(); rc.free()


def test_seed_batch_rejects_bad_read_descriptors(ctx):
    """reads beyond the arena, overlapping or out of order are refused (the minimizer scratch is indexed by read_off)"""
    from lra_b200 import capi
    case = seedgen.make_case(7, glen=120000, n_reads=4, k=17, w=10)
    reads = ctx.seq_upload(case["arena"][:-16]); genome = ctx.seq_upload(case["genome"][:-16])
    idx = ctx.index_upload(case["idx_t"], case["idx_pos"])
    ro, rl = case["read_off"].copy(), case["read_len"].copy()
    for bad_off, bad_len in ((ro, np.where(np.arange(len(rl)) == len(rl) - 1, rl + 10 ** 6, rl).astype(rl.dtype)), (ro[::-1].copy(), rl[::-1].copy()),
                             (np.where(np.arange(len(ro)) == 1, ro[0] + 1, ro).astype(ro.dtype), rl)):
        with pytest.raises(capi.LraB200Error) as e:
            ctx.seed_batch(reads, genome, idx, bad_off, bad_len, 17, 10, 150)
        assert e.value.code == capi.EINVAL
    ctx.index_free(idx); reads.free(); genome.free()
