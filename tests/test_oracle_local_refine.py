"""The C restatement of LocalIndex::IndexSeq (a12) and REFINEclusters (a13) pinned on the unmodified reference headers."""
import numpy as np
import pytest

from oracle import pyoracle as po
import refinegen

needs_ref = pytest.mark.skipif(po.ref() is None, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
B = np.frombuffer(b"ACGT", np.uint8)


def _same_index(a, b):
    return (a.seq_off == b.seq_off).all() and (a.bnd == b.bnd).all() and len(a.mins) == len(b.mins) and (a.mins == b.mins).all()


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_index_seq_matches_reference(seed):
    rng = np.random.default_rng(seed)
    seqs = []
    for L in (0, 5, 13, 14, 15, 100, 2047, 2048, 2049, 4096, 10000):
        s = B[rng.integers(0, 4, L)].copy()
        seqs.append(s)
        if L >= 100:
            t = s.copy(); t[L // 3:L // 3 + 9] = ord("N"); t[-3:] = ord("N"); seqs.append(t)                    # N runs, N at the end
            u = s.copy(); u[: L // 2] = np.tile(B[rng.integers(0, 4, 2)], L)[: L // 2]; seqs.append(u)            # dinucleotide repeat: ties, frequent tuples
            v = np.full(L, ord("A"), np.uint8); seqs.append(v)                                                      # homopolymer
            x = s.copy(); x[rng.random(L) < 0.3] = ord("N"); seqs.append(x)                                         # N-riddled: few valid windows
    for s in seqs:
        for mf in (5, 15):
            assert _same_index(po.local_index(s, max_freq=mf, which="port"), po.local_index(s, max_freq=mf, which="ref")), (len(s), mf)
    # a genome of several contigs: IndexFile appends contig after contig with a running offset
    contigs = [seqs[-1], seqs[3], seqs[-6]]
    assert _same_index(po.local_index(contigs, which="port"), po.local_index(contigs, which="ref"))


@needs_ref
@pytest.mark.parametrize("seed", [11, 12])
def test_refine_clusters_matches_reference(seed):
    case = refinegen.make_case(seed)
    a = refinegen.expected(case, "port"); b = refinegen.expected(case, "ref")
    assert len(a) == len(b) and len(a) > 10
    st = [x["status"] for x in b]
    assert 0 in st and 1 in st and 2 in st
    assert sum(len(x["rq"]) for x in b) > 1000
    for i, (x, y) in enumerate(zip(a, b)):
        assert refinegen.same(x, y), (i, case["clusters"][i]["read"], case["clusters"][i]["strand"])


def test_compare_lists_local_small():
    # tuples 5 and 9 are shared; 9 occurs twice in the query, 5 twice in the target
    q = np.array([3, 5 | (7 << 20), 9 | (1 << 20), 9 | (2 << 20), 12], np.uint32)
    t = np.array([5 | (100 << 20), 5 | (200 << 20), 8, 9 | (300 << 20), 11], np.uint32)
    rq, rt = po.compare_lists_local_port(q, t, 15)
    pairs = sorted(zip((rq >> 20).tolist(), (rt >> 20).tolist()))
    assert pairs == [(1, 300), (2, 300), (7, 100), (7, 200)]
    rq, rt = po.compare_lists_local_port(q, t, 1)     # a tuple occurring more than maxFreq times IN THE QUERY is skipped
    assert sorted(zip((rq >> 20).tolist(), (rt >> 20).tolist())) == [(7, 100), (7, 200)]


@needs_ref
@pytest.mark.parametrize("seed,limit", [(13, 1), (14, 1), (13, 0)])
def test_refine_splitchain_matches_reference(seed, limit):
    case = refinegen.make_case(seed)
    chains = refinegen.make_chains(case, seed)
    a = refinegen.expected_chains(case, chains, "port", limitrefine=limit); b = refinegen.expected_chains(case, chains, "ref", limitrefine=limit)
    assert len(a) == len(b) and len(a) >= 8
    assert sum(len(x["rq"]) for x in b) > 1000
    bad = [i for i, (x, y) in enumerate(zip(a, b)) if not refinegen.same_chain(x, y)]
    assert not bad, (bad, [(len(a[i]["rq"]), len(b[i]["rq"])) for i in bad])
