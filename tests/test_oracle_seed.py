"""Pins oracle/seed.c (a1-a5: the seeding prefix of MapRead) against the unmodified reference headers through
oracle/_ref/libref_lra.so, and against the committed golden (tests/golden/seed_small.npz: a reference-built .mms index of a
small synthetic genome + reads + the reference's own matches)."""
import os
import numpy as np
import pytest

from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(po.ref() is None, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
B = np.frombuffer(b"ACGT", np.uint8)


def rand_seq(rng, n, kind):
    if kind == "plain":
        s = B[rng.integers(0, 4, n)]
    elif kind == "lowcomplexity":      # many equal k-mers: ties in the window minimum, equal keys in the sort
        unit = B[rng.integers(0, 4, int(rng.integers(1, 9)))]
        s = np.tile(unit, n // len(unit) + 1)[:n].copy()
        m = rng.random(n) < 0.03
        s[m] = B[rng.integers(0, 4, int(m.sum()))]
    elif kind == "palindromic":        # a sequence followed by its reverse complement: same canonical k-mers on both strands
        h = B[rng.integers(0, 4, n // 2)]
        comp = {65: 84, 67: 71, 71: 67, 84: 65}
        s = np.concatenate([h, np.array([comp[c] for c in h[::-1]], np.uint8)])
    else:                              # N runs
        s = B[rng.integers(0, 4, n)]
        for _ in range(int(rng.integers(1, 6))):
            a = int(rng.integers(0, n)); s[a:a + int(rng.integers(1, 40))] = ord("N")
    return s


@needs_ref
@pytest.mark.parametrize("kind", ["plain", "lowcomplexity", "palindromic", "withN"])
def test_minimizers_and_sort_match_reference(kind):
    rng = np.random.default_rng(hash(kind) & 0xFFFF)
    for it in range(60):
        n = int(rng.choice([5, 20, 30, 45, 100, 400, 3000]))
        k, w = [(17, 10), (17, 20), (15, 10), (19, 10), (10, 5)][it % 5]
        s = rand_seq(rng, n, kind)
        t0, p0 = po.store_minimizers(s, k, w, "ref")
        t1, p1 = po.store_minimizers(s, k, w, "port")
        assert len(t0) == len(t1) and (t0 == t1).all() and (p0 == p1).all(), (kind, n, k, w)
        a0 = po.sort_minimizers(t0, p0, "ref"); a1 = po.sort_minimizers(t0, p0, "port")
        assert (a0[0] == a1[0]).all() and (a0[1] == a1[1]).all(), (kind, n, k, w)


@needs_ref
def test_sort_restates_libstdcxx_introsort_on_heavy_ties_and_big_inputs():
    rng = np.random.default_rng(5)
    for n in [0, 1, 2, 15, 16, 17, 18, 33, 100, 1000, 5000, 40000]:
        for nkeys in [1, 3, 50, 10**6]:
            t = rng.integers(0, nkeys, n).astype(np.uint64) | (rng.integers(0, 2, n).astype(np.uint64) << np.uint64(63))
            pos = np.arange(n, dtype=np.uint32)
            a0 = po.sort_minimizers(t, pos, "ref"); a1 = po.sort_minimizers(t, pos, "port")
            assert (a0[0] == a1[0]).all() and (a0[1] == a1[1]).all(), (n, nkeys)
    # organ-pipe / sawtooth patterns push quicksort towards its depth limit (heapsort path)
    for n in [3000, 20000]:
        t = np.concatenate([np.arange(n // 2), np.arange(n // 2)[::-1]]).astype(np.uint64)
        pos = np.arange(len(t), dtype=np.uint32)
        a0 = po.sort_minimizers(t, pos, "ref"); a1 = po.sort_minimizers(t, pos, "port")
        assert (a0[0] == a1[0]).all() and (a0[1] == a1[1]).all()


@needs_ref
@pytest.mark.parametrize("kind", ["plain", "lowcomplexity", "palindromic"])
def test_compare_lists_matches_reference(kind):
    rng = np.random.default_rng(11)
    for it in range(25):
        g = rand_seq(rng, int(rng.choice([300, 2000, 20000])), kind)
        k, w = [(17, 10), (15, 10), (10, 5)][it % 3]
        gt, gp = po.sort_minimizers(*po.store_minimizers(g, k, w, "ref"), "ref")
        a = int(rng.integers(0, max(1, len(g) - 200)))
        r = g[a:a + int(rng.integers(50, 1500))].copy()
        m = rng.random(len(r)) < 0.05
        r[m] = B[rng.integers(0, 4, int(m.sum()))]
        if rng.random() < 0.5:
            comp = np.zeros(256, np.uint8); comp[[65, 67, 71, 84, 78]] = [84, 71, 67, 65, 78]
            r = comp[r[::-1]]
        qt, qp = po.sort_minimizers(*po.store_minimizers(r, k, w, "ref"), "ref")
        for mf in [1, 2, 50]:
            x0 = po.compare_lists(qt, qp, gt, gp, mf, "ref"); x1 = po.compare_lists(qt, qp, gt, gp, mf, "port")
            assert all(len(u) == len(v) and (u == v).all() for u, v in zip(x0, x1)), (kind, it, mf)


@needs_ref
def test_seed_read_matches_reference_end_to_end():
    rng = np.random.default_rng(13)
    g = rand_seq(rng, 60000, "plain")
    g[20000:24000] = g[1000:5000]          # a repeat, so that some k-mers occur more than once
    k, w = 17, 10
    gt, gp = po.sort_minimizers(*po.store_minimizers(g, k, w, "ref"), "ref")
    for it in range(20):
        a = int(rng.integers(0, len(g) - 6000)); r = g[a:a + 5000].copy()
        m = rng.random(len(r)) < 0.08
        r[m] = B[rng.integers(0, 4, int(m.sum()))]
        x0 = po.seed_read(r, g, gt, gp, k, w, 150, "ref"); x1 = po.seed_read(r, g, gt, gp, k, w, 150, "port")
        assert len(x0[0]) > 10
        assert all(len(u) == len(v) and (u == v).all() for u, v in zip(x0, x1)), it
