"""a6 DiagonalSort / AntiDiagonalSort / CartesianSort / CartesianTargetSort (Sorting.h): oracle pinned on the reference, kernel logic
through the emulator, the real kernel through the C ABI."""
import numpy as np
import pytest

from oracle import pyoracle as po

needs_ref = pytest.mark.skipif(po.ref() is None, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")


def segments(seed, sizes=(0, 1, 2, 3, 17, 33, 100, 257, 1000, 2048, 2049, 5000)):
    """Anchor lists: near-diagonal anchors with duplicates and ties on every key, genome positions above 2^31 (wrap of q + t),
    a few read positions larger than the genome position (negative diagonals)."""
    rng = np.random.default_rng(seed)
    qs, ts, off = [], [], [0]
    for n in sizes:
        q = rng.integers(0, 30000, n).astype(np.uint32)
        base = int(rng.choice([0, 5000, 2**31 - 10000, 2**32 - 40000]))
        t = (q.astype(np.int64) + base + rng.integers(-30, 30, n)).clip(0, 2**32 - 1).astype(np.uint32)
        if n > 10:
            d = rng.integers(0, n, n // 5); q[d] = q[(d + 1) % n]; t[d] = t[(d + 1) % n]        # exact duplicates
            e = rng.integers(0, n, n // 5); q[e] = q[(e + 3) % n]                                  # equal q, different t
        qs.append(q); ts.append(t); off.append(off[-1] + n)
    return np.concatenate(qs), np.concatenate(ts), np.array(off, np.uint64)


@needs_ref
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_oracle_matches_reference(mode):
    q, t, off = segments(5 + mode)
    for s in range(len(off) - 1):
        a, b = int(off[s]), int(off[s + 1])
        pq, pt, perm = po.sort_matches(mode, q[a:b], t[a:b], "port"); rq, rt, _ = po.sort_matches(mode, q[a:b], t[a:b], "ref")
        assert (pq == rq).all() and (pt == rt).all()
        assert (q[a:b][perm] == pq).all() and (t[a:b][perm] == pt).all()


def check(q, t, off, mode, oq, ot, perm):
    for s in range(len(off) - 1):
        a, b = int(off[s]), int(off[s + 1])
        eq, et, _ = po.sort_matches(mode, q[a:b], t[a:b], "port")
        assert (oq[a:b] == eq).all() and (ot[a:b] == et).all(), (mode, s)
        if b > a:
            assert perm[a:b].min() >= a and perm[a:b].max() < b and len(set(perm[a:b].tolist())) == b - a, (mode, s)
    assert (q[perm] == oq).all() and (t[perm] == ot).all()


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_emu_sort(mode):
    import emu_lib
    q, t, off = segments(11 + mode, sizes=(0, 1, 2, 5, 33, 300, 2048, 2500))
    oq, ot, perm = emu_lib.sort_matches(mode, q, t, off)
    check(q, t, off, mode, oq, ot, perm)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_gpu_sort(mode):
    import lra_b200
    ctx = lra_b200.Context(0)
    q, t, off = segments(21 + mode, sizes=(0, 1, 2, 3, 17, 33, 100, 257, 1000, 2048, 2049, 5000, 70000) + (400,) * 300)
    oq, ot, perm = ctx.sort_matches_batch(mode, q, t, off)
    check(q, t, off, mode, oq, ot, perm)
    assert ctx.kernel_stats()[0]["name"].startswith("sort_pairs")
    e = ctx.sort_matches_batch(mode, np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(1, np.uint64))
    assert len(e[0]) == 0
    ctx.close()
