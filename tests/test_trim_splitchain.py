"""TrimSplitChainDiagonal (the step of the low-accuracy pipeline right after Refine_splitchain): restatement pinned on the unmodified reference, kernels (the
a6 Cartesian sort + the trim walk) through the emulator (CPU) and the C ABI (GPU).  Integer work: bit-exact."""
import numpy as np
import pytest

from oracle import pyoracle as po
import trimgen

HAVE_REF = po.ref() is not None


def cases(seed, n):
    rng = np.random.default_rng(seed)
    return [trimgen.case(rng) for _ in range(n)]


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
def test_oracle_matches_reference():
    rem = 0
    for cq, ct, st, q, t in cases(1, 2000):
        a = po.trim_splitchain(cq, ct, st, q, t, which="ref"); b = po.trim_splitchain(cq, ct, st, q, t)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
        rem += b[2]
    assert rem > 5000


def batch(cs):
    c_off = np.zeros(len(cs) + 1, np.uint64); c_off[1:] = np.cumsum([len(c[0]) for c in cs])
    m_off = np.zeros(len(cs) + 1, np.uint64); m_off[1:] = np.cumsum([len(c[3]) for c in cs])
    return (np.concatenate([c[0] for c in cs]), np.concatenate([c[1] for c in cs]), c_off, np.array([c[2] for c in cs], np.uint8),
            np.concatenate([c[3] for c in cs]), np.concatenate([c[4] for c in cs]), m_off)


def check(res, m_off, cs, which):
    q, t, keep, removed = res
    for c, (cq, ct, st, q0, t0) in enumerate(cs):
        x = po.trim_splitchain(cq, ct, st, q0, t0, which=which)
        a, b = int(m_off[c]), int(m_off[c + 1])
        k = keep[a:b].astype(bool)
        assert np.array_equal(q[a:b][k], x[0]) and np.array_equal(t[a:b][k], x[1]) and int(removed[c]) == x[2], c


def test_emu_trim_splitchains():
    import emu_lib
    cs = cases(3, 400)
    b = batch(cs)
    check(emu_lib.trim_splitchains(*b), b[6], cs, "port")


@pytest.mark.gpu
def test_gpu_trim_splitchains():
    import lra_b200
    ctx = lra_b200.Context(0)
    cs = cases(5, 5000)
    b = batch(cs)
    check(ctx.trim_splitchains_batch(*b), b[6], cs, "ref" if HAVE_REF else "port")
    ctx.close()
