"""a16 chain filters (RemoveSmallPairedIndels, RemovePairedIndels x3, RemoveSpuriousAnchors, RemoveSpuriousJump; Chain.h:546-960): oracle
pinned on the unmodified templates, kernel logic through the emulator, the real kernel through the C ABI."""
import numpy as np
import pytest

from oracle import pyoracle as po
import chainfgen

needs_ref = pytest.mark.skipif(po.ref() is None, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")


@needs_ref
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])
def test_oracle_matches_reference(mode):
    removed = 0
    for seed in (1, 2, 3, 4):
        for q, t, ln, st in chainfgen.chains(seed):
            a = po.chain_filter(mode, q, t, ln, st, "port"); b = po.chain_filter(mode, q, t, ln, st, "ref")
            assert (a == b).all(), (mode, seed, len(q))
            removed += int((b == 0).sum())
    assert removed > 0


def _batch(seeds):
    cs = [c for s in seeds for c in chainfgen.chains(s)]
    off = np.zeros(len(cs) + 1, np.uint64); off[1:] = np.cumsum([len(c[0]) for c in cs])
    cat = lambda k, dt: np.concatenate([c[k] for c in cs]).astype(dt)
    return cs, cat(0, np.uint32), cat(1, np.uint32), cat(2, np.uint32), cat(3, np.uint8), off


def _check(cs, off, mode, keep):
    for i, (q, t, ln, st) in enumerate(cs):
        a, b = int(off[i]), int(off[i + 1])
        assert (keep[a:b] == po.chain_filter(mode, q, t, ln, st, "port")).all(), (mode, i)


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])
def test_emu_chain_filter(mode):
    import emu_lib
    cs, q, t, ln, st, off = _batch([5])
    _check(cs, off, mode, emu_lib.chain_filter(mode, q, t, ln, st, off))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])
def test_gpu_chain_filter(mode):
    import lra_b200
    ctx = lra_b200.Context(0)
    cs, q, t, ln, st, off = _batch([6, 7, 8, 9, 10, 11])
    keep = ctx.chain_filter_batch(mode, q, t, ln, st, off)
    _check(cs, off, mode, keep)
    assert (keep == 0).sum() > 0
    ctx.close()
