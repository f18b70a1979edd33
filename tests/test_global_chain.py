"""a24 GlobalChain + PrioritySearchTree: the known answer of the reference's own driver (TestGlobalChain.cpp), the C restatement pinned on
the reference headers, the kernel logic through the emulator and the real kernel through the C ABI."""
import os
import numpy as np
import pytest

from oracle import pyoracle as po
import chaingen

needs_ref = pytest.mark.skipif(po.ref() is None, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "global_chain_kat.txt")


def kat_text(frag, chain, score):
    return "Opt of size %d\n" % len(chain) + "".join("%d\t%d\t%d\t%d\t%d\n" % (tuple(frag[i]) + (score[i],)) for i in chain)


def test_known_answer_of_the_reference_driver():
    """tests/golden/global_chain_kat.txt is the verbatim output of the reference's TestGlobalChain.cpp (oracle/Makefile: test_global_chain)."""
    f = chaingen.KAT
    chain, score, prev = po.global_chain(f, f[:, 2] - f[:, 0], "port")
    assert kat_text(f, chain, score) == open(GOLDEN).read()


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3, 4, 6])
def test_oracle_matches_reference(seed):
    for f, sc in chaingen.problems(seed):
        a = po.global_chain(f, sc, "port"); b = po.global_chain(f, sc, "ref")
        assert (a[0] == b[0]).all() and len(a[0]) == len(b[0]) and (a[1] == b[1]).all() and (a[2] == b[2]).all(), (seed, len(f))


def _batch(seeds):
    probs = [(chaingen.KAT, (chaingen.KAT[:, 2] - chaingen.KAT[:, 0]).astype(np.int32)), (np.zeros((0, 4), np.int32), np.zeros(0, np.int32))]
    for s in seeds:
        probs += chaingen.problems(s)
    off = np.zeros(len(probs) + 1, np.uint64); off[1:] = np.cumsum([len(f) for f, _ in probs])
    return probs, np.concatenate([f for f, _ in probs]), off, np.concatenate([s for _, s in probs])


def _check(probs, off, o):
    for p, (f, sc) in enumerate(probs):
        a, b = int(off[p]), int(off[p + 1])
        chain, score, prev = po.global_chain(f, sc, "port")
        assert o["chain_len"][p] == len(chain), p
        assert (o["chain"][a:a + len(chain)] == chain).all() and (o["score"][a:b] == score).all() and (o["prev"][a:b] == prev).all(), p
    assert kat_text(chaingen.KAT, o["chain"][:o["chain_len"][0]], o["score"][:9]) == open(GOLDEN).read()


def test_emu_global_chain():
    import emu_lib
    probs, frag, off, score = _batch([1, 6])
    _check(probs, off, emu_lib.global_chain(frag, off, score))


@pytest.mark.gpu
def test_gpu_global_chain():
    import lra_b200
    ctx = lra_b200.Context(0)
    probs, frag, off, score = _batch([1, 2, 3, 4, 6, 7, 8, 9])
    _check(probs, off, ctx.global_chain_batch(frag, off, score))
    assert ctx.global_chain_batch(np.zeros((0, 4), np.int32), np.zeros(1, np.uint64), np.zeros(0, np.int32))["chain_len"].size == 0
    ctx.close()
