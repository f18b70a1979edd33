"""Pins the C restatement oracle/aog.c of AffineOneGapAlign (reference AffineOneGapAlign.h:157-649).

(1) against the committed golden vectors (outputs of the unmodified reference; tools/make_golden.py);
(2) against the real reference header live (oracle/_ref/libref_lra.so) on seeded random jobs, when that
    library exists (it is built wherever /root/reference is present and shipped prebuilt to the GPU box).
"""
import os
import numpy as np
import pytest

from oracle import pyoracle as po
import jobgen

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["aog_kat", "aog_ccs", "aog_ont", "aog_clr"])
def test_restatement_matches_golden(name):
    recs = po.read_aog_capture(os.path.join(GOLD, name + ".bin"))
    assert len(recs) > 50
    for r in recs:
        s, b, st = po.aog_port(r["q"], r["t"], r["m"], r["mm"], r["indel"], r["k"])
        assert st == 0
        assert s == r["score"]
        assert b.shape == r["blocks"].shape and (b == r["blocks"]).all()


def test_golden_covers_both_modes():
    recs = po.read_aog_capture(os.path.join(GOLD, "aog_ccs.bin"))
    two = 0
    for r in recs:
        ql, tl = len(r["q"]), len(r["t"])
        d = max(1, min(ql, tl)); k = min(d, r["k"])
        two += d + 2 * k < max(ql, tl)
    assert two > 100 and len(recs) - two > 1000


@pytest.mark.skipif(po.ref() is None, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_restatement_matches_live_reference(seed):
    rng = np.random.default_rng(seed)
    n_two = 0
    for it in range(1500):
        q, t, k = jobgen.random_job(rng)
        m, mm, indel = jobgen.SCORINGS[it & 1]
        s0, b0 = po.aog_ref(q, t, m, mm, indel, k)
        s1, b1, st = po.aog_port(q, t, m, mm, indel, k)
        assert st == 0, (q, t, k)
        assert s0 == s1, (q, t, k, s0, s1)
        assert b0.shape == b1.shape and (b0 == b1).all(), (q, t, k)
        d = max(1, min(len(q), len(t))); kk = min(d, k)
        n_two += d + 2 * kk < max(len(q), len(t))
    assert n_two > 100


def test_batch_form_equals_single():
    rng = np.random.default_rng(5)
    qa, ta, qo, to, ql, tl, k = jobgen.batch(rng, 200)
    score, nb, off, blocks, st = po.aog_batch_port(qa, ta, qo, to, ql, tl, k, 4, -3, -4)
    for j in range(len(qo)):
        q = bytes(qa[qo[j]:qo[j] + ql[j]]); t = bytes(ta[to[j]:to[j] + tl[j]])
        s, b, _ = po.aog_port(q, t, 4, -3, -4, int(k[j]))
        assert s == score[j] and len(b) == nb[j]
        assert (blocks[off[j]:off[j] + nb[j]] == b).all()
