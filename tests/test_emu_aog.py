"""Kernel-logic parity on the CPU: the device code of lra_b200/csrc/aog_*.cuh executed through the SIMT emulator
(tests/simt/cuda_emu.h) must be bit-identical to the oracle for AffineOneGapAlign (reference AffineOneGapAlign.h:157-649).
The same comparisons run against the real kernels on the B200 in test_gpu_aog.py."""
import os
import numpy as np
import pytest

from oracle import pyoracle as po
import jobgen
import emu_lib

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def check_batch(batch, m, mm, indel, **kw):
    qa, ta, qo, to, ql, tl, k = batch
    es, enb, eoff, eblk, st = po.aog_batch_port(qa, ta, qo, to, ql, tl, k, m, mm, indel)
    assert (st == 0).all()
    err, s, nb, off, blk, cells = emu_lib.aog_batch(qa, ta, qo, to, ql, tl, k, m, mm, indel, **kw)
    assert err == 0
    bad = np.flatnonzero((s != es) | (nb != enb))
    assert len(bad) == 0, (bad[:5], s[bad[:5]], es[bad[:5]], ql[bad[:5]], tl[bad[:5]], k[bad[:5]])
    for j in range(len(qo)):
        a = blk[int(off[j]):int(off[j]) + nb[j]]
        e = eblk[eoff[j]:eoff[j] + enb[j]]
        assert (a == e).all(), (j, ql[j], tl[j], k[j], a[:4], e[:4])
    assert cells > 0


@pytest.mark.parametrize("mode", [dict(use_band=0), dict(use_band=1), dict(force_literal=1)])
@pytest.mark.parametrize("seed", [11, 12])
def test_emu_random_jobs(seed, mode):
    rng = np.random.default_rng(seed)
    m, mm, indel = jobgen.SCORINGS[seed & 1]
    check_batch(jobgen.batch(rng, 300), m, mm, indel, **mode)


@pytest.mark.parametrize("name", ["aog_kat", "aog_ccs", "aog_ont"])
def test_emu_golden(name):
    recs = po.read_aog_capture(os.path.join(GOLD, name + ".bin"))
    rng = np.random.default_rng(3)
    if len(recs) > 400:
        recs = [recs[i] for i in sorted(rng.choice(len(recs), 400, replace=False))]
    groups = {}
    for r in recs:
        groups.setdefault((r["m"], r["mm"], r["indel"]), []).append(r)
    for (m, mm, indel), rs in groups.items():
        batch = jobgen.pack([r["q"] for r in rs], [r["t"] for r in rs], [r["k"] for r in rs])
        err, s, nb, off, blk, _ = emu_lib.aog_batch(*batch, m, mm, indel, use_band=1)
        assert err == 0
        for j, r in enumerate(rs):
            assert s[j] == r["score"] and nb[j] == len(r["blocks"]), (j, len(r["q"]), len(r["t"]), r["k"])
            assert (blk[int(off[j]):int(off[j]) + nb[j]] == r["blocks"]).all()


def test_emu_block_overflow_is_reported():
    rng = np.random.default_rng(4)
    batch = jobgen.batch(rng, 64, kinds=["similar"])
    err, *_ = emu_lib.aog_batch(*batch, 4, -3, -4, block_cap=3)
    assert err & 1
