"""Pins the C restatement oracle/indel_refine.c of IndelRefineAlignment (reference IndelRefine.h:53-784) against the
committed golden segments (outputs of the unmodified reference captured by oracle/lra_capture.cpp; tools/make_golden.py)."""
import os
import numpy as np
import pytest

from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["ir_ccs", "ir_ont", "ir_clr"])
def test_restatement_matches_golden(name):
    recs = po.read_ir_capture(os.path.join(GOLD, name + ".bin"))
    assert len(recs) >= 10
    n_dp = 0
    for r in recs:
        bo, st, cells = po.indel_refine_port(r["read"], r["twin"], r["t_win_off"], r["contig_len"], r["blocks_in"], r["k"],
                                             r["match"], r["mismatch"], r["indel"], r["end_align"])
        assert st == 0
        assert bo.shape == r["blocks_out"].shape and (bo == r["blocks_out"]).all()
        n_dp += cells > 0
    assert n_dp >= len(recs) - 1


def test_group_dump_is_consistent():
    recs = po.read_ir_capture(os.path.join(GOLD, "ir_ccs.bin"))[:6]
    for r in recs:
        bo, st, groups = po.indel_refine_groups_port(r["read"], r["twin"], r["t_win_off"], r["contig_len"], r["blocks_in"], r["k"],
                                                     r["match"], r["mismatch"], r["indel"], r["end_align"])
        assert (bo == r["blocks_out"]).all() and len(groups) > 0
        for g in groups:
            assert len(g["qS"]) == g["tLen"] and (np.diff(g["qS"]) >= 0).all() and (np.diff(g["qE"]) >= 0).all()
            b = g["blocks"]
            assert b[0, 0] == g["qStart"] and b[0, 1] == g["tStart"]
            assert b[-1, 0] + b[-1, 2] == g["qStart"] + g["qSeqLen"] and b[-1, 1] + b[-1, 2] == g["tStart"] + g["tSeqLen"]


@pytest.mark.skipif(po.ref() is None, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
@pytest.mark.parametrize("profile,n", [("ont", 12), ("ccs", 40), ("clr", 10)])
def test_restatement_matches_live_reference_on_synthetic_segments(profile, n):
    """Whole IndelRefineAlignment through the unmodified reference header (oracle/ref_wrap.cpp) vs the restatement."""
    import synth, workload
    genome = synth.gen_ref(3_000_000, 1, 77)[0][1]
    sb = workload.make_segments(profile, n, 5, len(genome), workload.host_genome_fetcher(genome))
    nref, off, blk = po.indel_refine_batch_ref(sb, sb["t_arena_compact"], sb["t_base_compact"], nthreads=2)
    outs = po.indel_refine_batch_port(sb, sb["t_arena_compact"], sb["t_base_compact"])
    for s, o in enumerate(outs):
        assert len(o) == nref[s] and (o == blk[int(off[s]):int(off[s]) + nref[s]]).all(), (profile, s)
