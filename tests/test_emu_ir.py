"""Kernel-logic parity on the CPU for a19: the IndelRefine DP kernels (lra_b200/csrc/ir_kernels.cuh) executed through the
SIMT emulator must reproduce, group by group, the blocks of the pinned oracle on segments captured from the reference."""
import os
import numpy as np
import pytest

from oracle import pyoracle as po
import irgen
import emu_lib

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def run(name, n_rec, force_generic=0, max_groups=None):
    recs = po.read_ir_capture(os.path.join(GOLD, name + ".bin"))[:n_rec]
    rg = irgen.groups_of_records(recs)
    if max_groups:
        rg = rg[:max_groups]
    r0 = recs[0]
    gb, expect = irgen.pack_groups(recs, rg, (r0["match"], r0["mismatch"], r0["indel"]))
    err, nb, off, blk, cells = emu_lib.ir_dp_batch(gb, force_generic=force_generic)
    assert err == 0
    assert cells == int(sum((g["qE"] - g["qS"] + 1).sum() for _, g in rg))
    for j, e in enumerate(expect):
        assert nb[j] == len(e), (name, j, nb[j], len(e))
        assert (blk[int(off[j]):int(off[j]) + nb[j]] == e).all(), (name, j)
    return len(expect)


def test_emu_ir_ccs():
    assert run("ir_ccs", 4) > 50


def test_emu_ir_ont():
    assert run("ir_ont", 1) >= 1


def test_emu_ir_clr_w64():
    assert run("ir_clr", 1) >= 1


def test_emu_ir_pipe_kernel_lane_orders():
    """the row-pipeline kernel exchanges rows through shared memory between warp barriers: replay it with the lanes
    scheduled in descending and in random order between collectives (hardware leaves that order unspecified)"""
    try:
        for mode in (1, 20261017):
            emu_lib.set_lane_order(mode)
            assert run("ir_ont", 1) >= 1
    finally:
        emu_lib.set_lane_order(0)


def test_emu_ir_pipe_two_cells_per_step():
    assert run("ir_ont", 1, force_generic=5) >= 1


def test_emu_ir_scan_warp_kernel():
    assert run("ir_ont", 1, force_generic=4) >= 1


def test_emu_ir_thread_kernels_only():
    assert run("ir_ont", 1, force_generic=3) >= 1
    assert run("ir_ccs", 2, force_generic=3) > 20


def test_emu_ir_generic_kernel():
    assert run("ir_ccs", 2, force_generic=1, max_groups=40) > 10


@pytest.mark.parametrize("name,n", [("ir_ccs", 6), ("ir_ont", 2), ("ir_clr", 2)])
def test_emu_whole_function_on_captured_segments(name, n):
    """group + band + AffineOneGapAlign fallback + DP + assemble == the reference's IndelRefineAlignment output."""
    recs = po.read_ir_capture(os.path.join(GOLD, name + ".bin"))[:n]
    sb = irgen.pack_segments(recs)
    err, nb, off, blk, info = emu_lib.ir_segments(sb)
    assert err == 0
    assert info[1] > 0
    for s, r in enumerate(recs):
        e = r["blocks_out"]
        assert nb[s] == len(e), (name, s, nb[s], len(e))
        assert (blk[int(off[s]):int(off[s]) + nb[s]] == e).all(), (name, s)


@pytest.mark.parametrize("profile,n", [("ccs", 30), ("ont", 2)])
def test_emu_whole_function_on_synthetic_segments(profile, n):
    """Bench-shaped synthetic segments (CCS ones start at read position 0 with end padding: the band builder's `== 0` quirk).
    The harness also cross-checks the closed-form band kernel against the step-by-step one on every group (error bit 20)."""
    import synth, workload
    genome = synth.gen_ref(2_000_000, 1, 31)[0][1]
    sb = workload.make_segments(profile, n, 8, len(genome), workload.host_genome_fetcher(genome), max_len=6000)
    sb2 = dict(sb); sb2["t_arena"] = sb["t_arena_compact"]; sb2["t_base"] = sb["t_base_compact"]
    err, nb, off, blk, info = emu_lib.ir_segments(sb2)
    assert err == 0
    outs = po.indel_refine_batch_port(sb, sb["t_arena_compact"], sb["t_base_compact"])
    for s, e in enumerate(outs):
        assert nb[s] == len(e) and (blk[int(off[s]):int(off[s]) + nb[s]] == e).all(), (profile, s)
