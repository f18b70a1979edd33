"""Kernel-logic parity on the CPU for a1-a5 (lra_b200/csrc/seed_kernels.cuh through the SIMT emulator) against the oracle."""
import numpy as np
import pytest

from oracle import pyoracle as po
import emu_lib
import seedgen


def pack_expected_rc(arena, read_off, read_len):
    n = len(arena) - 16
    code = np.full(n, 4, np.uint8)
    for r in range(len(read_off)):
        o, L = int(read_off[r]), int(read_len[r])
        rc = seedgen.COMP[arena[o:o + L][::-1]]
        lut = np.full(256, 4, np.uint8); lut[[65, 67, 71, 84]] = [0, 1, 2, 3]
        code[o:o + L] = lut[rc]
    return code


@pytest.mark.parametrize("seed,k,w,mf", [(1, 17, 10, 150), (2, 15, 10, 2), (3, 17, 20, 150)])
def test_emu_seed_batch(seed, k, w, mf):
    case = seedgen.make_case(seed, glen=100000, n_reads=14, k=k, w=w)
    exp = seedgen.expected(case, mf)
    o = emu_lib.seed_batch(case["arena"], case["read_off"], case["read_len"], case["genome"], case["idx_t"], case["idx_pos"], k, w, mf)
    assert o["n"] == sum(len(e[0]) for e in exp) and o["n"] > 100
    for r, e in enumerate(exp):
        a, b = int(o["match_off"][r]), int(o["match_off"][r + 1])
        assert b - a == len(e[0]), r
        for key, ev in zip(["q_t", "q_pos", "t_t", "t_pos", "strand"], e):
            assert (o[key][a:b] == ev).all(), (r, key)
        # the sorted minimizer list of the read (a2 + a3)
        mt, mp = po.sort_minimizers(*po.store_minimizers(case["reads"][r], k, w))
        ro = int(case["read_off"][r])
        assert o["n_mm"][r] == len(mt)
        assert (o["mm_t"][ro:ro + len(mt)] == mt).all() and (o["mm_pos"][ro:ro + len(mt)] == mp).all()


def test_emu_revcomp():
    case = seedgen.make_case(4, glen=100000, n_reads=9)
    b2, nm = emu_lib.seq_revcomp(case["arena"], case["read_off"], case["read_len"])
    code = pack_expected_rc(case["arena"], case["read_off"], case["read_len"])
    n = len(code)
    got2 = (np.repeat(b2, 16)[:n] >> (2 * (np.arange(n) % 16)).astype(np.uint32)) & 3
    gotn = (np.repeat(nm, 32)[:n] >> (np.arange(n) % 32).astype(np.uint32)) & 1
    got = np.where(gotn == 1, 4, got2)
    assert (got == code).all()
