"""a4 in the mapper worker: the planned CompareLists (lra_b200/csrc/mp_compare.cuh: bounds of every read minimizer found lane-parallel, the
reference's two-ended walk replayed on the bounds, descriptors expanded) against the literal device form (seed_kernels.cuh mm_compare, itself
pinned on the reference header by tests/test_oracle_seed.py / test_gpu_seed.py): same pairs in the same push order, on random sorted lists with
repeated tuples, mixed strand bits, long runs, and lists that end inside a run.  CPU, SIMT emulator, 1 and 32 lanes."""
import ctypes as C
import os
import sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import emu_mp  # noqa: E402

_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
REV = np.uint64(1 << 63)
MASK = np.uint64((1 << 63) - 1)


def bind(L):
    for f in (L.emu_compare_plan, L.emu_compare_literal):
        f.restype = C.c_longlong
        f.argtypes = [_u64p, C.c_int, _u64p, C.c_longlong, C.c_longlong, _i32p, _u32p, C.c_longlong]


def make_lists(rng, nq, nt, universe, p_rev, run_boost):
    def one(n, boost):
        keys = rng.integers(1, universe, size=n).astype(np.uint64)
        if boost:       # long runs of a few tuples
            hot = rng.integers(1, universe, size=3).astype(np.uint64)
            sel = rng.random(n) < boost
            keys[sel] = hot[rng.integers(0, 3, size=int(sel.sum()))]
        strand = (rng.random(n) < p_rev)
        t = keys | np.where(strand, REV, np.uint64(0))
        # sorted by the masked tuple; the order inside a run of equal masked tuples is arbitrary (strand bits mixed)
        o = np.argsort(keys, kind="stable")
        t = t[o]
        # shuffle inside runs
        k = t & MASK
        starts = np.flatnonzero(np.concatenate([[True], k[1:] != k[:-1]]))
        ends = np.concatenate([starts[1:], [n]])
        for a, b in zip(starts, ends):
            if b - a > 1:
                t[a:b] = t[a:b][rng.permutation(b - a)]
        return np.ascontiguousarray(t)
    return one(nq, run_boost), one(nt, run_boost / 2)


@pytest.mark.parametrize("lanes", [1, 32])
def test_planned_compare_lists_equals_literal(lanes):
    L = emu_mp.lib(lanes)
    bind(L)
    rng = np.random.default_rng(99 + lanes)
    cases = 0
    for it in range(260 if lanes == 1 else 60):
        nq = int(rng.integers(1, 400)); nt = int(rng.integers(1, 3000))
        universe = int(rng.choice([8, 40, 300, 5000, 10 ** 6]))
        q, t = make_lists(rng, nq, nt, universe, float(rng.choice([0.0, 0.5, 0.9])), float(rng.choice([0.0, 0.0, 0.3])))
        max_freq = int(rng.choice([1, 2, 5, 150]))
        cap = 4_000_000
        a_q = np.zeros(cap, np.int32); a_t = np.zeros(cap, np.uint32); b_q = np.zeros(cap, np.int32); b_t = np.zeros(cap, np.uint32)
        na = L.emu_compare_plan(q, nq, t, nt, max_freq, a_q, a_t, cap)
        nb = L.emu_compare_literal(q, nq, t, nt, max_freq, b_q, b_t, cap)
        assert na == nb, (it, nq, nt, universe, max_freq, na, nb)
        n = min(na, cap)
        assert (a_q[:n] == b_q[:n]).all() and (a_t[:n] == b_t[:n]).all(), (it, nq, nt, universe, max_freq)
        cases += n > 0
    assert cases > 20
