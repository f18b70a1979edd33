"""a20 RefineBreakpoint: oracle pinned on the reference, kernel logic through the emulator, the real kernels through the C ABI."""
import numpy as np
import pytest

from oracle import pyoracle as po
import bpgen

needs_ref = pytest.mark.skipif(po.ref() is None, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")


def strands(c):
    return (c["read"] if c["lstrand"] == 0 else c["read_rc"]), (c["read"] if c["rstrand"] == 0 else c["read_rc"])


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_matches_reference(seed):
    refined = 0
    for i, c in enumerate(bpgen.cases(seed)):
        lr, rr = strands(c)
        a = po.refine_breakpoint_ref(lr, rr, c["L"], c["G1"], c["G2"], c["lblocks"], c["lstrand"], c["rblocks"], c["rstrand"])
        r, bl, br, _ = po.refine_breakpoint_port(lr, rr, c["L"], c["G1"], c["G2"], c["lblocks"], c["lstrand"], c["rblocks"], c["rstrand"])
        assert len(a[0]) == len(bl) and (a[0] == bl).all() and len(a[1]) == len(br) and (a[1] == br).all(), (seed, i, c["span"], c["lstrand"], c["rstrand"])
        refined += r
    assert refined >= 20


def _check(cs, o):
    from oracle import pyoracle as po
    for p, c in enumerate(cs):
        lr, rr = strands(c)
        r, bl, br, (mode, n_out, bound, out) = po.refine_breakpoint_port(lr, rr, c["L"], c["G1"], c["G2"], c["lblocks"], c["lstrand"], c["rblocks"], c["rstrand"])
        assert o["refined"][p] == r, p
        assert (o["mode"][p] == mode).all() and (o["n_out"][p] == n_out).all(), (p, o["mode"][p], mode, o["n_out"][p], n_out)
        if r:
            assert (o["bound"][p].reshape(-1) == bound).all(), p
            for s in (0, 1):
                assert (o["out"][p, s, :n_out[s]] == out[s, :n_out[s]]).all(), (p, s)


def test_emu_refine_breakpoint():
    import emu_lib
    rng = np.random.default_rng(5)
    cs = [bpgen.make_case(rng, span=s, lstrand=i & 1, rstrand=(i >> 1) & 1, edge=(i // 4) % 3) for i, s in enumerate([1, 2, 9, 33, 40, 57, 64, 20, 31, 45, 3, 50])]
    fwd, rcs, gen, bp = bpgen.pack(cs)
    _check(cs, emu_lib.refine_breakpoint(fwd, rcs, gen, bp))


@pytest.mark.gpu
def test_gpu_refine_breakpoint():
    import lra_b200
    ctx = lra_b200.Context(0)
    cs = bpgen.cases(7, n=48) + bpgen.cases(8, n=24)
    fwd, rcs, gen, bp = bpgen.pack(cs)
    f = ctx.seq_upload(fwd[:-16]); r = ctx.seq_upload(rcs[:-16]); g = ctx.seq_upload(gen[:-16])
    o = ctx.refine_breakpoint_batch(f, r, g, bp)
    _check(cs, o)
    assert o["refined"].sum() >= 60 and o["n_out"].sum() > 50
    for x in (f, r, g):
        x.free()
    ctx.close()
