"""a10 SparseDP family (SparseDP.h:2139, :2287, SparseDP_Forward.h:312) against the reference itself: every SparseDP /
SparseDP_ForwardOnly call of real `lra align -ONT` / `-CLR` runs, captured by oracle/_ref/lra_capture ($LRA_CAPTURE_SDP),
must come back bit-identical (anchors of every chain, link bits, float value bits, chain bounds, number of chains).
CPU: the kernel source under the SIMT emulator (1 lane: all calls; 32 lanes: a subset).  GPU: through the C ABI."""
import ctypes as C
import os
import subprocess
import sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth  # noqa: E402
import sdpgen  # noqa: E402
import mapgen  # noqa: E402

PRESET = {"ont": dict(pwl=(7.0, 10.0, 1.5, 1500, 3000), alnthres=0.65, NumAln=2), "clr": dict(pwl=(7.0, 10.0, 1.5, 1500, 3000), alnthres=0.5, NumAln=2),
          "ccs": dict(pwl=(4.0, 15.0, 1.5, 2000, 3000), alnthres=0.7, NumAln=2)}


def ref_pwl(p):
    from oracle import pyoracle as po
    L = po.ref()
    stops = np.zeros(25, np.int64); slope = np.zeros(25, np.float32); inter = np.zeros(25, np.float32)
    L.ref_init_pwl.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_init_pwl(p[0], p[1], p[2], p[3], p[4], stops.ctypes.data, slope.ctypes.data, inter.ctypes.data)
    return stops, slope, inter, p[3], p[4]


def test_init_pwl_matches_reference():
    import lra_b200
    for p in [(7.0, 10.0, 1.5, 1500, 3000), (4.0, 15.0, 1.5, 2000, 3000), (4.0, 20.0, 1.5, 3000, 5000), (2.0, 10.0, 2.0, 1500, 3000)]:
        a = ref_pwl(p); b = lra_b200.init_pwl(*p)
        assert (a[0] == b[0]).all() and (a[1].view(np.uint32) == b[1].view(np.uint32)).all() and (a[2].view(np.uint32) == b[2].view(np.uint32)).all()


@pytest.mark.parametrize("preset,n_reads", [("ont", 60), ("clr", 60)])
def test_sdp_emulated_on_captured_calls(preset, n_reads, tmp_path):
    import emu_mp
    w = mapgen.workdir(tmp_path, preset, n_reads=n_reads, ref_len=600_000, contigs=2, repeats=True)
    recs = sdpgen.parse_capture(mapgen.capture_sdp(w))
    assert len(recs) >= 2 * n_reads - 10
    P = PRESET[preset]
    pb = sdpgen.pack(recs)
    out = emu_mp.sdp_batch(pb, ref_pwl(P["pwl"]), P["alnthres"], P["NumAln"], lanes=1)
    assert out["err"] == 0
    assert sdpgen.compare(recs, pb, out, 2) == []
    sub = recs[:6]
    pb = sdpgen.pack(sub)
    out = emu_mp.sdp_batch(pb, ref_pwl(P["pwl"]), P["alnthres"], P["NumAln"], lanes=32)
    assert out["err"] == 0
    assert sdpgen.compare(sub, pb, out, 2) == []


@pytest.mark.gpu
@pytest.mark.parametrize("preset,n_reads,repeats", [("ont", 2500, False), ("clr", 2500, False), ("ont", 600, True), ("clr", 600, True)])
def test_sdp_gpu_on_captured_calls(preset, n_reads, repeats, tmp_path):
    import lra_b200
    w = mapgen.workdir(tmp_path, preset, n_reads=n_reads, ref_len=5_000_000, contigs=3, repeats=repeats)
    recs = sdpgen.parse_capture(mapgen.capture_sdp(w))
    assert len(recs) >= 2 * n_reads - 20
    P = PRESET[preset]
    ctx = lra_b200.Context(0)
    pwl = lra_b200.init_pwl(*P["pwl"])
    bad = []
    for s in range(0, len(recs), 2048):
        sub = recs[s:s + 2048]
        pb = sdpgen.pack(sub)
        out = ctx.sdp_batch(pb, pwl, P["alnthres"], P["NumAln"])
        bad += [(s + k, m) for k, m in sdpgen.compare(sub, pb, out, 2)]
    ctx.close()
    assert bad == [], bad[:10]


def test_sdp_emulated_on_captured_highacc_calls(tmp_path):
    """SparseDP.h:1766 (the second SparseDP of the high-accuracy pipeline, over the Cluster_SameDiag anchors of a split chain) on the calls of a
    real `lra align -CCS` run, and SparseDP.h:1956 (the first one, over the split clusters, with DecidePrimaryChains :1586): chains, link bits,
    float value bits, boxes, NumOfAnchors0 bit-identical."""
    import emu_mp
    w = mapgen.workdir(tmp_path, "ccs", n_reads=60, ref_len=600_000, contigs=2, repeats=True)
    recs = [r for r in sdpgen.parse_capture(mapgen.capture_sdp(w)) if r["kind"] in (3, 4)]
    assert sum(r["kind"] == 3 for r in recs) >= 50 and sum(r["kind"] == 4 for r in recs) >= 50
    P = PRESET["ccs"]
    pb = sdpgen.pack(recs)
    out = emu_mp.sdp_batch(pb, ref_pwl(P["pwl"]), P["alnthres"], P["NumAln"], lanes=1)
    assert out["err"] == 0
    assert sdpgen.compare(recs, pb, out, 2) == []
    sub = recs[:4]
    pb = sdpgen.pack(sub)
    out = emu_mp.sdp_batch(pb, ref_pwl(P["pwl"]), P["alnthres"], P["NumAln"], lanes=32)
    assert sdpgen.compare(sub, pb, out, 2) == []


@pytest.mark.gpu
def test_sdp_gpu_on_captured_highacc_calls(tmp_path):
    import lra_b200
    w = mapgen.workdir(tmp_path, "ccs", n_reads=1500, ref_len=5_000_000, contigs=3, repeats=True)
    recs = [r for r in sdpgen.parse_capture(mapgen.capture_sdp(w)) if r["kind"] in (3, 4)]
    assert sum(r["kind"] == 3 for r in recs) >= 1400 and sum(r["kind"] == 4 for r in recs) >= 1400
    P = PRESET["ccs"]
    ctx = lra_b200.Context(0)
    pwl = lra_b200.init_pwl(*P["pwl"])
    bad = []
    for s in range(0, len(recs), 2048):
        sub = recs[s:s + 2048]
        pb = sdpgen.pack(sub)
        out = ctx.sdp_batch(pb, pwl, P["alnthres"], P["NumAln"])
        bad += [(s + k, m) for k, m in sdpgen.compare(sub, pb, out, 2)]
    ctx.close()
    assert bad == [], bad[:10]
