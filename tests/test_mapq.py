"""a22 SetFromSegAlignment / AlignmentsOrder::Update / SimpleMapQV: oracle pinned on the reference, kernel logic through the emulator, the
real kernel through the C ABI."""
import numpy as np
import pytest

from oracle import pyoracle as po
import mapqgen

needs_ref = pytest.mark.skipif(po.ref() is None, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
KEYS = ["flag", "typeofaln", "issec", "supp", "mapq", "g_issec", "g_value", "g_n0", "g_n1", "g_nm", "order"]
CONFIGS = [(1, 0, 17, 0), (1, 1, 15, 0), (0, 2, 17, 1), (0, 3, 19, 1)]      # (bypassClustering, readType, globalK, Update schedule): ONT, CLR, CCS, CONTIG


@needs_ref
@pytest.mark.parametrize("cfg", CONFIGS)
def test_oracle_matches_reference(cfg):
    bypass, rt, K, sched = cfg
    nonzero = 0
    for seed in (1, 2, 3, 4):
        for i, rd in enumerate(mapqgen.reads(seed, sched)):
            a = po.mapq(rd, bypass, rt, K, "port"); b = po.mapq(rd, bypass, rt, K, "ref")
            for k in KEYS:
                assert (a[k] == b[k]).all(), (cfg, seed, i, k, a[k], b[k])
            nonzero += int((b["mapq"] > 0).sum())
    assert nonzero > 10


def _batch(seeds, sched):
    rds = [rd for s in seeds for rd in mapqgen.reads(s, sched)]
    grp_off = np.zeros(len(rds) + 1, np.int32); grp_off[1:] = np.cumsum([len(rd["seg_off"]) - 1 for rd in rds])
    seg_off = [0]
    for rd in rds:
        seg_off += (rd["seg_off"][1:] + seg_off[-1]).tolist()
    upd_off = np.zeros(len(rds) + 1, np.int32); upd_off[1:] = np.cumsum([len(rd["update_at"]) for rd in rds])
    ag = dict(grp_off=grp_off, seg_off=np.array(seg_off, np.int32), upd_off=upd_off, update_at=np.concatenate([rd["update_at"] for rd in rds]))
    for k in ["value", "n0", "n1", "nm", "nmm", "ndel", "nins", "strand", "flag", "typeofaln", "issec", "supp"]:
        ag[k] = np.concatenate([rd[k] for rd in rds])
    return rds, ag


def _check(rds, ag, cfg, o):
    bypass, rt, K, sched = cfg
    for r, rd in enumerate(rds):
        e = po.mapq(rd, bypass, rt, K, "port")
        g0, g1 = int(ag["grp_off"][r]), int(ag["grp_off"][r + 1]); s0, s1 = int(ag["seg_off"][g0]), int(ag["seg_off"][g1])
        for k in ("flag", "typeofaln", "issec", "supp", "mapq"):
            assert (o[k][s0:s1] == e[k]).all(), (cfg, r, k, o[k][s0:s1], e[k])
        for k in ("g_issec", "g_value", "g_n0", "g_n1", "g_nm", "order"):
            assert (o[k][g0:g1] == e[k]).all(), (cfg, r, k)


@pytest.mark.parametrize("cfg", CONFIGS)
def test_emu_mapq(cfg):
    import ctypes as C
    import emu_lib
    bypass, rt, K, sched = cfg
    rds, ag = _batch([5], sched)
    libm = C.CDLL("libm.so.6"); libm.logf.restype = C.c_float; libm.logf.argtypes = [C.c_float]
    f = np.float32
    logv = np.array([libm.logf(f(v) / f(K)) if v > 3 else 0.0 for v in ag["value"]], np.float32)
    lenpen = np.array([int(f(f(4.343) * f(libm.logf(float(rd["update_at"][-1])))) + f(.499)) for rd in rds], np.int32)
    _check(rds, ag, cfg, emu_lib.mapq(ag, logv, lenpen, bypass, rt))


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CONFIGS)
def test_gpu_mapq(cfg):
    import lra_b200
    bypass, rt, K, sched = cfg
    ctx = lra_b200.Context(0)
    rds, ag = _batch([6, 7, 8, 9, 10, 11, 12, 13], sched)
    o = ctx.mapq_batch(ag, bypass, rt, K)
    _check(rds, ag, cfg, o)
    assert (o["mapq"] > 0).sum() > 20
    ctx.close()
