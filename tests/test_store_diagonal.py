"""StoreDiagonalClusters (the cluster builder of CleanMatches without ExtractDiagonalFromClean): restatement pinned on the unmodified reference, kernel through
the emulator (CPU) and the C ABI (GPU).  anchorfreq is a binary32 running sum / count: compared bit for bit."""
import numpy as np
import pytest

from oracle import pyoracle as po
import sdgen

HAVE_REF = po.ref() is not None
KEYS = ["start", "end", "box", "freq", "chrom"]


def cases(seed, n):
    rng = np.random.default_rng(seed)
    return [sdgen.anchor_list(rng) for _ in range(n)]


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
def test_oracle_matches_reference():
    rng = np.random.default_rng(1)
    nc = 0
    for it, (q, t, qt, f, st) in enumerate(cases(2, 1500)):
        args = (17, 500, int(rng.choice([1, 2, 3])), int(rng.choice([17, 50, 200])), it & 1)
        a = po.store_diagonal(q, t, qt, f, st, sdgen.HDR, *args, which="ref"); b = po.store_diagonal(q, t, qt, f, st, sdgen.HDR, *args)
        for k in KEYS:
            assert a[k].tobytes() == b[k].tobytes(), (it, k)
        nc += len(b["start"])
    assert nc > 1000


def batch(cs):
    l_off = np.zeros(len(cs) + 1, np.uint64); l_off[1:] = np.cumsum([len(c[0]) for c in cs])
    return dict(l_off=l_off, q=np.concatenate([c[0] for c in cs]), t=np.concatenate([c[1] for c in cs]), qt=np.concatenate([c[2] for c in cs]),
                freq=np.concatenate([c[3] for c in cs]), strand=np.array([c[4] for c in cs], np.uint8))


def check(o, cl, cs, args, which):
    for l, (q, t, qt, f, st) in enumerate(cs):
        x = po.store_diagonal(q, t, qt, f, st, sdgen.HDR, *args, which=which)
        a = int(cl["l_off"][l]); n = int(o["n_cl"][l])
        assert n == len(x["start"]), l
        for k, kk in [("start", "c_start"), ("end", "c_end"), ("box", "c_box"), ("freq", "c_freq"), ("chrom", "c_chrom")]:
            assert np.ascontiguousarray(o[kk][a:a + n]).tobytes() == x[k].tobytes(), (l, k)


@pytest.mark.parametrize("args", [(17, 500, 2, 50, 0), (17, 500, 1, 17, 1), (15, 500, 3, 200, 1)])
def test_emu_store_diagonal(args):
    import emu_lib
    cs = cases(5, 300)
    cl = batch(cs)
    check(emu_lib.store_diagonal(cl, sdgen.HDR, *args), cl, cs, args, "port")


@pytest.mark.gpu
@pytest.mark.parametrize("args", [(17, 500, 2, 50, 0), (17, 500, 1, 17, 1)])
def test_gpu_store_diagonal(args):
    import lra_b200
    ctx = lra_b200.Context(0)
    cs = cases(7, 4000)
    cl = batch(cs)
    check(ctx.store_diagonal_batch(cl, sdgen.HDR, *args), cl, cs, args, "ref" if HAVE_REF else "port")
    ctx.close()
