"""a7 CleanOffDiagonal / SecondRoundCleanOffDiagonal / AVGfreq (Clustering.h:549-868): oracle pinned on the reference, kernel logic through
the emulator, the real kernel through the C ABI."""
import numpy as np
import pytest

from oracle import pyoracle as po
import codgen

needs_ref = pytest.mark.skipif(po.ref() is None, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")


@needs_ref
@pytest.mark.parametrize("preset", ["ccs", "clr", "ont", "noextract"])
def test_oracle_matches_reference(preset):
    opts = po.COD_PRESETS[preset]
    kept = removed = ncl = 0
    for seed in (1, 2, 3):
        for i, (q, t, qt, strand) in enumerate(codgen.lists(seed)):
            a = po.clean_off_diagonal(q, t, qt, strand, opts, codgen.HDR, "port"); b = po.clean_off_diagonal(q, t, qt, strand, opts, codgen.HDR, "ref")
            assert len(a["kq"]) == len(b["kq"]) and (a["kq"] == b["kq"]).all() and (a["kt"] == b["kt"]).all(), (preset, seed, i)
            assert (a["kfreq"] == b["kfreq"]).all(), (preset, seed, i)
            assert a["cl"].shape == b["cl"].shape and (a["cl"] == b["cl"]).all() and (a["cl_freq"] == b["cl_freq"]).all(), (preset, seed, i)
            kept += len(b["kq"]); removed += len(q) - len(b["kq"]); ncl += len(b["cl"])
    assert kept > 500 and removed > 200 and (ncl > 10 or preset == "noextract")


def _batch(seeds):
    ls = [l for s in seeds for l in codgen.lists(s)]
    off = np.zeros(len(ls) + 1, np.uint64); off[1:] = np.cumsum([len(l[0]) for l in ls])
    return ls, np.concatenate([l[0] for l in ls]), np.concatenate([l[1] for l in ls]), np.concatenate([l[2] for l in ls]), off, np.array([l[3] for l in ls], np.uint8)


def _check(ls, off, opts, o):
    for s, (q, t, qt, strand) in enumerate(ls):
        a, b = int(off[s]), int(off[s + 1])
        e = po.clean_off_diagonal(q, t, qt, strand, opts, codgen.HDR, "port")
        assert (o["keep"][a:b] == e["keep"]).all(), s
        k = e["keep"] == 1
        assert (o["freq"][a:b][k] == e["freq"][k]).all() and (o["cnt"][a:b][k] == e["cnt"][k]).all(), s
        n = len(e["cl"])
        assert o["n_cl"][s] == n, (s, o["n_cl"][s], n)
        assert (o["cl"][a:a + n] == e["cl"]).all() and (o["cl_freq"][a:a + n] == e["cl_freq"]).all(), s


@pytest.mark.parametrize("preset", ["ccs", "ont", "noextract"])
def test_emu_clean_off_diagonal(preset):
    import emu_lib
    opts = po.COD_PRESETS[preset]
    ls, q, t, qt, off, strand = _batch([4])
    _check(ls, off, opts, emu_lib.clean_off_diagonal(q, t, qt, off, strand, [opts[k] for k in po.COD_FIELDS], codgen.HDR))


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["ccs", "clr", "ont", "noextract"])
def test_gpu_clean_off_diagonal(preset):
    import lra_b200
    ctx = lra_b200.Context(0)
    opts = po.COD_PRESETS[preset]
    ls, q, t, qt, off, strand = _batch([5, 6, 7, 8, 9, 10])
    o = ctx.clean_off_diagonal_batch(q, t, qt, off, strand, opts, codgen.HDR)
    _check(ls, off, opts, o)
    assert o["keep"].sum() > 1000 and (o["keep"] == 0).sum() > 500
    ctx.close()
