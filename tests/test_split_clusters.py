"""a9 SplitClusters + DecideSplitClustersValue (SplitClusters.h:63-248): oracle pinned on the reference, kernel logic through the emulator,
the real kernel through the C ABI."""
import numpy as np
import pytest

from oracle import pyoracle as po
import splitgen

needs_ref = pytest.mark.skipif(po.ref() is None, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")


def same(a, b):
    return ((a["split"] == b["split"]).all() and (a["val_cluster"] == b["val_cluster"]).all() and a["sp"].shape == b["sp"].shape and (a["sp"] == b["sp"]).all()
            and (a["sp_val"] == b["sp_val"]).all() and (a["sp_n0"] == b["sp_n0"]).all())


@needs_ref
@pytest.mark.parametrize("contig", [0, 1])
def test_oracle_matches_reference(contig):
    pieces = 0
    for seed in (1, 2, 3, 4, 5):
        for i, (box, strand, freq, mq, m_off) in enumerate(splitgen.reads(seed)):
            a = po.split_clusters(box, strand, freq, contig, mq, m_off, 17, "port"); b = po.split_clusters(box, strand, freq, contig, mq, m_off, 17, "ref")
            assert same(a, b), (contig, seed, i)
            pieces += len(b["sp"])
    assert pieces > 500


def _batch(seeds):
    rs = [r for s in seeds for r in splitgen.reads(s)]
    cl_off = np.zeros(len(rs) + 1, np.uint64); cl_off[1:] = np.cumsum([len(r[1]) for r in rs])
    box = np.concatenate([r[0].reshape(-1, 4) for r in rs]); strand = np.concatenate([r[1] for r in rs]); freq = np.concatenate([r[2] for r in rs])
    mq = np.concatenate([r[3] for r in rs])
    m_off = [0]
    for r in rs:
        m_off += (r[4][1:] + np.uint64(m_off[-1])).tolist()
    return rs, cl_off, box, strand, freq, np.array(m_off, np.uint64), mq


def _check(rs, cl_off, contig, o):
    for r, (box, strand, freq, mq, m_off) in enumerate(rs):
        e = po.split_clusters(box, strand, freq, contig, mq, m_off, 17, "port")
        c0, c1 = int(cl_off[r]), int(cl_off[r + 1]); a, b = int(o["sp_off"][r]), int(o["sp_off"][r + 1])
        assert (o["split"][c0:c1] == e["split"]).all() and (o["val_cluster"][c0:c1] == e["val_cluster"]).all(), r
        assert b - a == len(e["sp"]) and (o["sp"][a:b] == e["sp"]).all() and (o["sp_val"][a:b] == e["sp_val"]).all() and (o["sp_n0"][a:b] == e["sp_n0"]).all(), r


@pytest.mark.parametrize("contig", [0, 1])
def test_emu_split_clusters(contig):
    import emu_lib
    rs, cl_off, box, strand, freq, m_off, mq = _batch([6])
    _check(rs, cl_off, contig, emu_lib.split_clusters(cl_off, box, strand, freq, m_off, mq, contig, 17))


@pytest.mark.gpu
@pytest.mark.parametrize("contig", [0, 1])
def test_gpu_split_clusters(contig):
    import lra_b200
    ctx = lra_b200.Context(0)
    rs, cl_off, box, strand, freq, m_off, mq = _batch([7, 8, 9, 10, 11, 12])
    o = ctx.split_clusters_batch(cl_off, box, strand, freq, m_off, mq, contig, 17)
    _check(rs, cl_off, contig, o)
    assert o["n_pieces"] > 1000
    with pytest.raises(lra_b200.LraB200Error) as ei:
        ctx.split_clusters_batch(cl_off, box, strand, freq, m_off, mq, contig, 17, piece_cap=5)
    assert ei.value.code == lra_b200.capi.EOVERFLOW
    ctx.close()
