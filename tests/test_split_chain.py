"""a11 (low-accuracy pipeline): SPLITChain(UltimateChain) + push_new + MergeSplitchainINS + RemoveSpuriousSplitChain -- the restatement pinned on the
unmodified reference, the kernel through the emulator (CPU) and the C ABI (GPU).  Integer work (one binary64 product, two binary32): bit-exact."""
import numpy as np
import pytest

from oracle import pyoracle as po
import spchaingen

HAVE_REF = po.ref() is not None
KEYS = ["sp_off", "sptc", "sp_lk", "ci_off", "ci", "box", "chrom", "type", "strand", "link"]


def batch(chs):
    c_off = np.zeros(len(chs) + 1, np.uint64); c_off[1:] = np.cumsum([len(c["q"]) for c in chs])
    cat = lambda k, dt: np.concatenate([np.asarray(c[k], dt) for c in chs]) if chs else np.zeros(0, dt)
    link = np.concatenate([np.append(c["link"], 0).astype(np.uint8)[:len(c["q"])] for c in chs]) if chs else np.zeros(0, np.uint8)
    return dict(c_off=c_off, q=cat("q", np.uint32), t=cat("t", np.uint32), len=cat("len", np.int32), strand=cat("strand", np.uint8), cnum=cat("cnum", np.int32), link=link)


def check(o, ac, chs, bypass, which):
    from lra_b200.capi import split_chain_view
    for k, ch in enumerate(chs):
        x = po.split_chain(ch, spchaingen.HDR, 50000, bypass, which=which)
        v = split_chain_view(o, ac["c_off"], k)
        sl = x["sp_lk"].copy(); vl = np.array(v["sp_lk"]).copy()
        for key in KEYS:
            assert np.array_equal(np.asarray(v[key]).reshape(-1), np.asarray(x[key]).reshape(-1)), (k, key)


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
@pytest.mark.parametrize("bypass", [0, 1])
def test_oracle_matches_reference(bypass):
    types, merged = set(), 0
    for ch in spchaingen.chains(3 + bypass, 400):
        a = po.split_chain(ch, spchaingen.HDR, 50000, bypass, which="ref"); b = po.split_chain(ch, spchaingen.HDR, 50000, bypass, which="port")
        for k in KEYS:
            assert np.array_equal(a[k], b[k]), k
        types.update(bytes(b["type"]).decode())
        for s in range(len(b["type"])):
            x = b["sptc"][b["sp_off"][s]:b["sp_off"][s + 1]]
            merged += int(len(x) > 1 and np.abs(np.diff(x)).max() > 1)
    assert types == set("NTI") and merged > 10       # every event type occurs, and MergeSplitchainINS re-joins excursions


def test_known_answers():
    """A forward chain of 6 anchors with a translocated excursion of 3 anchors in the middle: the two sides are merged around it (type of the later
    side), pieces of forward strand come out reversed, the one-anchor tail piece of a second chain is removed."""
    q = np.array([9000, 8900, 8800, 8700, 8600, 8500, 8400, 8300, 8200], np.uint32)
    t = np.array([109000, 108900, 108800, 400700, 400600, 400500, 108400, 108300, 108200], np.uint32)
    ch = dict(q=q, t=t, len=np.full(9, 50, np.int32), strand=np.zeros(9, np.uint8), cnum=np.array([0, 0, 0, 1, 1, 1, 2, 2, 2], np.int32), link=np.zeros(8, np.uint8))
    o = po.split_chain(ch, spchaingen.HDR, 50000, 0)
    assert bytes(o["type"]).decode() == "NT" and o["sp_off"].tolist() == [0, 6, 9]
    assert o["sptc"].tolist() == [8, 7, 6, 2, 1, 0, 5, 4, 3] and o["ci"].tolist() == [0, 1] and o["link"].tolist() == [0]
    assert o["box"].tolist() == [[8200, 9050, 108200, 109050], [8500, 8750, 400500, 400750]]
    o1 = po.split_chain(ch, spchaingen.HDR, 50000, 1)
    assert o1["ci"].tolist() == [0, 2, 1]
    one = dict(q=q[:4], t=np.array([109000, 108900, 108800, 400700], np.uint32), len=np.full(4, 50, np.int32), strand=np.zeros(4, np.uint8), cnum=np.zeros(4, np.int32),
               link=np.zeros(3, np.uint8))
    o2 = po.split_chain(one, spchaingen.HDR, 50000, 0)
    assert o2["sp_off"].tolist() == [0, 3] and bytes(o2["type"]).decode() == "T" and len(o2["link"]) == 0


@pytest.mark.parametrize("bypass", [0, 1])
def test_emu_split_chains(bypass):
    import emu_lib
    chs = spchaingen.chains(11 + bypass, 150)
    ac = batch(chs)
    check(emu_lib.split_chains(ac, spchaingen.HDR, 50000, bypass), ac, chs, bypass, "port")


@pytest.mark.gpu
@pytest.mark.parametrize("bypass", [0, 1])
def test_gpu_split_chains(bypass):
    import lra_b200
    ctx = lra_b200.Context(0)
    chs = spchaingen.chains(21 + bypass, 3000)
    ac = batch(chs)
    check(ctx.split_chains_batch(ac, spchaingen.HDR, 50000, bypass), ac, chs, bypass, "ref" if HAVE_REF else "port")
    e = ctx.split_chains_batch(batch([]), spchaingen.HDR)
    assert e["n_sp"][0] == 0
    ctx.close()
