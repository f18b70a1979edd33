"""a8 (first half): SplitRoughClustersWithGaps over the rough clusters of an anchor list -- the restatement pinned on the unmodified reference, the kernel
through the emulator (CPU) and the C ABI (GPU).  Integer work: bit-exact (anchorfreq is only compared with 10 and carried along)."""
import numpy as np
import pytest

from oracle import pyoracle as po
import roughgen

HAVE_REF = po.ref() is not None
OPT = (17, 1000, 500)       # globalK, RoughClustermaxGap, maxDiag
KEYS = ["start", "end", "box", "strand", "coarse", "freq", "chrom"]


def cases(seed, n):
    rng = np.random.default_rng(seed)
    return [roughgen.rough_list(rng, with_chrom=bool(i & 1)) + (int(rng.choice([1, 2, 3])),) for i in range(n)]


def expect(c, which):
    q, t, rc, mcs = c
    return po.split_rough(q, t, rc, OPT[0], OPT[1], mcs, OPT[2], which=which)


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_lra.so not built (no /root/reference)")
def test_oracle_matches_reference():
    merges = 0
    for c in cases(1, 1500):
        a, b = expect(c, "ref"), expect(c, "port")
        for k in KEYS:
            assert np.array_equal(a[k], b[k]), k
        assert a["smi"] == b["smi"]
        merges += b["n_piece"] - len(b["start"])
    assert merges > 50          # pieces do get re-joined to the previous split cluster


def batch(cs):
    l_off, lr_off = [0], [0]
    cols = {k: [] for k in ["q", "t", "r_start", "r_end", "r_box", "r_strand", "r_freq", "r_chrom"]}
    for q, t, rc, _ in cs:
        l_off.append(l_off[-1] + len(q)); lr_off.append(lr_off[-1] + len(rc["start"]))
        cols["q"].append(q); cols["t"].append(t)
        for k, kk in [("r_start", "start"), ("r_end", "end"), ("r_box", "box"), ("r_strand", "strand"), ("r_freq", "freq"), ("r_chrom", "chrom")]:
            cols[k].append(rc[kk])
    out = {k: np.concatenate(v) for k, v in cols.items()}
    out.update(l_off=np.array(l_off, np.uint64), lr_off=np.array(lr_off, np.uint64))
    return out


def check(o, rl, cs, which, mcs):
    for l, c in enumerate(cs):
        x = po.split_rough(c[0], c[1], c[2], OPT[0], OPT[1], mcs, OPT[2], which=which)
        base = int(rl["l_off"][l]) + int(rl["lr_off"][l]); ns = int(o["n_split"][l]); npc = int(o["n_piece"][l])
        assert ns == len(x["start"]), l
        for k, kk in [("start", "s_start"), ("end", "s_end"), ("box", "s_box"), ("strand", "s_strand"), ("coarse", "s_coarse"), ("freq", "s_freq"), ("chrom", "s_chrom")]:
            assert np.array_equal(o[kk][base:base + ns], x[k]), (l, k)
        smi = [[] for _ in range(ns)]
        for j in range(base, base + npc):
            smi[o["p_cluster"][j]] += list(range(int(o["p_start"][j]), int(o["p_end"][j])))
        assert smi == x["smi"], l


@pytest.mark.parametrize("mcs", [1, 2, 3])
def test_emu_split_rough(mcs):
    import emu_lib
    cs = cases(10 + mcs, 200)
    rl = batch(cs)
    check(emu_lib.split_rough(rl, OPT[0], OPT[1], mcs, OPT[2]), rl, cs, "port", mcs)


@pytest.mark.gpu
@pytest.mark.parametrize("mcs", [1, 2, 3])
def test_gpu_split_rough(mcs):
    import lra_b200
    ctx = lra_b200.Context(0)
    cs = cases(20 + mcs, 3000)
    rl = batch(cs)
    check(ctx.split_rough_batch(rl, OPT[0], OPT[1], mcs, OPT[2]), rl, cs, "ref" if HAVE_REF else "port", mcs)
    ctx.close()
