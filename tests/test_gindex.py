"""f2 (global half): the global minimizer index built on the GPU (lra_b200_gindex_build) against <ref>.mms written by the reference binary
(`lra_ref index -MODE`, StoreIndex MMIndex.h:286-399): same header, same records -- the same (tuple, position) set, sorted by the masked
tuple; inside a run of equal tuples the reference keeps introsort's order, the GPU builder position order (include/lra_b200.h), so runs are
compared as sets.  And the reference maps the same reads to the same SAM with either index."""
import os
import subprocess
import sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import mapgen  # noqa: E402
import lra_b200  # noqa: E402
from lra_b200 import capi  # noqa: E402

B = np.frombuffer(b"ACGT", np.uint8)
MASK = np.uint64(0x7FFFFFFFFFFFFFFF)


def tricky_contigs(seed):
    rng = np.random.default_rng(seed)
    out = []
    for i, L in enumerate((300001, 40, 26, 25, 123457, 2047, 16)):
        s = B[rng.integers(0, 4, L)].copy()
        if L == 123457:
            s[5000:5100] = ord("N"); s[40000:52000] = np.tile(B[rng.integers(0, 4, 3)], 4000); s[90000:96000] = ord("A"); s[-1] = ord("N")
            s[70000] = ord("N"); s[70020] = ord("N"); s[-30] = ord("N")
        if L == 300001:
            s[1000:9000] = np.tile(B[rng.integers(0, 4, 2)], 4000)          # long tandem repeat: ties in every window
            s[200000:203000] = s[100000:103000]                              # repeated segment: equal tuples at two places
            s[250000:251500] = synth_rc(s[100000:101500])
        out.append(("chr%d" % (i + 1), s))
    return out


def synth_rc(s):
    comp = np.zeros(256, np.uint8)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    return comp[s[::-1]]


def canon(t, pos):
    """records sorted by (masked tuple, position, strand bit): the set, independent of the order inside runs of equal tuples"""
    o = np.lexsort((pos, t & MASK))
    return t[o], pos[o]


def write_fa(path, contigs):
    import synth
    synth.write_fasta(path, contigs)


@pytest.mark.gpu
@pytest.mark.skipif(not mapgen.have_reference_binaries(), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("preset", ["ont", "clr", "contig"])
def test_gindex_matches_reference_mms(preset, tmp_path):
    contigs = tricky_contigs(7)
    fa = str(tmp_path / "g.fa")
    write_fa(fa, contigs)
    subprocess.run([mapgen.REF_BIN, "index", mapgen.MODE[preset], fa], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    want = capi.read_mms(fa + ".mms")
    ctx = lra_b200.Context(0)
    lens = np.array([len(s) for _, s in contigs], np.uint32)
    start = np.zeros(len(contigs), np.uint64); start[1:] = np.cumsum(lens[:-1].astype(np.uint64))
    arena = ctx.seq_upload(np.concatenate([s for _, s in contigs]))
    t, pos = ctx.gindex_build(arena, start, lens, **capi.INDEX_PRESET[preset])
    arena.free(); ctx.close()
    assert len(t) == len(want["t"])
    assert ((t & MASK) == (want["t"] & MASK)).all()                      # sorted by the masked tuple, same multiset of tuples
    a, b = canon(t, pos), canon(want["t"], want["pos"])
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


@pytest.mark.gpu
@pytest.mark.skipif(not mapgen.have_reference_binaries(), reason="oracle/_ref binaries not built")
def test_reference_maps_identically_with_gpu_built_index(tmp_path):
    """build_index_files writes <ref>.mms / <ref>.gli the reference loads; its SAM on 150 ONT reads (repeat-enriched reference) is the same
    as with its own index files."""
    w = mapgen.workdir(tmp_path / "a", "ont", n_reads=150, ref_len=3_000_000, contigs=3, repeats=True)
    sam_ref = open(mapgen.reference_sam(w)).read().split("\n")
    gli_ref = open(w["ref"] + ".gli", "rb").read()
    os.remove(w["ref"] + ".mms"); os.remove(w["ref"] + ".gli")
    info = lra_b200.build_index_files(w["ref"], w["ref_records"], "ont")
    assert info["n_mms"] > 0 and open(w["ref"] + ".gli", "rb").read() == gli_ref
    sam_ours = open(mapgen.reference_sam(w)).read().split("\n")
    strip = lambda ls: sorted(l for l in ls if not l.startswith("@PG"))
    import re
    mask = lambda ls: [re.sub(r"\tRT:i:\d+", "", l) for l in ls]
    assert mask(strip(sam_ours)) == mask(strip(sam_ref))
