"""End-to-end parity of the MapRead seam (lra_b200_map_batch) with the reference: the SAM text of `oracle/_ref/lra_ref align -MODE -t 1 -p s`
on seeded synthetic reads must be reproduced byte for byte after canonicalisation (drop @PG, mask RT:i, sort records).
CPU: the mapper worker + finalize kernels under the SIMT emulator with the reference standing in for the separately tested a12 / a19 / a21 kernels
(tests/mapemu.py).  GPU: everything through the C ABI."""
import os
import re
import sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import mapgen  # noqa: E402


def canon_ours(text):
    return sorted(re.sub(r"\tRT:i:\d+", "\tRT:i:0", l) + "\n" for l in text.split("\n") if l)


def diff_report(ours, ref, limit=5):
    out = []
    for a, b in zip(ours, ref):
        if a != b:
            fa, fb = a.split("\t"), b.split("\t")
            d = [i for i, (x, y) in enumerate(zip(fa, fb)) if x != y]
            out.append((fa[0], d, [(fa[i][:40], fb[i][:40]) for i in d if i not in (5, 9)][:4]))
            if len(out) >= limit:
                break
    return out


@pytest.mark.parametrize("preset,n_reads,repeats,sv", [("ont", 60, False, False), ("clr", 60, True, False), ("ccs", 80, True, False), ("ccs", 48, False, True),
                                                       ("contig", 6, False, True)])
def test_map_emulated_matches_reference_sam(preset, n_reads, repeats, sv, tmp_path):
    """-CCS / -CONTIG take MapRead_highacc (Map_highacc.h:37-798): fine clusters, split clusters, both high-accuracy SparseDPs, RefineBreakpoint between the
    segments of the reads with structural variants (sv: deletions, insertions, inversions, translocations, duplications)."""
    import mapemu
    w = mapgen.workdir(tmp_path, preset, n_reads=n_reads, ref_len=1_500_000, contigs=3, repeats=repeats, sv=sv)
    _, ref = mapgen.canonical_sam(mapgen.reference_sam(w))
    inp, mo, res, text = mapemu.run(w, lanes=1)
    assert mo["err"] == 0 and (mo["status"] <= 1).all()
    ours = canon_ours(text)
    assert len(ours) == len(ref)
    assert ours == ref, diff_report(ours, ref)


@pytest.mark.parametrize("preset", ["ont", "ccs"])
def test_map_emulated_32_lanes(preset, tmp_path):
    import mapemu
    w = mapgen.workdir(tmp_path, preset, n_reads=2, ref_len=300_000, contigs=2, repeats=False, sv=preset == "ccs")
    _, ref = mapgen.canonical_sam(mapgen.reference_sam(w))
    inp, mo, res, text = mapemu.run(w, lanes=32)
    assert canon_ours(text) == ref


def gpu_sam(w, batch=None):
    import lra_b200
    import mapemu
    inp = mapemu.load_inputs(w)
    ctx = lra_b200.Context(0)
    mp = lra_b200.Mapper(ctx, inp["opts"], inp["genome"], inp["hdr"], inp["mms"], inp["gli"])
    n = len(inp["read_len"]); batch = batch or n
    text = []
    stats = dict(status=[], n_records=0)
    for s in range(0, n, batch):
        e = min(n, s + batch)
        o0 = int(inp["read_off"][s]); o1 = int(inp["read_off"][e - 1]) + int(inp["read_len"][e - 1])
        res = mp.map_batch(inp["reads"][o0:o1], inp["read_off"][s:e] - np.uint64(o0), inp["read_len"][s:e])
        stats["status"].append(res["status"][:e - s].copy()); stats["n_records"] += res["n_records"]
        text.append(lra_b200.format_sam(inp["opts"], res, inp["names"][s:e], inp["reads"][o0:o1], inp["read_off"][s:e] - np.uint64(o0), inp["read_len"][s:e], inp["contig_names"]))
    mp.close(); ctx.close()
    stats["status"] = np.concatenate(stats["status"])
    return "".join(text), stats


@pytest.mark.gpu
@pytest.mark.parametrize("preset,n_reads,repeats,batch", [("ont", 1000, False, None), ("clr", 1000, False, 400), ("ont", 600, True, None), ("clr", 600, True, 250), ("ont", 4000, True, 1500),
                                                          ("ccs", 1000, False, None), ("ccs", 1000, True, 300), ("contig", 24, False, None)])
def test_map_batch_gpu_matches_reference_sam(preset, n_reads, repeats, batch, tmp_path):
    w = mapgen.workdir(tmp_path, preset, n_reads=n_reads, ref_len=5_000_000, contigs=3, repeats=repeats)
    _, ref = mapgen.canonical_sam(mapgen.reference_sam(w))
    text, st = gpu_sam(w, batch)
    assert (st["status"] <= 1).all(), np.bincount(st["status"])
    ours = canon_ours(text)
    assert len(ours) == len(ref)
    assert ours == ref, diff_report(ours, ref)


@pytest.mark.gpu
def test_map_batch_gpu_small_scratch_retries(tmp_path, monkeypatch):
    """Reads whose worker scratch does not fit the first pass (forced here with 2 MB arenas) are mapped again with 8x the scratch: same SAM."""
    w = mapgen.workdir(tmp_path, "ont", n_reads=300, ref_len=5_000_000, contigs=3, repeats=False)
    _, ref = mapgen.canonical_sam(mapgen.reference_sam(w))
    monkeypatch.setenv("LRA_B200_MAP_ARENA_MB", "2")
    text, st = gpu_sam(w, None)
    assert (st["status"] <= 1).all(), np.bincount(st["status"])
    assert canon_ours(text) == ref
    monkeypatch.setenv("LRA_B200_MAP_NO_RETRY", "1")
    text2, st2 = gpu_sam(w, None)
    assert (st2["status"] == 2).any()           # the first pass alone does leave reads behind at this size


def test_map_emulated_structural_variant_reads(tmp_path):
    """Reads with a deletion / insertion / inversion / translocation / duplication inside: split chains of every type, inversion probing,
    supplementary segments and SA:Z tags (emulator, one lane)."""
    import mapemu
    w = mapgen.workdir(tmp_path, "ont", n_reads=30, ref_len=1_500_000, contigs=3, repeats=False, sv=True)
    _, ref = mapgen.canonical_sam(mapgen.reference_sam(w))
    inp, mo, res, text = mapemu.run(w, lanes=1)
    assert mo["err"] == 0 and (mo["status"] <= 1).all()
    ours = canon_ours(text)
    assert sum("SA:Z:" in l for l in ref) >= 5          # the case really has multi-segment alignments
    assert len(ours) == len(ref)
    assert ours == ref, diff_report(ours, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("preset,n_reads", [("ont", 360), ("clr", 360), ("ccs", 360), ("contig", 30)])
def test_map_batch_gpu_structural_variant_reads(preset, n_reads, tmp_path):
    w = mapgen.workdir(tmp_path, preset, n_reads=n_reads, ref_len=5_000_000, contigs=3, repeats=False, sv=True)
    _, ref = mapgen.canonical_sam(mapgen.reference_sam(w))
    text, st = gpu_sam(w, None)
    assert (st["status"] <= 1).all(), np.bincount(st["status"])
    ours = canon_ours(text)
    assert sum("SA:Z:" in l for l in ref) >= n_reads // 8
    assert len(ours) == len(ref)
    assert ours == ref, diff_report(ours, ref)


@pytest.mark.gpu
def test_map_batch_gpu_megabase_contigs(tmp_path, monkeypatch):
    """-CONTIG on 1-2 Mb contigs with seeded structural variants (BASELINE configs[4] is 1-10 Mb): worker arenas of hundreds of MB per warp, few CTAs."""
    monkeypatch.setitem(mapgen.PROFILE, "contig", "contig_long")
    w = mapgen.workdir(tmp_path, "contig", n_reads=6, ref_len=9_000_000, contigs=3, repeats=False, sv=True)
    _, ref = mapgen.canonical_sam(mapgen.reference_sam(w))
    text, st = gpu_sam(w, None)
    assert (st["status"] <= 1).all(), np.bincount(st["status"])
    ours = canon_ours(text)
    assert len(ours) == len(ref)
    assert ours == ref, diff_report(ours, ref)


def edge_case_workdir(tmp_path, preset):
    """Degenerate reads: shorter than k, shorter than a minimizer window, all N, N runs, homopolymer, dinucleotide repeat, an exact copy of the
    reference, a read covering a whole small contig, the same read twice."""
    import synth
    w = mapgen.workdir(tmp_path, preset, n_reads=4, ref_len=600_000, contigs=3, repeats=False)
    ref = w["ref_records"]
    rng = np.random.default_rng(9)
    B = np.frombuffer(b"ACGT", np.uint8)
    c0 = ref[0][1]
    reads = list(w["read_records"])
    reads.append(("short10", c0[1000:1010].copy()))
    reads.append(("short25", c0[2000:2025].copy()))
    reads.append(("short200", c0[3000:3200].copy()))
    reads.append(("allN", np.full(700, ord("N"), np.uint8)))
    x = c0[50000:58000].copy(); x[1000:1400] = ord("N"); x[5000] = ord("N")
    reads.append(("withN", x))
    reads.append(("polyA", np.full(3000, ord("A"), np.uint8)))
    reads.append(("dinuc", np.tile(np.frombuffer(b"AC", np.uint8), 2000)))
    reads.append(("exact", c0[100000:112000].copy()))
    reads.append(("exact_rc", synth.COMP[c0[120000:129000][::-1]]))
    reads.append(("random", B[rng.integers(0, 4, 6000)]))
    reads.append(("dup1", reads[0][1].copy())); reads.append(("dup2", reads[0][1].copy()))
    reads.append(("contig_end", ref[1][1][-7000:].copy()))
    reads.append(("contig_start", ref[2][1][:7000].copy()))
    reads.append(("across_contigs", np.concatenate([ref[0][1][-4000:], ref[1][1][:4000]])))
    synth.write_fasta(w["reads"], reads, width=1 << 30)
    w["read_records"] = reads
    return w


@pytest.mark.parametrize("preset", ["ont", "ccs"])
def test_map_emulated_edge_case_reads(preset, tmp_path):
    import mapemu
    w = edge_case_workdir(tmp_path, preset)
    _, ref = mapgen.canonical_sam(mapgen.reference_sam(w))
    inp, mo, res, text = mapemu.run(w, lanes=1)
    assert mo["err"] == 0 and (mo["status"] <= 1).all()
    ours = canon_ours(text)
    assert len(ours) == len(ref)
    assert ours == ref, diff_report(ours, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["ont", "clr", "ccs", "contig"])
def test_map_batch_gpu_edge_case_reads(preset, tmp_path):
    w = edge_case_workdir(tmp_path, preset)
    _, ref = mapgen.canonical_sam(mapgen.reference_sam(w))
    text, st = gpu_sam(w, None)
    assert (st["status"] <= 1).all(), np.bincount(st["status"])
    ours = canon_ours(text)
    assert len(ours) == len(ref)
    assert ours == ref, diff_report(ours, ref)
