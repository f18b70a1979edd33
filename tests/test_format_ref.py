"""The printers that read the reference bases (lra_b200_format_records_ref): `-p a` (Alignment::PrintPairwise, Alignment.h:564-589) and the MD:Z: tag of
`--printMD` (AlignmentStringsToMD, Alignment.h:204-245) against `lra_ref align`.  The records come from the emulated pipeline (tests/mapemu.py); the printers are host code."""
import os
import subprocess
import sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import mapgen  # noqa: E402
from test_map_e2e import canon_ours  # noqa: E402


@pytest.mark.parametrize("preset", ["ont", "ccs"])
def test_pairwise_and_md_match_the_reference(preset, tmp_path):
    import mapemu
    from lra_b200 import capi
    w = mapgen.workdir(tmp_path, preset, n_reads=24, ref_len=1_000_000, contigs=2, repeats=False, sv=True)
    inp, mo, res, _ = mapemu.run(w, lanes=1)
    args = (inp["opts"], res, inp["names"], inp["reads"], inp["read_off"], inp["read_len"], inp["contig_names"], inp["genome"], inp["hdr"])
    # -p a
    out = str(tmp_path / "ref.a")
    subprocess.run([mapgen.REF_BIN, "align", mapgen.MODE[preset], w["ref"], w["reads"], "-t", "1", "-p", "a", "-o", out], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ours = capi.format_records_ref(*args, fmt="a")
    ref = open(out).read()
    assert len(ref) > 10000
    assert ours == ref, [(a, b) for a, b in zip(ours.split("\n"), ref.split("\n")) if a != b][:3]
    # -p s --printMD
    out = str(tmp_path / "ref.md.sam")
    subprocess.run([mapgen.REF_BIN, "align", mapgen.MODE[preset], w["ref"], w["reads"], "-t", "1", "-p", "s", "--printMD", "-o", out], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    _, refsam = mapgen.canonical_sam(out)
    ours = canon_ours(capi.format_records_ref(*args, fmt="s", print_md=True))
    assert sum("MD:Z:" in l for l in refsam) >= 20
    assert ours == refsam, [(a[-200:], b[-200:]) for a, b in zip(ours, refsam) if a != b][:2]
