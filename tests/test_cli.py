"""The C++ host program above the C ABI (lra_b200/cli/lra_b200_cli.cpp, built by lra_b200/build.py): `lra_b200 index -ONT ref.fa` +
`lra_b200 align -ONT ref.fa reads.fa -p s|p|pc|b` against the reference binary on the same files.  SAM: byte-identical after the
canonicalisation of SURVEY 8(c); PAF / BED: identical apart from the RT:i tag.  This is also the test that links the ABI from C++."""
import os
import re
import subprocess
import sys
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import mapgen  # noqa: E402

CLI = os.path.join(ROOT, "lra_b200", "lra_b200")


def test_cli_is_built_and_fails_loudly_without_a_gpu_or_arguments():
    from lra_b200 import build
    build.build()
    assert os.path.exists(CLI)
    p = subprocess.run([CLI], capture_output=True, text=True)
    assert p.returncode != 0 and "usage" in p.stderr
    p = subprocess.run([CLI, "align", "-CCS", "a.fa", "b.fa"], capture_output=True, text=True)
    assert p.returncode != 0 and "Cannot open" in p.stderr


def lines(path, drop_rt=True):
    out = []
    for l in open(path):
        if l.startswith("@PG"):
            continue
        out.append(re.sub(r"\tRT:i:\d+", "", l) if drop_rt else l)
    return sorted(out)


@pytest.mark.gpu
@pytest.mark.skipif(not mapgen.have_reference_binaries(), reason="oracle/_ref binaries not built")
def test_cli_index_and_align_match_the_reference(tmp_path):
    w = mapgen.workdir(tmp_path, "ont", n_reads=200, ref_len=3_000_000, contigs=3, repeats=True)
    ref_out = {}
    for fmt in ("s", "p", "pc", "b"):
        o = str(tmp_path / ("ref." + fmt))
        subprocess.run([mapgen.REF_BIN, "align", "-ONT", w["ref"], w["reads"], "-t", "1", "-p", fmt, "-o", o], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        ref_out[fmt] = lines(o)
    gli = open(w["ref"] + ".gli", "rb").read()
    os.remove(w["ref"] + ".mms"); os.remove(w["ref"] + ".gli")
    subprocess.run([CLI, "index", "-ONT", w["ref"]], check=True)
    assert open(w["ref"] + ".gli", "rb").read() == gli
    for fmt in ("s", "p", "pc", "b"):
        o = str(tmp_path / ("ours." + fmt))
        subprocess.run([CLI, "align", "-ONT", w["ref"], w["reads"], "-t", "4", "-p", fmt, "-o", o, "--batch-bases", "1000000"], check=True)
        ours = lines(o)
        assert len(ours) == len(ref_out[fmt]), fmt
        assert ours == ref_out[fmt], (fmt, [(a, b) for a, b in zip(ours, ref_out[fmt]) if a != b][:2])
    # the printers that read the reference bases: -p a, --printMD
    o_ref = str(tmp_path / "ref.a"); o = str(tmp_path / "ours.a")
    subprocess.run([mapgen.REF_BIN, "align", "-ONT", w["ref"], w["reads"], "-t", "1", "-p", "a", "-o", o_ref], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([CLI, "align", "-ONT", w["ref"], w["reads"], "-p", "a", "-o", o], check=True)
    assert open(o).read() == open(o_ref).read()
    o_ref = str(tmp_path / "ref.md"); o = str(tmp_path / "ours.md")
    subprocess.run([mapgen.REF_BIN, "align", "-ONT", w["ref"], w["reads"], "-t", "1", "-p", "s", "--printMD", "-o", o_ref], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([CLI, "align", "-ONT", w["ref"], w["reads"], "-p", "s", "--printMD", "-o", o], check=True)
    assert lines(o) == lines(o_ref)
    # the high-accuracy preset through the command line (index written by the CLI's own `index -CCS`)
    wc = mapgen.workdir(tmp_path / "ccs", "ccs", n_reads=120, ref_len=3_000_000, contigs=3, repeats=False, sv=True)
    o_ref = str(tmp_path / "ref_ccs.s"); o = str(tmp_path / "ours_ccs.s")
    subprocess.run([mapgen.REF_BIN, "align", "-CCS", wc["ref"], wc["reads"], "-t", "1", "-p", "s", "-o", o_ref], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    os.remove(wc["ref"] + ".mms"); os.remove(wc["ref"] + ".gli")
    subprocess.run([CLI, "index", "-CCS", wc["ref"]], check=True)
    subprocess.run([CLI, "align", "-CCS", wc["ref"], wc["reads"], "-p", "s", "-o", o], check=True)
    a, b = lines(o), lines(o_ref)
    assert len(a) == len(b) and a == b, [(x[:80], y[:80]) for x, y in zip(a, b) if x != y][:2]
    # FASTQ input (lower-case bases, extra words in the header, qualities): plain for the reference, gzip-compressed for the CLI
    import gzip
    rng = __import__("numpy").random.default_rng(3)
    fqp = str(tmp_path / "reads.fq"); fq = fqp + ".gz"
    with open(fqp, "wb") as f1, gzip.open(fq, "wb") as f2:
        for name, seq in w["read_records"]:
            q = (rng.integers(35, 75, size=len(seq)).astype("uint8")).tobytes()
            rec = b"@" + name.encode() + b" extra words\n" + seq.tobytes().lower() + b"\n+\n" + q + b"\n"
            f1.write(rec); f2.write(rec)
    o_ref = str(tmp_path / "ref_fq.s"); o = str(tmp_path / "ours_fq.s")
    subprocess.run([mapgen.REF_BIN, "align", "-ONT", w["ref"], fqp, "-t", "1", "-p", "s", "-o", o_ref], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([CLI, "align", "-ONT", w["ref"], fq, "-p", "s", "-o", o], check=True)
    a, b = lines(o), lines(o_ref)
    assert len(a) == len(b) and a == b, [(x[:80], y[:80]) for x, y in zip(a, b) if x != y][:2]
