"""Work directories for the end-to-end and captured-call parity tests (TEST INFRASTRUCTURE): a seeded synthetic reference
and reads (tools/synth.py), the reference's own index (`oracle/_ref/lra_ref index -MODE`), the reference's SAM
(`lra_ref align -MODE -t 1 -p s`) and the capture streams of `oracle/_ref/lra_capture`.  Nothing here reads /root/reference:
the binaries under oracle/_ref/ were built from it by oracle/Makefile and travel with the repository snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "lra_ref")
CAP_BIN = os.path.join(ROOT, "oracle", "_ref", "lra_capture")
MODE = {"ont": "-ONT", "clr": "-CLR", "ccs": "-CCS", "contig": "-CONTIG"}
PROFILE = {"ont": "ont", "clr": "clr", "ccs": "ccs10k", "contig": "contig"}
SEED = {"ont": 2, "clr": 4, "ccs": 11, "contig": 17}


def have_reference_binaries():
    return os.path.exists(REF_BIN) and os.path.exists(CAP_BIN)


def workdir(tmp, preset, n_reads, ref_len=5_000_000, contigs=3, repeats=False, seed=None, sv=False):
    """Creates <tmp>/ref.fa (+ .mms / .gli written by the reference) and <tmp>/reads.fa; returns a dict of paths."""
    tmp = str(tmp)
    os.makedirs(tmp, exist_ok=True)
    ref = synth.gen_ref(ref_len, contigs, 1234)
    if repeats:
        ref = synth.add_repeats(ref)
    w = dict(dir=tmp, preset=preset, ref=os.path.join(tmp, "ref.fa"), reads=os.path.join(tmp, "reads.fa"), n_reads=n_reads)
    synth.write_fasta(w["ref"], ref)
    gen = synth.gen_sv_reads if sv else synth.gen_reads      # sv: deletions, insertions, inversions, translocations, duplications inside the reads
    reads = gen(ref, n_reads, PROFILE[preset], SEED[preset] if seed is None else seed)
    synth.write_fasta(w["reads"], reads, width=1 << 30)
    subprocess.run([REF_BIN, "index", MODE[preset], w["ref"]], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    w["ref_records"] = ref
    w["read_records"] = reads
    return w


def reference_sam(w, threads=1, extra=()):
    out = os.path.join(w["dir"], "ref.sam")
    subprocess.run([REF_BIN, "align", MODE[w["preset"]], w["ref"], w["reads"], "-t", str(threads), "-p", "s", "-o", out] + list(extra), check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return out


def capture_sdp(w):
    out = os.path.join(w["dir"], "sdp.bin")
    env = dict(os.environ, LRA_CAPTURE_SDP=out)
    subprocess.run([CAP_BIN, "align", MODE[w["preset"]], w["ref"], w["reads"], "-t", "1", "-p", "s", "-o", os.path.join(w["dir"], "cap.sam")], check=True, env=env,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return out


def canonical_sam(path):
    """Drop @PG, mask RT:i:<n>, sort the records."""
    import re
    hdr, rec = [], []
    with open(path) as f:
        for line in f:
            if line.startswith("@PG"):
                continue
            if line.startswith("@"):
                hdr.append(line)
            else:
                rec.append(re.sub(r"\tRT:i:\d+", "\tRT:i:0", line))
    return hdr, sorted(rec)
