#!/usr/bin/env python
"""Generate tests/golden/* from the UNMODIFIED reference (needs /root/reference and oracle/_ref built).

Golden sets for a18 AffineOneGapAlign (record format = oracle/lra_capture.cpp):
  aog_kat.bin      the query/target pairs (21) of the reference's TestAffineOneGapAlign.cpp:19-68 (the stale
                     driver stores no expected output; outputs here come from the real header via
                     oracle/_ref/libref_lra.so) under three (m,mm,indel,k) settings
  aog_{ccs,ont,clr}.bin   calls captured from `lra align -MODE` on seeded synthetic reads (tools/synth.py)
                     vs a 5 Mb synthetic reference: every two-sided ("alignTop") call + a seeded subsample
  aog_shapes_{ccs,ont,clr}.npy  (qLen,tLen,k) of every captured call: the job-shape table bench.py draws from
Golden sets for a19 IndelRefineAlignment:
  ir_{ccs,ont,clr}.bin    whole segments (read strand, contig window, blocks before / after) captured from the same runs
Run:  python tools/make_golden.py
"""
import os, re, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth
from oracle import pyoracle as po

GOLD = os.path.join(ROOT, "tests", "golden")
CAP = os.path.join(ROOT, "oracle", "_ref", "lra_capture")


def write_records(path, recs):
    with open(path, "wb") as f:
        for r in recs:
            b = np.asarray(r["blocks"], dtype=np.uint32).reshape(-1, 3)
            f.write(np.array([len(r["q"]), len(r["t"]), r["m"], r["mm"], r["indel"], r["k"], r["score"], len(b)],
                             dtype=np.int32).tobytes())
            f.write(bytes(r["q"])); f.write(bytes(r["t"])); f.write(b.tobytes())


def write_ir_records(path, recs):
    with open(path, "wb") as f:
        for r in recs:
            f.write(np.array([len(r["read"]), r["contig_len"], r["k"], r["match"], r["mismatch"], r["indel"], r["end_align"],
                              len(r["blocks_in"]), len(r["blocks_out"]), r["t_win_off"], len(r["twin"]), r["strand"]], dtype=np.int32).tobytes())
            f.write(bytes(r["read"])); f.write(bytes(r["twin"]))
            f.write(np.asarray(r["blocks_in"], np.uint32).tobytes()); f.write(np.asarray(r["blocks_out"], np.uint32).tobytes())


def kat20():
    src = open("/root/reference/TestAffineOneGapAlign.cpp").read()
    body = src[src.index("int main"):src.index("/*\n\tstring target")]
    pairs = re.findall(r'Test\(\s*"([A-Za-z]*)"\s*,\s*"([A-Za-z]*)"\s*\)', body)
    assert len(pairs) >= 20, len(pairs)
    recs = []
    for (m, mm, indel, k) in [(4, -4, -3, 15), (4, -3, -4, 15), (4, -1, -2, 30), (4, -3, -4, 7)]:
        for q, t in pairs:
            q, t = q.encode(), t.encode()
            s, b = po.aog_ref(q, t, m, mm, indel, k)
            recs.append(dict(q=q, t=t, m=m, mm=mm, indel=indel, k=k, score=s, blocks=b))
    write_records(os.path.join(GOLD, "aog_kat.bin"), recs)
    print("aog_kat:", len(recs))


def captured():
    tmp = tempfile.mkdtemp(prefix="lra_gold_")
    ref = synth.gen_ref(5_000_000, 1, 1234)
    synth.write_fasta(os.path.join(tmp, "ref.fa"), ref)
    cfg = {"ccs": ("-CCS", "ccs10k", 11, 200), "ont": ("-ONT", "ont", 2, 60), "clr": ("-CLR", "clr", 4, 60)}
    for name, (mode, prof, seed, n) in cfg.items():
        d = os.path.join(tmp, name); os.makedirs(d)
        os.symlink(os.path.join(tmp, "ref.fa"), os.path.join(d, "ref.fa"))
        synth.write_fasta(os.path.join(d, "reads.fa"), synth.gen_reads(ref, n, prof, seed), width=1 << 30)
        subprocess.run([CAP, "index", mode, "ref.fa"], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        env = dict(os.environ, LRA_CAPTURE_AOG=os.path.join(d, "aog.cap"), LRA_CAPTURE_IR=os.path.join(d, "ir.cap"))
        subprocess.run([CAP, "align", mode, "ref.fa", "reads.fa", "-t", "1", "-p", "s", "-o", "out.sam"], cwd=d, env=env,
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        recs = po.read_aog_capture(os.path.join(d, "aog.cap"))
        ql = np.array([len(r["q"]) for r in recs]); tl = np.array([len(r["t"]) for r in recs])
        kk = np.array([r["k"] for r in recs])
        np.save(os.path.join(GOLD, "aog_shapes_%s.npy" % name), np.stack([ql, tl, kk], 1).astype(np.int32))
        dg = np.maximum(1, np.minimum(ql, tl)); k2 = np.minimum(dg, kk)
        two = dg + 2 * k2 < np.maximum(ql, tl)
        rng = np.random.default_rng(7)
        pick = set(np.flatnonzero(two).tolist())
        pick |= set(rng.choice(len(recs), size=min(len(recs), 1500), replace=False).tolist())
        pick |= set(np.argsort(-(ql + tl))[:40].tolist())          # the largest jobs
        sel = [recs[i] for i in sorted(pick)]
        write_records(os.path.join(GOLD, "aog_%s.bin" % name), sel)
        print("aog_%s: %d of %d calls (%d two-sided)" % (name, len(sel), len(recs), int(two.sum())))
        # a19 IndelRefineAlignment: whole-segment before/after records (record format: oracle/lra_capture.cpp)
        irs = po.read_ir_capture(os.path.join(d, "ir.cap"))
        keep = {"ccs": 24, "ont": 10, "clr": 10}[name]
        pick = sorted(np.random.default_rng(9).choice(len(irs), size=min(keep, len(irs)), replace=False).tolist())
        write_ir_records(os.path.join(GOLD, "ir_%s.bin" % name), [irs[i] for i in pick])
        print("ir_%s: %d of %d segments" % (name, len(pick), len(irs)))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    kat20()
    captured()
