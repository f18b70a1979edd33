#!/usr/bin/env python
"""Times lra_b200_map_batch on synthetic reads against `lra_ref align -t <cores>` on the same inputs (run on a GPU box)."""
import argparse, os, subprocess, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import mapgen, mapemu, lra_b200

ap = argparse.ArgumentParser()
ap.add_argument("--preset", default="ont"); ap.add_argument("--reads", type=int, default=4096); ap.add_argument("--ref-len", type=int, default=5_000_000)
ap.add_argument("--contigs", type=int, default=3); ap.add_argument("--reps", type=int, default=3); ap.add_argument("--no-ref", action="store_true")
a = ap.parse_args()
d = tempfile.mkdtemp(prefix="lra_map_")
t0 = time.time()
w = mapgen.workdir(d, a.preset, n_reads=a.reads, ref_len=a.ref_len, contigs=a.contigs)
print("workdir %.1fs" % (time.time() - t0), flush=True)
inp = mapemu.load_inputs(w)
ctx = lra_b200.Context(0)
mp = lra_b200.Mapper(ctx, inp["opts"], inp["genome"], inp["hdr"], inp["mms"], inp["gli"])
bases = int(inp["read_len"].sum())
for rep in range(a.reps):
    t0 = time.time()
    res = mp.map_batch(inp["reads"], inp["read_off"], inp["read_len"])
    dt = time.time() - t0
    print("rep %d: %.3fs  %.0f reads/s  %.2f Mbp/s  status %s  records %d aligned %.1f Mbp" % (rep, dt, a.reads / dt, bases / dt / 1e6, np.bincount(res["status"][:a.reads], minlength=4), res["n_records"], res["aligned_bases"] / 1e6), flush=True)
for s in ctx.kernel_stats():
    print("  %-28s %9.3f ms  jobs %d" % (s["name"], s["ms"], s["jobs"]))
if not a.no_ref:
    nproc = os.cpu_count()
    t0 = time.time()
    mapgen.reference_sam(w, threads=nproc)
    dt = time.time() - t0
    print("lra_ref align -t %d: %.2fs  %.0f reads/s  %.2f Mbp/s" % (nproc, dt, a.reads / dt, bases / dt / 1e6))
