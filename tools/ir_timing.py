#!/usr/bin/env python
"""Quick per-kernel timing of lra_b200_indel_refine_batch on synthetic segments (development aid; run on the GPU box)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import synth, workload, lra_b200

profile = sys.argv[1] if len(sys.argv) > 1 else "ont"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
genome = synth.gen_ref(100_000_000, 1, 78)[0][1]
max_len = int(sys.argv[3]) if len(sys.argv) > 3 else None
sb = workload.make_segments(profile, n, 6, len(genome), workload.host_genome_fetcher(genome), max_len=max_len)
print("read_len max", int(sb["read_len"].max()), "mean", float(sb["read_len"].mean()))
ctx = lra_b200.Context(0)
q = ctx.seq_upload(sb["q_arena"][:-16]); t = ctx.seq_upload(genome)
for it in range(3):
    t0 = time.time()
    r = ctx.indel_refine_batch(q, t, sb)
    dt = time.time() - t0
    print("iter", it, "wall ms", round(dt * 1e3, 2), "reads/s", round(n / dt), "cells", r["cells"], "gcups(wall)", round(r["cells"] / dt / 1e9, 2),
          "groups", r["n_dp_groups"], "aog", r["n_aog_jobs"])
for s in ctx.kernel_stats():
    print("  %-34s %9.3f ms  jobs %8d  gcups %8.2f  algoGB/s %8.2f" % (s["name"], s["ms"], s["jobs"], s["cells"] / max(s["ms"], 1e-6) / 1e6,
                                                                 s["algo_bytes"] / max(s["ms"], 1e-6) / 1e6))
