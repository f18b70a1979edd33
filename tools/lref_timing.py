#!/usr/bin/env python
"""Per-kernel timing of a12 (lra_b200_lindex_build) and a13 (lra_b200_refine_clusters_batch) on synthetic reads (development aid;
run on the GPU box).  Clusters: one per read, the read's true diagonal sampled every ~60 bases (what seeding + clustering hand over)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import synth, lra_b200

profile = sys.argv[1] if len(sys.argv) > 1 else "ont"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
glen = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000_000
rng = np.random.default_rng(3)
genome = synth.gen_ref(glen, 1, 78)[0][1]
lens = np.minimum(synth.read_lengths(profile, n, rng), glen // 2).astype(np.int64)
starts = rng.integers(0, glen - lens)
reads, mq, mt, m_off, box = [], [], [], [0], []
for i in range(n):
    r = genome[starts[i]:starts[i] + lens[i]]          # error-free copy: the timing does not depend on the errors' exact places
    reads.append(r)
    q = np.arange(0, lens[i] - 17, 60, dtype=np.uint32)
    mq.append(q); mt.append((q + starts[i]).astype(np.uint32)); m_off.append(m_off[-1] + len(q))
    box.append([q[0], q[-1] + 17, q[0] + starts[i], q[-1] + starts[i] + 17])
read_len = lens.astype(np.uint32); read_off = np.zeros(n, np.uint64); read_off[1:] = np.cumsum(lens[:-1])
cl = dict(m_q=np.concatenate(mq), m_t=np.concatenate(mt), m_off=np.array(m_off, np.uint64), box=np.array(box, np.uint32), strand=np.zeros(n, np.uint8),
          read_id=np.arange(n, dtype=np.uint32), hdr_pos=np.array([0, glen], np.uint64), global_k=17, small_k=10, window=100, local_max_freq=15)
ctx = lra_b200.Context(0)
g = ctx.seq_upload(genome); rd = ctx.seq_upload(np.concatenate(reads))
t0 = time.time(); gl = ctx.lindex_build(g, np.array([0], np.uint64), np.array([glen], np.uint32)); dt = time.time() - t0
print("genome LocalIndex: %d windows, %d tuples, wall %.1f ms" % (gl.sizes() + (dt * 1e3,)))
for s in ctx.kernel_stats():
    print("  %-24s %9.3f ms  jobs %9d  algoGB/s %8.2f" % (s["name"], s["ms"], s["jobs"], s["algo_bytes"] / max(s["ms"], 1e-6) / 1e6))
rc = ctx.seq_revcomp(rd, read_off, read_len)
for it in range(2):
    t0 = time.time(); rf = ctx.lindex_build(rd, read_off, read_len); st1 = ctx.kernel_stats(); rr = ctx.lindex_build(rc, read_off, read_len); dt = time.time() - t0
    print("read LocalIndex x2: %d windows each, wall %.1f ms" % (rf.sizes()[0], dt * 1e3))
    for s in st1:
        print("  %-24s %9.3f ms  jobs %9d  algoGB/s %8.2f" % (s["name"], s["ms"], s["jobs"], s["algo_bytes"] / max(s["ms"], 1e-6) / 1e6))
    t0 = time.time(); o = ctx.refine_clusters_batch(gl, rf, rr, cl, anchor_cap=int(lens.sum())); dt = time.time() - t0
    print("refine: %d clusters, %d units, %d tasks, %d anchors, wall %.1f ms (%.0f reads/s)" % (n, o["n_units"], o["n_tasks"], o["n_anchors"], dt * 1e3, n / dt))
    for s in ctx.kernel_stats():
        print("  %-24s %9.3f ms  jobs %9d" % (s["name"], s["ms"], s["jobs"]))
    if it == 0:
        rf.free(); rr.free()
