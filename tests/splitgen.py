"""Seeded cluster sets for the SplitClusters (a9) parity tests: per read a handful of clusters along a diagonal whose boxes overlap in q
and in t (so every cluster is cut by the others' ends), both strands, equal coordinates shared between clusters, tiny clusters, slopes far
from 1 (the q/t comparator then disagrees with itself), anchorfreq values on both sides of the contig thresholds."""
import numpy as np


def make_read(rng, n):
    box = np.zeros((n, 4), np.uint32); strand = np.zeros(n, np.uint8); freq = np.zeros(n, np.float32)
    mq, m_off = [], [0]
    q0 = int(rng.integers(0, 2000)); t0 = int(rng.integers(10000, 1_000_000_000))
    for m in range(n):
        ql = int(rng.choice([4, 30, 200, 900, 3000, 9000])); tl = max(3, int(ql * float(rng.choice([0.5, 0.9, 1.0, 1.0, 1.1, 2.0]))) + int(rng.integers(-2, 3)))
        st = int(rng.random() < 0.3)
        box[m] = [q0, q0 + ql, t0, t0 + tl]; strand[m] = st
        freq[m] = float(rng.choice([1.0, 1.04, 2.5, 3.0, 4.5, 5.0, 7.0]))
        k = int(rng.integers(0, 40))
        a = np.sort(rng.integers(q0, max(q0 + 1, q0 + ql - 17), k)).astype(np.uint32) if k else np.zeros(0, np.uint32)
        mq.append(a); m_off.append(m_off[-1] + k)
        # next cluster: overlapping, abutting (shared coordinate) or after a gap
        mode = rng.random()
        if mode < 0.4:
            q0 = q0 + int(ql * rng.random()); t0 = t0 + int(tl * rng.random())
        elif mode < 0.6:
            q0 = q0 + ql; t0 = t0 + tl
        else:
            q0 = q0 + ql + int(rng.integers(1, 500)); t0 = t0 + tl + int(rng.integers(-200, 3000)); t0 = max(t0, 100)
    return box, strand, freq, (np.concatenate(mq) if mq else np.zeros(0, np.uint32)), np.array(m_off, np.uint64)


def reads(seed, sizes=(0, 1, 2, 3, 5, 8, 14, 30, 60)):
    rng = np.random.default_rng(seed)
    return [make_read(rng, n) for n in sizes]
