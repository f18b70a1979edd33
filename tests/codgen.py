"""Seeded anchor lists for the CleanOffDiagonal (a7) parity tests: diagonal runs of every size class the function distinguishes (< 10,
< cleanClustersize, several multiples of it), repeated read tuples at nearby diagonals (tandem repeats: avgfreq from 1.0x to > 4), isolated
off-diagonal noise, runs reaching the end of the list; sorted with the reference's own order (DiagonalSort / AntiDiagonalSort)."""
import numpy as np
from oracle import pyoracle as po

HDR = np.array([0, 1_500_000_000, 2_400_000_000, 3_000_000_000], np.uint64)


def make_list(rng, n_runs, strand, repeat_level):
    q, t = [], []
    for r in range(n_runs):
        size = int(rng.choice([1, 2, 4, 9, 12, 40, 99, 100, 150, 260, 420]))
        q0 = int(rng.integers(0, 40000)); t0 = int(rng.integers(1000, 2_900_000_000))
        step = rng.integers(5, 60, size)
        qq = q0 + np.cumsum(step)
        jitter = rng.integers(-8, 9, size)
        if strand == 0:
            tt = t0 + (qq - q0) + jitter
        else:
            tt = t0 + 60000 - (qq - q0) + jitter
        q.append(qq); t.append(tt)
        # tandem repeat: the same read position (same tuple) again at shifted genome positions
        copies = int(rng.choice([0, 0, 1, 2, 4])) if repeat_level else 0
        for c in range(copies):
            sel = rng.random(size) < float(rng.choice([0.3, 0.8, 1.0]))
            q.append(qq[sel]); t.append(tt[sel] + (c + 1) * int(rng.integers(15, 90)) * (1 if rng.random() < 0.5 else -1))
    noise = int(rng.integers(0, 30))
    q.append(rng.integers(0, 40000, noise)); t.append(rng.integers(1000, 2_900_000_000, noise))
    q = np.concatenate(q).astype(np.uint32); t = np.clip(np.concatenate(t), 0, 2**32 - 20).astype(np.uint32)
    qt = (q.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(3)
    if repeat_level > 1 and len(q) > 10:      # different read positions carrying the same tuple
        d = rng.integers(0, len(q), len(q) // 6); qt[d] = qt[(d + 7) % len(q)]
    sq, st, perm = po.sort_matches(1 if strand else 0, q, t, "port")
    return sq, st, qt[perm]


def lists(seed):
    rng = np.random.default_rng(seed)
    out = [(np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.uint64), 0), ]
    for i in range(14):
        strand = i & 1
        out.append(make_list(rng, int(rng.integers(1, 7)), strand, i % 3) + (strand,))
    out.append((np.array([5], np.uint32), np.array([900], np.uint32), np.array([1], np.uint64), 0))
    return out
