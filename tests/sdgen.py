"""Inputs for the StoreDiagonalClusters test: cleaned anchors of one strand in diagonal order -- runs along a diagonal with small drift, jumps of the diagonal
around opts.maxDiag, runs that are too short / too small, runs made of one repeated read k-mer, per-anchor frequencies (binary32), anchors near contig starts."""
import numpy as np

HDR = np.array([0, 1_000_000, 2_000_000, 3_000_000], np.uint64)


def anchor_list(rng, maxDiag=500):
    strand = int(rng.random() < 0.4)
    Q, T, QT, F = [], [], [], []
    t = int(rng.integers(10_000, 2_900_000)); q = int(rng.integers(0, 2000))
    for _ in range(int(rng.integers(1, 9))):
        k = int(rng.choice([1, 2, 3, 4, 8, 30]))
        rep = rng.random() < 0.15
        tup = int(rng.integers(1, 1 << 34))
        for _ in range(k):
            step = int(rng.integers(0, 60))
            q += step
            t += (step + int(rng.integers(-20, 21))) * (1 if strand == 0 else -1)
            Q.append(q); T.append(max(t, 0)); QT.append(tup if rep else int(rng.integers(1, 1 << 34))); F.append(float(rng.choice([1.0, 1.0, 2.0, 3.5, 7.25])))
        jump = int(rng.choice([maxDiag - 1, maxDiag, maxDiag + 1, 5 * maxDiag, 400_000]))
        t += jump * int(rng.choice([-1, 1]))
        t = min(max(t, 1000), 2_990_000)
        if rng.random() < 0.1:
            t = 1_000_000 * int(rng.integers(1, 3)) + int(rng.integers(-40, 40))      # straddle a contig start
    return np.array(Q, np.uint32), np.array(T, np.uint32), np.array(QT, np.uint64), np.array(F, np.float32), strand
