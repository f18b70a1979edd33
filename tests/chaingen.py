"""Seeded fragment sets for the GlobalChain (a24) parity tests."""
import numpy as np

KAT = np.array([[0, 0, 10, 10], [20, 20, 30, 30], [40, 40, 50, 50], [60, 60, 70, 70], [80, 80, 90, 90], [100, 100, 110, 110], [120, 120, 130, 130],
                [140, 140, 150, 150], [81, 31, 91, 41]], np.int32)            # TestGlobalChain.cpp:29-38


def problems(seed, sizes=(1, 2, 3, 9, 40, 200, 1500)):
    """Fragment sets: a noisy diagonal with off-diagonal decoys, touching fragments (an end point equal to another fragment's start point:
    the tie std::sort leaves in an order of its own), duplicates, zero-length fragments, negative coordinates (unsigned keys)."""
    rng = np.random.default_rng(seed)
    out = []
    for n in sizes:
        x = np.sort(rng.integers(0, 40 * n + 10, n)).astype(np.int64)
        ln = rng.integers(0, 30, n)
        y = x + rng.integers(-15, 15, n)
        decoy = rng.random(n) < 0.25
        y[decoy] = rng.integers(0, 40 * n + 10, int(decoy.sum()))
        f = np.stack([x, y, x + ln, y + ln], 1)
        if n >= 9:
            for i in range(1, n, 4):           # fragment i starts exactly where fragment i-1 ends
                f[i, 0], f[i, 1] = f[i - 1, 2], f[i - 1, 3]; f[i, 2], f[i, 3] = f[i, 0] + ln[i], f[i, 1] + ln[i]
            f[n // 2] = f[n // 2 - 1]          # a duplicate
            f[n // 3, 1] -= 50; f[n // 3, 3] -= 50
        f = f[rng.permutation(n)] if seed % 2 else f
        if seed % 3 == 0:
            f[:, [1, 3]] -= 20                 # some negative y: keys wrap as unsigned
        sc = np.where(rng.random(n) < 0.8, f[:, 2] - f[:, 0], rng.integers(0, 50, n)).astype(np.int32)
        out.append((f.astype(np.int32), sc))
    return out
