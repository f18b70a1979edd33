"""Seeded chains for the chain-filter (a16) parity tests: anchors along a diagonal with paired indels of all the sizes the filters test
for (6..50, 31.., 100.., 300.., 500..), short and long anchors between them, strand switches, outliers at the ends."""
import numpy as np


def make_chain(rng, n, strand_mix=False):
    q = np.zeros(n, np.int64); t = np.zeros(n, np.int64)
    ln = rng.choice([12, 17, 25, 40, 49, 50, 51, 60, 99, 100, 120], n)
    st = np.zeros(n, np.uint8)
    cq, ct = int(rng.integers(0, 500)), int(rng.integers(10000, 2_000_000_000))
    pending = 0
    for i in range(n):
        step = int(rng.integers(20, 200))
        cq += step; ct += step
        r = rng.random()
        if pending and rng.random() < 0.7:                 # close a paired indel (almost equal size, opposite sign) one or two anchors later
            ct -= pending + int(rng.integers(-15, 15)); pending = 0
        elif r < 0.12:
            g = int(rng.choice([6, 20, 35, 50, 80, 120, 320, 520, 700])) * int(rng.choice([-1, 1]))
            ct += g; pending = g
        elif r < 0.15:
            ct += int(rng.integers(-3000, 3000))
        q[i], t[i] = cq, max(ct, 0)
        cq += int(ln[i]) // 2
    if strand_mix and n > 6:
        a = int(rng.integers(1, n - 3)); st[a:a + int(rng.integers(1, 4))] = 1
    if n > 8 and rng.random() < 0.5:                       # outliers at the ends (refineEnds)
        t[0] += 50000; q[-1] += 30000
    return q.astype(np.uint32), t.astype(np.uint32), ln.astype(np.uint32), st


def chains(seed, sizes=(0, 1, 2, 3, 5, 9, 30, 120, 600)):
    rng = np.random.default_rng(seed)
    return [make_chain(rng, n, strand_mix=(i % 2 == 1)) for i, n in enumerate(sizes * 2)]
