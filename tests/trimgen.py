"""Inputs for the TrimSplitChainDiagonal test: a split chain's anchors (ascending read position for a forward chain, as SPLITChain leaves them reversed;
descending for a reverse chain) and refined anchors scattered around its diagonal, some beyond +-100, some before / after the chain's span, in random order."""
import numpy as np


def case(rng):
    strand = int(rng.random() < 0.4)
    nch = int(rng.choice([1, 2, 3, 6, 12]))
    q0 = int(rng.integers(100, 3000)); t0 = int(rng.integers(10_000, 2_000_000))
    cq = q0 + np.cumsum(rng.integers(20, 400, nch)); diag = t0 + np.cumsum(rng.integers(-60, 61, nch))
    ct = cq + diag if strand == 0 else diag + 50_000 - cq
    if strand == 1:
        cq = cq[::-1].copy(); ct = ct[::-1].copy()
    n = int(rng.integers(0, 120))
    q = rng.integers(max(0, int(cq.min()) - 300), int(cq.max()) + 300, n)
    near = diag[rng.integers(0, nch, n)] + rng.choice([0, 5, -40, 99, 100, 101, -101, 250, -3000], n)
    t = q + near if strand == 0 else near + 50_000 - q
    t = np.clip(t, 0, None)
    return cq.astype(np.uint32), ct.astype(np.uint32), strand, q.astype(np.uint32), t.astype(np.uint32)
