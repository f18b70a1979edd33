"""Inputs for the a11 chain-glue tests (MergeChain, switchindex)."""
import numpy as np


def merge_case(rng):
    """Clusters along a split chain in chain order (read position descending): mostly close neighbours on one strand, with gaps beyond 500 on either
    axis, strand / contig changes, overlaps."""
    n = int(rng.integers(1, 12))
    q = 100_000; strand = int(rng.random() < 0.3); chrom = 0
    t = 500_000
    box, st, ch = [], [], []
    for _ in range(n):
        lq = int(rng.integers(50, 3000)); lt = lq + int(rng.integers(-30, 31))
        gq = int(rng.choice([0, 10, 200, 499, 500, 501, 800, 3000])) - int(rng.choice([0, 0, 0, 40]))
        gt = int(rng.choice([0, 10, 200, 499, 500, 501, 800, 3000])) - int(rng.choice([0, 0, 0, 40]))
        q -= gq + lq
        if strand == 0:
            t -= gt + lt
        else:
            t += gt
        box.append([q, q + lq, t, t + lt]); st.append(strand); ch.append(chrom)
        if strand == 1:
            t += lt
        if rng.random() < 0.1:
            strand ^= 1
        if rng.random() < 0.07:
            chrom += 1
    order = rng.permutation(n)                      # cluster ids are arbitrary; the split chain lists them in chain order
    inv = np.argsort(order)
    return dict(sp=order.astype(np.int32), chrom=np.array(ch, np.int32)[inv], strand=np.array(st, np.uint8)[inv], box=np.array(box, np.uint32)[inv])


def switch_case(rng):
    """A chain over split clusters whose coarse clusters repeat (adjacent repeats, A..A sandwiches, nested repeats), and cluster read ranges of which
    some are covered by their predecessor's."""
    n_cl = int(rng.integers(1, 8)); n_sc = int(rng.integers(n_cl, 4 * n_cl + 1))
    coarse = np.sort(rng.integers(0, n_cl, n_sc)).astype(np.int32)
    n = int(rng.integers(1, 25))
    if rng.random() < 0.6:                          # mostly runs of neighbouring split clusters
        ch = np.clip(np.cumsum(rng.integers(-1, 3, n)) % n_sc, 0, n_sc - 1).astype(np.int32)
    else:
        ch = rng.integers(0, n_sc, n).astype(np.int32)
    link = rng.integers(0, 2, n - 1).astype(np.uint8)
    qs = rng.integers(0, 5000, n_cl); qe = qs + rng.integers(1, 5000, n_cl)
    if n_cl > 1 and rng.random() < 0.5:             # nest one cluster inside another
        a, b = rng.choice(n_cl, 2, replace=False)
        qs[a] = qs[b] + 1; qe[a] = qe[b] - 1 if qe[b] - 1 > qs[a] else qe[b]
    return dict(ch=ch, link=link, coarse=coarse, cq=np.stack([qs, qe], 1).astype(np.uint32))
