"""Inputs for the a11 tests (SPLITChain on an UltimateChain + MergeSplitchainINS + RemoveSpuriousSplitChain): anchor chains in the order the second
sparse DP leaves them (read position descending), made of runs on one strand / contig joined by the events the splitter looks for -- unchained
gaps (>= 1000 on both axes along one diagonal), translocations (beyond opts.splitdist, or another contig), strand switches, insertion-like
excursions that come back within 1500 bases (merged again), runs of one or two anchors (removed), and runs that straddle a contig boundary
(dropped by SplitChain::CHROMIndex)."""
import numpy as np

HDR = np.array([0, 300_000, 600_000, 900_000], np.uint64)     # three contigs of 300 kb


def chain(rng, n_runs=None):
    n_runs = int(rng.integers(1, 9)) if n_runs is None else n_runs
    q = 200_000
    strand = int(rng.random() < 0.3)
    t = int(rng.integers(150_000, 250_000)) + 300_000 * int(rng.integers(0, 3))
    cnum = 0
    Q, T, L, S, C = [], [], [], [], []
    saved = None
    for r in range(n_runs):
        k = int(rng.choice([1, 2, 3, 5, 12, 40], p=[0.1, 0.1, 0.1, 0.2, 0.3, 0.2]))
        for j in range(k):
            ln = int(rng.integers(17, 70))
            gap = int(rng.integers(0, 150))
            big = rng.random() < 0.04
            if big:
                gap = int(rng.integers(1000, 4000))
            jit = int(rng.integers(-8, 9)) if not big else int(rng.integers(-200, 201))
            q -= ln + gap
            if strand == 0:
                t -= ln + gap + jit
            else:
                t += (L[-1] if L and S[-1] == 1 else 17) + gap + jit
            if q < 100 or t < 100 or t > 899_000:
                break
            Q.append(q); T.append(t); L.append(ln); S.append(strand); C.append(cnum)
        ev = rng.random()
        if saved is not None and rng.random() < 0.8:              # come back from an excursion: close to where the chain left
            t, strand = saved; saved = None
            if strand == 0:
                t -= int(rng.integers(0, 1200))
            else:
                t += int(rng.integers(0, 1200))
            cnum += int(rng.random() < 0.7)
        elif ev < 0.35:                                            # translocation (far on the same contig, or another contig), maybe an excursion
            if rng.random() < 0.6:
                saved = (t, strand)
            t = int(rng.integers(20_000, 280_000)) + 300_000 * int(rng.integers(0, 3))
            cnum += 1
        elif ev < 0.6:                                             # inversion
            strand ^= 1; cnum += 1
            t += int(rng.integers(-2000, 2000))
        elif ev < 0.7:                                             # jump next to a contig boundary
            t = 300_000 * int(rng.integers(1, 3)) + int(rng.integers(-300, 300)); cnum += 1
        else:                                                      # same place, maybe another cluster
            cnum += int(rng.random() < 0.5)
    if not Q:
        return chain(rng, n_runs)
    n = len(Q)
    # clusters must have one strand: renumber so that (cnum, strand) pairs are distinct clusters
    key = {}
    cn = np.array([key.setdefault((c, s), len(key)) for c, s in zip(C, S)], np.int32)
    return dict(q=np.array(Q, np.uint32), t=np.array(T, np.uint32), len=np.array(L, np.int32), strand=np.array(S, np.uint8), cnum=cn,
                link=rng.integers(0, 2, max(n - 1, 0)).astype(np.uint8))


def chains(seed, n):
    rng = np.random.default_rng(seed)
    return [chain(rng) for _ in range(n)]
