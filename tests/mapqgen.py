"""Seeded reads for the MAPQ / ordering (a22) parity tests: 1..40 alignments per read (more than 16 exercises the introsort partition), 1..4
segments each, ties in value and in NumOfAnchors0, near-equal top two values (the x >= 0.990 branch), tiny anchor counts, perfect identity,
zero values, supplementary / inversion segments; the Update schedule of either pipeline."""
import numpy as np


def make_read(rng, G, schedule):
    nseg = rng.integers(1, 5, G)
    so = np.zeros(G + 1, np.int32); so[1:] = np.cumsum(nseg)
    S = int(so[-1])
    base = rng.choice([5.0, 300.0, 2500.0, 12000.0])
    value = (base * rng.choice([1.0, 0.995, 0.98, 0.7, 0.3, 0.0], S) + rng.choice([0.0, 0.0, 0.25, 17.5], S)).astype(np.float32)
    n0 = rng.choice([1, 3, 4, 5, 9, 10, 11, 20, 21, 60, 400], S).astype(np.int32)
    if G > 3:
        value[so[1]:so[2]] = value[so[0]:so[1]][:1]          # ties in the group sums are likely
    nm = rng.integers(0, 20000, S).astype(np.int32); nmm = rng.integers(0, 600, S).astype(np.int32)
    ndel = rng.integers(0, 300, S).astype(np.int32); nins = rng.integers(0, 300, S).astype(np.int32)
    perfect = rng.random(S) < 0.15
    nmm[perfect] = 0; ndel[perfect] = 0; nins[perfect] = 0
    rd = dict(seg_off=so, value=value, n0=n0, n1=rng.integers(0, 3000, S).astype(np.int32), nm=nm, nmm=nmm, ndel=ndel, nins=nins,
              strand=(rng.random(S) < 0.5).astype(np.uint8), flag=np.zeros(S, np.int32), typeofaln=rng.choice([0, 0, 0, 1, 3], S).astype(np.int32),
              issec=(rng.random(S) < 0.3).astype(np.uint8), supp=(rng.random(S) < 0.4).astype(np.uint8))
    if schedule == 0:
        rd["update_at"] = np.array([G], np.int32)                          # Map_lowacc.h:609
    else:
        cuts = np.sort(rng.choice(np.arange(1, G + 1), size=min(G, int(rng.integers(1, 5))), replace=False)).tolist()
        if cuts[-1] != G:
            cuts.append(G)
        rd["update_at"] = np.array(cuts + [G], np.int32)                   # Map_highacc.h:737 (per primary chain) and :743 (once more)
    return rd


def reads(seed, schedule):
    rng = np.random.default_rng(seed)
    return [make_read(rng, G, schedule) for G in (1, 1, 2, 2, 2, 3, 5, 9, 17, 40)]
