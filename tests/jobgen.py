"""Seeded random AffineOneGapAlign jobs for parity tests (shapes follow SURVEY.md Appendix E)."""
import numpy as np

ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)


def _mutate(seq, err, rng):
    out = []
    for c in seq:
        r = rng.random()
        if r < err / 3:
            continue
        if r < 2 * err / 3:
            out.append(c); out.append(ALPHA[rng.integers(4)])
        elif r < err:
            out.append(ALPHA[(np.searchsorted(ALPHA, c) + rng.integers(1, 4)) & 3])
        else:
            out.append(c)
    return np.array(out, dtype=np.uint8)


def random_job(rng, kind=None):
    """Returns (q bytes, t bytes, k)."""
    kind = kind or rng.choice(["similar", "similar", "similar", "lopsided", "tiny", "junk", "lowrep"])
    if kind == "tiny":
        ql, tl = int(rng.integers(0, 12)), int(rng.integers(0, 12))  # the reference is called with empty windows too
        q = ALPHA[rng.integers(0, 4, ql)]; t = ALPHA[rng.integers(0, 4, tl)]
    elif kind == "junk":
        ql, tl = int(rng.integers(1, 120)), int(rng.integers(1, 120))
        q = ALPHA[rng.integers(0, 4, ql)]; t = ALPHA[rng.integers(0, 4, tl)]
    elif kind == "lowrep":  # low-complexity: many score ties
        n = int(rng.integers(2, 150))
        t = ALPHA[rng.integers(0, 2, n)]
        q = _mutate(t, 0.15, rng)
    elif kind == "lopsided":  # exercises the two-sided (one free gap) mode
        n = int(rng.integers(5, 200))
        t = ALPHA[rng.integers(0, 4, n)]
        q = _mutate(t, float(rng.choice([0.0, 0.02, 0.1])), rng)
        ins = ALPHA[rng.integers(0, 4, int(rng.integers(5, 300)))]
        cut = int(rng.integers(0, len(q) + 1))
        q = np.concatenate([q[:cut], ins, q[cut:]])
        if rng.random() < 0.5:
            q, t = t, q
    else:
        n = int(np.exp(rng.uniform(np.log(2), np.log(1000))))
        t = ALPHA[rng.integers(0, 4, n)]
        q = _mutate(t, float(rng.choice([0.0, 0.01, 0.08, 0.15, 0.3])), rng)
    if kind != "tiny":
        if len(q) == 0:
            q = ALPHA[rng.integers(0, 4, 1)]
        if len(t) == 0:
            t = ALPHA[rng.integers(0, 4, 1)]
    q = q.copy(); t = t.copy()
    if rng.random() < 0.1:  # non-ACGT symbols compare equal to each other (seqMapN, SeqUtils.h:42-75)
        for s in (q, t):
            m = rng.random(len(s)) < 0.05
            s[m] = ord("N")
    if rng.random() < 0.05:
        q = np.frombuffer(bytes(q).lower(), dtype=np.uint8).copy()
    k = int(rng.choice([1, 1, 1, 3, 3, 5, 7, 9, 11, 13, 15, 20, 30, 50]))
    return bytes(q), bytes(t), k


SCORINGS = [(4, -3, -4), (4, -1, -2)]  # CCS/CONTIG and CLR/ONT presets (lra.cpp:268-431, Options.h)


def batch(rng, n, kinds=None):
    """SoA batch: (q_arena, t_arena, q_off, t_off, q_len, t_len, k) as numpy arrays."""
    qs, ts, ks = [], [], []
    for _ in range(n):
        q, t, k = random_job(rng, None if kinds is None else rng.choice(kinds))
        qs.append(q); ts.append(t); ks.append(k)
    return pack(qs, ts, ks)


def pack(qs, ts, ks):
    q_len = np.array([len(x) for x in qs], dtype=np.int32)
    t_len = np.array([len(x) for x in ts], dtype=np.int32)
    q_off = np.zeros(len(qs), dtype=np.uint32); t_off = np.zeros(len(ts), dtype=np.uint32)
    if len(qs) > 1:
        q_off[1:] = np.cumsum(q_len[:-1]); t_off[1:] = np.cumsum(t_len[:-1])
    q_arena = np.frombuffer(b"".join(qs) + b"\0" * 16, dtype=np.uint8).copy()
    t_arena = np.frombuffer(b"".join(ts) + b"\0" * 16, dtype=np.uint8).copy()
    return q_arena, t_arena, q_off, t_off, q_len, t_len, np.array(ks, dtype=np.int32)
