"""Seeded breakpoint cases for the RefineBreakpoint (a20) parity tests: a read whose two halves align to two loci (any strand combination),
separated by an unaligned span of 1..499 bases that partly continues each locus; loci near contig ends (short target windows)."""
import numpy as np

B = np.frombuffer(b"ACGT", np.uint8)
COMP = np.zeros(256, np.uint8); COMP[[65, 67, 71, 84, 78]] = [84, 71, 67, 65, 78]


def rc(a):
    return COMP[a[::-1]]


def mutate(a, rate, rng):
    """Substitutions at `rate`, plus a few one- and two-base insertions / deletions (so that an extension has several blocks)."""
    a = a.copy()
    m = rng.random(len(a)) < rate
    a[m] = B[rng.integers(0, 4, int(m.sum()))]
    out, i = [], 0
    while i < len(a):
        r = rng.random()
        if r < 0.012:
            i += int(rng.integers(1, 3)); continue            # deletion
        if r < 0.024:
            out.extend(B[rng.integers(0, 4, int(rng.integers(1, 3)))].tolist())   # insertion
        out.append(int(a[i])); i += 1
    return np.array(out, np.uint8)


def make_case(rng, span=None, lstrand=None, rstrand=None, edge=0):
    span = int(rng.integers(1, 500)) if span is None else span
    lstrand = int(rng.integers(0, 2)) if lstrand is None else lstrand
    rstrand = int(rng.integers(0, 2)) if rstrand is None else rstrand
    G1 = B[rng.integers(0, 4, 6000)].copy(); G2 = B[rng.integers(0, 4, 6000)].copy()
    lenA, lenB = int(rng.integers(60, 400)), int(rng.integers(60, 400))
    # loci; edge cases put them close to a contig end so that the target window is shorter than the span
    ta = int(rng.integers(700, 4500)); tb = int(rng.integers(700, 4500))
    if edge == 1:
        ta = len(G1) - lenA - int(rng.integers(0, 40)) if lstrand == 0 else int(rng.integers(0, 40))
    if edge == 2:
        tb = int(rng.integers(0, 40)) if rstrand == 0 else len(G2) - lenB - int(rng.integers(0, 40))
    A = G1[ta:ta + lenA] if lstrand == 0 else rc(G1[ta:ta + lenA])
    contA = G1[ta + lenA:ta + lenA + span] if lstrand == 0 else rc(G1[max(0, ta - span):ta])
    Bp = G2[tb:tb + lenB] if rstrand == 0 else rc(G2[tb:tb + lenB])
    preB = G2[max(0, tb - span):tb] if rstrand == 0 else rc(G2[tb + lenB:tb + lenB + span])
    h = int(rng.integers(0, span + 1))
    gap = np.concatenate([mutate(contA, 0.08, rng)[:h], B[rng.integers(0, 4, span)]])[:span]
    tail = mutate(preB, 0.08, rng)[-(span - h):] if span - h > 0 else np.zeros(0, np.uint8)
    if len(tail):
        gap[span - len(tail):] = tail
    pre, post = B[rng.integers(0, 4, int(rng.integers(0, 300)))], B[rng.integers(0, 4, int(rng.integers(0, 300)))]
    read = np.concatenate([pre, A, gap, Bp, post]).astype(np.uint8)
    L = len(read)
    qa, qb = len(pre), len(pre) + lenA + span
    def blocks(q0, t0, n, strand):        # two blocks with a one-base deletion between them (the refinement only touches the ends)
        n1 = n // 2
        if strand == 0:
            return np.array([[q0, t0, n1], [q0 + n1, t0 + n1, n - n1]], np.uint32)
        q0r = L - (q0 + n)
        return np.array([[q0r, t0, n1], [q0r + n1, t0 + n1, n - n1]], np.uint32)
    return dict(read=read, read_rc=rc(read), L=L, G1=G1, G2=G2, lblocks=blocks(qa, ta, lenA, lstrand), rblocks=blocks(qb, tb, lenB, rstrand),
                lstrand=lstrand, rstrand=rstrand, span=span)


def cases(seed, n=24):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        out.append(make_case(rng, lstrand=i & 1, rstrand=(i >> 1) & 1, edge=(i // 4) % 3, span=[None, 1, 2, 499, 37, None][i % 6]))
    # not within MAX_GAP: overlapping alignments and a gap of 500
    c = make_case(rng, span=3, lstrand=0, rstrand=0); c["rblocks"] = c["rblocks"].copy(); c["rblocks"][:, 0] -= 10; out.append(c)
    return out


def pack(cs):
    """Batch layout of the C ABI: read arenas (forward and reverse complement, same offsets), one genome arena with every case's two contigs."""
    N = np.full(16, ord("N"), np.uint8)
    read_off, roff, goff = [], 0, 0
    fwd, rcs, gen = [], [], []
    bp = {k: [] for k in ["lf", "ll", "rf", "rl", "lstrand", "rstrand", "read_off", "read_len", "lchrom_off", "rchrom_off", "lchrom_len", "rchrom_len"]}
    for c in cs:
        bp["read_off"].append(roff); bp["read_len"].append(c["L"])
        fwd.append(c["read"]); rcs.append(c["read_rc"]); roff += c["L"]
        bp["lchrom_off"].append(goff); bp["lchrom_len"].append(len(c["G1"])); gen.append(c["G1"]); goff += len(c["G1"])
        bp["rchrom_off"].append(goff); bp["rchrom_len"].append(len(c["G2"])); gen.append(c["G2"]); goff += len(c["G2"])
        bp["lf"].append(c["lblocks"][0]); bp["ll"].append(c["lblocks"][-1]); bp["rf"].append(c["rblocks"][0]); bp["rl"].append(c["rblocks"][-1])
        bp["lstrand"].append(c["lstrand"]); bp["rstrand"].append(c["rstrand"])
    return np.concatenate(fwd + [N]), np.concatenate(rcs + [N]), np.concatenate(gen + [N]), bp
