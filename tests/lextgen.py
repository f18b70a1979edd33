"""Inputs for the a15 tests (LinearExtend on GenomePairs + DecideCoordinates + TrimOverlappedAnchors): reads that are mutated copies of
genome windows, with the K-mer anchors a minimizer sampler would leave along the true alignment, a few off-diagonal noise anchors, tandem
repeats around indels (so that extended anchors of neighbouring diagonals overlap and TrimOverlappedAnchors has work), forward and reverse
parts (the reference compares bases without complementing on strand 1, so a "reverse" read here is the plain reversal of the window),
1-3 parts per group."""
import numpy as np

B = np.frombuffer(b"ACGT", np.uint8)


def make_genome(rng, n_contigs=3, length=60_000):
    contigs = []
    for _ in range(n_contigs):
        g = B[rng.integers(0, 4, length)].copy()
        for _ in range(length // 1500):          # short tandem repeats, period 1-4, 16-36 bases
            p = int(rng.integers(1, 5)); n = int(rng.integers(16, 37)); s = int(rng.integers(0, length - 64))
            unit = B[rng.integers(0, 4, p)]
            g[s:s + n] = np.resize(unit, n)
        contigs.append(g)
    off = np.zeros(n_contigs, np.uint64); off[1:] = np.cumsum([len(c) for c in contigs[:-1]])
    return np.concatenate(contigs + [np.full(16, ord("A"), np.uint8)]), off, np.array([len(c) for c in contigs], np.int32)


def mutate(rng, win, sub, indel):
    """Returns (read, pairs) with pairs = [(q, t)] for every aligned (match or mismatch) column."""
    out, pairs = [], []
    t = 0
    while t < len(win):
        x = rng.random()
        if x < indel / 2:                        # deletion of 1-3 bases (of one repeat unit, often)
            t += int(rng.integers(1, 4)); continue
        if x < indel:                            # insertion: copy of the previous bases (keeps tandem repeats in phase) or random
            n = int(rng.integers(1, 4))
            for i in range(n):
                out.append(out[-n] if len(out) >= n and rng.random() < 0.7 else B[rng.integers(0, 4)])
            continue
        b = win[t]
        if rng.random() < sub:
            b = B[(int(np.searchsorted(B, b)) + int(rng.integers(1, 4))) & 3]
        pairs.append((len(out), t)); out.append(b); t += 1
    return np.array(out, np.uint8), pairs


def one_read(rng, arena, c_off, c_len, K, read_len, sub=0.03, indel=0.03, max_parts=4, noise=0.04, unsorted=False):
    """One read with 1..max_parts parts.  Returns (read bytes, dict(q, t, p_off, p_strand, chrom_off, chrom_len))."""
    n_parts = int(rng.integers(1, max_parts + 1))
    pieces, part_rows = [], []
    qbase = 0
    for p in range(n_parts):
        c = int(rng.integers(0, len(c_len))); L = int(rng.integers(read_len // 2, read_len))
        s = int(rng.integers(0, int(c_len[c]) - L))
        win = arena[int(c_off[c]) + s:int(c_off[c]) + s + L]
        strand = int(rng.random() < 0.35)
        src = win[::-1] if strand else win
        piece, pairs = mutate(rng, src, sub, indel)
        # exact K-mer matches along the alignment: runs of >= K aligned matching columns without an indel in between
        qs = np.array([a for a, _ in pairs]); ts = np.array([b for _, b in pairs])
        eq = piece[qs] == src[ts]
        anchors = []
        i = 0
        while i + K <= len(pairs):
            j = i + K - 1
            if qs[j] - qs[i] == K - 1 and ts[j] - ts[i] == K - 1 and eq[i:j + 1].all():
                tt = s + int(ts[i]) if strand == 0 else s + (L - 1 - int(ts[i])) - (K - 1)      # contig position of the K-mer's lowest base
                anchors.append((qbase + int(qs[i]), tt))
                i += int(rng.integers(1, 13))
            else:
                i += 1
        for _ in range(int(noise * len(anchors)) + 1):
            anchors.append((qbase + int(rng.integers(0, max(1, len(piece) - K))), int(rng.integers(0, int(c_len[c]) - K))))
        if rng.random() < 0.15:                  # anchors at the very ends of the contig / read exercise the clamps
            anchors.append((qbase + max(0, len(piece) - K), int(c_len[c]) - K)); anchors.append((qbase, 0))
        a = np.array(anchors, np.int64).reshape(-1, 2)
        if len(a) and not unsorted:
            d = a[:, 0] - a[:, 1]
            a = a[np.lexsort((a[:, 0], d))]
        elif len(a):
            a = a[rng.permutation(len(a))]
        if rng.random() < 0.05:
            a = a[:0]
        pieces.append(piece); qbase += len(piece)
        part_rows.append((a, strand, int(c_off[c]), int(c_len[c])))
    read = np.concatenate(pieces)
    p_off = np.zeros(n_parts + 1, np.int32); p_off[1:] = np.cumsum([len(r[0]) for r in part_rows])
    allq = np.concatenate([r[0][:, 0] for r in part_rows]).astype(np.uint32) if p_off[-1] else np.zeros(0, np.uint32)
    allt = np.concatenate([r[0][:, 1] for r in part_rows]).astype(np.uint32) if p_off[-1] else np.zeros(0, np.uint32)
    return read, dict(q=allq, t=allt, p_off=p_off, p_strand=np.array([r[1] for r in part_rows], np.uint8),
                      chrom_off=np.array([r[2] for r in part_rows], np.uint64), chrom_len=np.array([r[3] for r in part_rows], np.int32))


def group(rng, rd, single):
    """g_off: one part per group (Map_lowacc.h:132-136) or 1-3 consecutive parts per group (the merged split chains of :460-474)."""
    n = len(rd["p_strand"])
    if single:
        g = np.arange(n + 1, dtype=np.int32)
    else:
        cuts = [0]
        while cuts[-1] < n:
            cuts.append(min(n, cuts[-1] + int(rng.integers(1, 4))))
        g = np.array(cuts, np.int32)
    rd = dict(rd); rd["g_off"] = g
    return rd


def handmade(K=17):
    """Deterministic cases: two diagonals around a 2-base insertion inside a 24-base dinucleotide repeat, both anchors extended over the repeat
    (>= 40 long, overlapping by <= 30: trimmed); identical long anchors (a tie for the LongAnchors sort); an anchor whose extension reaches the
    next one exactly (merged); an extension that stops at a mismatch; anchors at the contig end."""
    rng = np.random.default_rng(5)
    left = B[rng.integers(0, 4, 60)]; right = B[rng.integers(0, 4, 60)]
    rep = np.resize(np.frombuffer(b"AC", np.uint8), 24)
    left[-1] = ord("G"); right[0] = ord("T")
    contig = np.concatenate([left, rep, right, B[rng.integers(0, 4, 200)]])
    read = np.concatenate([left, rep, np.frombuffer(b"AC", np.uint8), right, contig[144:344]])
    read[215] = B[(int(np.searchsorted(B, read[215])) + 1) & 3]
    clen = len(contig)
    # diagonal 0: (0,0), (30,30), (60,60), (67,67) -> one anchor 0..84 that runs through the repeat; diagonal +2 starts inside the repeat at (62,60) and runs
    # on through (90,88), (110,108), (150,148), (190,188); the extension after (190,188) stops at the mismatch at q = 215, (230,228) starts a new anchor
    q = np.array([0, 30, 60, 67, 62, 90, 110, 150, 190, 230, clen + 2 - K], np.uint32)
    t = np.array([0, 30, 60, 67, 60, 88, 108, 148, 188, 228, clen - K], np.uint32)
    arena = np.concatenate([contig, np.full(16, ord("A"), np.uint8)])
    rd = dict(q=q, t=t, p_off=np.array([0, len(q)], np.int32), p_strand=np.zeros(1, np.uint8), chrom_off=np.zeros(1, np.uint64), chrom_len=np.array([clen], np.int32),
              g_off=np.array([0, 1], np.int32))
    # a second case: the same part twice in one group (duplicate long anchors = ties under LongAnchors)
    rd2 = dict(q=np.concatenate([q, q]), t=np.concatenate([t, t]), p_off=np.array([0, len(q), 2 * len(q)], np.int32), p_strand=np.zeros(2, np.uint8),
               chrom_off=np.zeros(2, np.uint64), chrom_len=np.array([clen, clen], np.int32), g_off=np.array([0, 2], np.int32))
    return arena, read, rd, rd2


def reads(seed, n_reads, K=17, read_len=3000, single=True, unsorted=False, **kw):
    rng = np.random.default_rng(seed)
    arena, c_off, c_len = make_genome(rng)
    out = []
    for _ in range(n_reads):
        read, rd = one_read(rng, arena, c_off, c_len, K, read_len, unsorted=unsorted, **kw)
        out.append((read, group(rng, rd, single)))
    return arena, out


def to_batch(items):
    """Concatenate per-read inputs into the batch layout of lra_b200_linear_extend_batch."""
    roff = np.zeros(len(items), np.uint64); rlen = np.array([len(r) for r, _ in items], np.uint32)
    roff[1:] = np.cumsum(rlen[:-1])
    read_arena = np.concatenate([r for r, _ in items] + [np.full(16, ord("A"), np.uint8)])
    g_off, p_off, q, t, st, co, cl, pro, prl = [0], [0], [], [], [], [], [], [], []
    for i, (_, rd) in enumerate(items):
        np_ = len(rd["p_strand"])
        g_off += [len(st) + int(x) for x in rd["g_off"][1:]]
        p_off += [p_off[-1] + int(rd["p_off"][k + 1] - rd["p_off"][0]) for k in range(np_)]
        q.append(rd["q"]); t.append(rd["t"]); st += list(rd["p_strand"]); co += list(rd["chrom_off"]); cl += list(rd["chrom_len"])
        pro += [int(roff[i])] * np_; prl += [int(rlen[i])] * np_
    return read_arena, dict(g_off=np.array(g_off, np.uint64), p_off=np.array(p_off, np.uint64), p_strand=np.array(st, np.uint8), chrom_off=np.array(co, np.uint64),
                            chrom_len=np.array(cl, np.uint32), read_off=np.array(pro, np.uint64), read_len=np.array(prl, np.uint32),
                            q=np.concatenate(q).astype(np.uint32) if q else np.zeros(0, np.uint32), t=np.concatenate(t).astype(np.uint32) if t else np.zeros(0, np.uint32))


# ------------------------------------------------------------------------------------------------ the high-accuracy overload (chains of clusters)

def chain_read(rng, arena, c_off, c_len, K, read_len, sub=0.02, indel=0.02):
    """One read with 1-2 chains over clusters cut from mutated windows with OVERLAPPING read ranges (so that the neighbours' box corners fall
    inside a cluster and CheckOverlap has work).  Returns (read, dict(cl_off, q, t, box, strand, freq, chrom_off, chrom_len), [chain, ...])."""
    pieces, clusters, chains = [], [], []
    qbase = 0
    for _ in range(int(rng.integers(1, 3))):
        c = int(rng.integers(0, len(c_len))); L = int(rng.integers(read_len // 2, read_len))
        s = int(rng.integers(0, int(c_len[c]) - L))
        win = arena[int(c_off[c]) + s:int(c_off[c]) + s + L]
        strand = int(rng.random() < 0.35)
        src = win[::-1] if strand else win
        piece, pairs = mutate(rng, src, sub, indel)
        qs = np.array([a for a, _ in pairs]); ts = np.array([b for _, b in pairs])
        eq = piece[qs] == src[ts]
        anchors = []
        i = 0
        while i + K <= len(pairs):
            j = i + K - 1
            if qs[j] - qs[i] == K - 1 and ts[j] - ts[i] == K - 1 and eq[i:j + 1].all():
                tt = s + int(ts[i]) if strand == 0 else s + (L - 1 - int(ts[i])) - (K - 1)
                anchors.append((qbase + int(qs[i]), tt)); i += int(rng.integers(1, 10))
            else:
                i += 1
        a = np.array(anchors, np.int64).reshape(-1, 2)
        ncl = int(rng.integers(1, 5))
        cuts = np.sort(rng.integers(qbase, qbase + len(piece), ncl - 1)) if ncl > 1 else np.zeros(0, np.int64)
        bounds = np.concatenate([[qbase], cuts, [qbase + len(piece)]])
        chain = []
        for k in range(ncl):
            ov = int(rng.integers(0, 90))
            sel = a[(a[:, 0] >= bounds[k] - ov) & (a[:, 0] < bounds[k + 1] + ov)] if len(a) else a
            if len(sel) == 0:
                continue
            sel = sel[rng.permutation(len(sel))]
            box = [int(sel[:, 0].min()), int(sel[:, 0].max()) + K, int(sel[:, 1].min()), int(sel[:, 1].max()) + K]
            chain.append(len(clusters))
            clusters.append((sel, box, strand, float(rng.choice([1.0, 1.05, 1.1, 1.3])), int(c_off[c]), int(c_len[c])))
        if rng.random() < 0.3:
            chain = chain[::-1]
        if chain:
            chains.append(np.array(chain, np.int32))
        pieces.append(piece); qbase += len(piece)
    read = np.concatenate(pieces)
    if not clusters:
        return chain_read(rng, arena, c_off, c_len, K, read_len, sub, indel)
    cl_off = np.zeros(len(clusters) + 1, np.int32); cl_off[1:] = np.cumsum([len(c[0]) for c in clusters])
    cd = dict(cl_off=cl_off, q=np.concatenate([c[0][:, 0] for c in clusters]).astype(np.uint32), t=np.concatenate([c[0][:, 1] for c in clusters]).astype(np.uint32),
              box=np.array([c[1] for c in clusters], np.uint32), strand=np.array([c[2] for c in clusters], np.uint8), freq=np.array([c[3] for c in clusters], np.float32),
              chrom_off=np.array([c[4] for c in clusters], np.uint64), chrom_len=np.array([c[5] for c in clusters], np.int32))
    return read, cd, chains


def chain_reads(seed, n_reads, K=17, read_len=3000):
    rng = np.random.default_rng(seed)
    arena, c_off, c_len = make_genome(rng)
    return arena, [chain_read(rng, arena, c_off, c_len, K, read_len) for _ in range(n_reads)]


def chains_to_batch(items):
    """Concatenate per-read (read, cd, chains) into the batch layout of lra_b200_linear_extend_chains_batch."""
    roff = np.zeros(len(items), np.uint64); rlen = np.array([len(r) for r, _, _ in items], np.uint32)
    roff[1:] = np.cumsum(rlen[:-1])
    read_arena = np.concatenate([r for r, _, _ in items] + [np.full(16, ord("A"), np.uint8)])
    ch_off, ch, cl_off = [0], [], [0]
    cols = {k: [] for k in ["q", "t", "box", "strand", "freq", "chrom_off", "chrom_len", "read_off", "read_len"]}
    for i, (_, cd, chains) in enumerate(items):
        base = len(cl_off) - 1
        ncl = len(cd["strand"])
        cl_off += [cl_off[-1] + int(cd["cl_off"][k + 1]) for k in range(ncl)]
        for c in chains:
            ch += [base + int(x) for x in c]; ch_off.append(len(ch))
        for k in ["q", "t", "box", "strand", "freq", "chrom_off", "chrom_len"]:
            cols[k].append(cd[k])
        cols["read_off"].append(np.full(ncl, roff[i], np.uint64)); cols["read_len"].append(np.full(ncl, rlen[i], np.uint32))
    out = {k: np.concatenate(v) for k, v in cols.items()}
    out.update(ch_off=np.array(ch_off, np.uint64), ch=np.array(ch, np.uint32), cl_off=np.array(cl_off, np.uint64))
    return read_arena, out
