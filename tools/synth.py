#!/usr/bin/env python
"""Seeded synthetic references and reads for the BASELINE.json configs.

Reference: uniform i.i.d. ACGT, upper-case, 80-column FASTA, contigs chr1..chrN.
Reads: start uniform over a contig, strand 50/50, i.i.d. per-base errors with
sub:ins:del = 1:1:1, FASTA, names r<idx>_<chr>_<start>_<strand>.

Used by tests (small sizes), by bench.py (to draw job shapes) and by tools/make_golden.py.
"""
import argparse
import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGTN", b"TGCAN"):
    COMP[a] = b


def gen_ref(total_len, n_contigs=1, seed=1234):
    """Return list of (name, uint8 array of ASCII bases)."""
    rng = np.random.default_rng(seed)
    per = total_len // n_contigs
    out = []
    for c in range(n_contigs):
        n = per if c < n_contigs - 1 else total_len - per * (n_contigs - 1)
        out.append(("chr%d" % (c + 1), BASES[rng.integers(0, 4, size=n, dtype=np.uint8)]))
    return out


def add_repeats(ref, frac=0.05, seed=77):
    """Repeat-enriched variant: overwrite `frac` of every contig with copies of 1-10 kb segments of the same
    contig set at 0-2 % divergence, so that minimizers repeat and anchor / chain ties occur."""
    rng = np.random.default_rng(seed)
    out = [(n, s.copy()) for n, s in ref]
    for ci, (name, seq) in enumerate(out):
        budget = int(len(seq) * frac)
        while budget > 0:
            L = int(rng.integers(1000, 10001))
            sc = int(rng.integers(0, len(out)))
            src = out[sc][1]
            if len(src) <= L or len(seq) <= L:
                break
            a = int(rng.integers(0, len(src) - L))
            b = int(rng.integers(0, len(seq) - L))
            piece = mutate(src[a:a + L], float(rng.uniform(0.0, 0.02)), rng)[:L]
            if rng.integers(0, 2):
                piece = COMP[piece[::-1]]
            seq[b:b + len(piece)] = piece
            budget -= L
    return out


def write_fasta(path, records, width=80):
    with open(path, "wb") as f:
        for name, seq in records:
            f.write(b">" + name.encode() + b"\n")
            n = len(seq)
            full = (n // width) * width
            if full:
                body = np.empty((full // width, width + 1), dtype=np.uint8)
                body[:, :width] = seq[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(seq[full:].tobytes() + b"\n")


def mutate(seq, err, rng):
    """Apply i.i.d. errors (sub:ins:del = 1:1:1) at total rate `err`."""
    n = len(seq)
    if err <= 0:
        return seq.copy()
    r = rng.random(n)
    kind = np.zeros(n, dtype=np.uint8)  # 0 keep, 1 sub, 2 ins (after base), 3 del
    kind[r < err] = 1
    kind[r < 2 * err / 3] = 2
    kind[r < err / 3] = 3
    sub_shift = rng.integers(1, 4, size=n, dtype=np.uint8)
    ins_base = BASES[rng.integers(0, 4, size=n, dtype=np.uint8)]
    code = np.searchsorted(BASES, seq).astype(np.uint8) & 3
    subbed = BASES[(code + sub_shift) & 3]
    base = np.where(kind == 1, subbed, seq)
    cnt = np.ones(n, dtype=np.int64)
    cnt[kind == 3] = 0
    cnt[kind == 2] = 2
    total = int(cnt.sum())
    out = np.empty(total, dtype=np.uint8)
    pos = np.cumsum(cnt) - cnt
    keep = kind != 3
    out[pos[keep]] = base[keep]
    ins = kind == 2
    out[pos[ins] + 1] = ins_base[ins]
    return out


def read_lengths(profile, n, rng):
    if profile == "ccs10k":
        return np.full(n, 10000, dtype=np.int64)
    if profile == "hifi":
        return np.maximum(5000, rng.normal(15000, 2000, size=n)).astype(np.int64)
    if profile == "ont":  # log-normal, N50 ~ 20 kb, sigma 0.5, min 1 kb
        sigma = 0.5
        mu = np.log(20000.0) - sigma * sigma  # N50 of a log-normal = exp(mu + sigma^2)
        return np.maximum(1000, rng.lognormal(mu, sigma, size=n)).astype(np.int64)
    if profile == "clr":
        sigma = 0.5
        mu = np.log(12000.0) - sigma * sigma
        return np.maximum(1000, rng.lognormal(mu, sigma, size=n)).astype(np.int64)
    if profile == "contig":  # assembly contigs scaled to the test references: 100-400 kb
        return rng.integers(100000, 400000, size=n).astype(np.int64)
    if profile == "contig_long":  # megabase-scale contigs (BASELINE configs[4] is 1-10 Mb)
        return rng.integers(1000000, 2000000, size=n).astype(np.int64)
    raise ValueError(profile)


PROFILE_ERR = {"contig_long": 0.001, "ccs10k": 0.01, "hifi": 0.005, "ont": 0.08, "clr": 0.12, "contig": 0.001}


def gen_reads(ref, n_reads, profile, seed, err=None):
    """Return list of (name, uint8 seq). `ref` is the list from gen_ref."""
    rng = np.random.default_rng(seed)
    err = PROFILE_ERR[profile] if err is None else err
    lens = read_lengths(profile, n_reads, rng)
    sizes = np.array([len(s) for _, s in ref], dtype=np.float64)
    contig = rng.choice(len(ref), size=n_reads, p=sizes / sizes.sum())
    out = []
    for i in range(n_reads):
        name, seq = ref[contig[i]]
        L = int(min(lens[i], len(seq)))
        start = int(rng.integers(0, len(seq) - L + 1))
        frag = seq[start:start + L]
        strand = int(rng.integers(0, 2))
        if strand:
            frag = COMP[frag[::-1]]
        out.append(("r%d_%s_%d_%s" % (i, name, start, "-" if strand else "+"), mutate(frag, err, rng)))
    return out


def gen_reads_torch(genome, hdr, names, n_reads, profile, seed, device="cpu", err=None):
    """The same read model as gen_reads, vectorised over the whole batch with torch (bench.py: 16 k reads = 0.3 Gbase per call).
    genome: uint8 torch tensor of the concatenated contigs (ASCII) on `device`; hdr: cumulative contig offsets (len n_contigs + 1).
    Returns (ascii uint8 numpy array of all reads back to back, read_off uint64, read_len uint32, names list).
    The random streams differ from gen_reads (different generator), the distribution does not."""
    import torch
    rng = np.random.default_rng(seed)
    err = PROFILE_ERR[profile] if err is None else err
    hdr = np.asarray(hdr, dtype=np.int64)
    sizes = np.diff(hdr).astype(np.float64)
    lens = read_lengths(profile, n_reads, rng)
    contig = rng.choice(len(sizes), size=n_reads, p=sizes / sizes.sum())
    lens = np.minimum(lens, np.diff(hdr)[contig])
    start = (rng.random(n_reads) * (np.diff(hdr)[contig] - lens + 1)).astype(np.int64)
    strand = rng.integers(0, 2, size=n_reads)
    g = torch.Generator(device=device); g.manual_seed(int(seed) * 7919 + 13)
    L = torch.from_numpy(lens).to(device); off = torch.cumsum(L, 0) - L
    total = int(lens.sum())
    rid = torch.repeat_interleave(torch.arange(n_reads, device=device), L, output_size=total)
    pin = torch.arange(total, device=device) - off[rid]
    gs = torch.from_numpy(hdr[contig] + start).to(device)[rid]
    st = torch.from_numpy(strand).to(device)[rid]
    src = torch.where(st == 0, gs + pin, gs + L[rid] - 1 - pin)
    del pin, gs
    base = genome[src]
    del src
    comp = torch.from_numpy(COMP).to(device)
    base = torch.where(st == 0, base, comp[base.long()])
    del st
    r = torch.rand(total, generator=g, device=device)
    code = torch.zeros(256, dtype=torch.uint8, device=device)
    for i, b in enumerate(b"ACGT"):
        code[b] = i
    bases_t = torch.from_numpy(BASES.copy()).to(device)
    shift = torch.randint(1, 4, (total,), generator=g, device=device, dtype=torch.uint8)
    subbed = bases_t[((code[base.long()] + shift) & 3).long()]
    insb = bases_t[torch.randint(0, 4, (total,), generator=g, device=device)]
    is_sub = (r < err) & (r >= 2 * err / 3)
    is_ins = (r < 2 * err / 3) & (r >= err / 3)
    is_del = r < err / 3
    base = torch.where(is_sub, subbed, base)
    cnt = torch.ones(total, dtype=torch.int64, device=device)
    cnt[is_del] = 0
    cnt[is_ins] = 2
    pos = torch.cumsum(cnt, 0) - cnt
    n_out = int(cnt.sum())
    out = torch.empty(n_out, dtype=torch.uint8, device=device)
    keep = ~is_del
    out[pos[keep]] = base[keep]
    out[pos[is_ins] + 1] = insb[is_ins]
    new_off = pos[off]
    new_off_np = new_off.cpu().numpy().astype(np.uint64)
    new_len = np.diff(np.concatenate([new_off_np, [np.uint64(n_out)]]).astype(np.int64)).astype(np.uint32)
    rd_names = ["r%d_%s_%d_%s" % (i, names[contig[i]], start[i], "-" if strand[i] else "+") for i in range(n_reads)]
    return out.cpu().numpy(), new_off_np, new_len, rd_names


def write_reads_fasta(path, ascii_reads, read_off, read_len, names):
    with open(path, "wb") as f:
        for i, nm in enumerate(names):
            o = int(read_off[i])
            f.write(b">" + nm.encode() + b"\n" + ascii_reads[o:o + int(read_len[i])].tobytes() + b"\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-len", type=int, default=5_000_000)
    ap.add_argument("--contigs", type=int, default=1)
    ap.add_argument("--ref-seed", type=int, default=1234)
    ap.add_argument("--reads", type=int, default=1000)
    ap.add_argument("--profile", default="ccs10k")
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--ref-out", required=True)
    ap.add_argument("--reads-out", required=True)
    a = ap.parse_args()
    ref = gen_ref(a.ref_len, a.contigs, a.ref_seed)
    write_fasta(a.ref_out, ref)
    write_fasta(a.reads_out, gen_reads(ref, a.reads, a.profile, a.seed), width=1 << 30)


if __name__ == "__main__":
    main()


def gen_sv_reads(ref, n_reads, profile, seed, err=None):
    """Reads with structural differences from the reference (for the multi-segment paths: split chains, inversion probing, supplementary
    segments, SA:Z): per read one of deletion (0.5-8 kb of reference skipped), insertion (0.3-3 kb of random sequence), inversion (0.5-4 kb
    reverse-complemented in place), translocation (second half from another locus / contig), tandem duplication (1-3 kb repeated), or none."""
    rng = np.random.default_rng(seed)
    err = PROFILE_ERR[profile] if err is None else err
    lens = read_lengths(profile, n_reads, rng)
    out = []
    kinds = ["del", "ins", "inv", "tra", "dup", "none"]
    for i in range(n_reads):
        c = int(rng.integers(0, len(ref)))
        name, seq = ref[c]
        L = int(min(max(lens[i], 6000), len(seq) - 20000))
        start = int(rng.integers(0, len(seq) - L - 10000))
        kind = kinds[i % len(kinds)]
        a = L // 3 + int(rng.integers(0, L // 3))
        if kind == "del":
            d = int(rng.integers(500, 8000))
            frag = np.concatenate([seq[start:start + a], seq[start + a + d:start + L + d]])
        elif kind == "ins":
            d = int(rng.integers(300, 3000))
            frag = np.concatenate([seq[start:start + a], BASES[rng.integers(0, 4, size=d, dtype=np.uint8)], seq[start + a:start + L]])
        elif kind == "inv":
            d = int(rng.integers(500, 4000))
            d = min(d, L - a - 100)
            frag = np.concatenate([seq[start:start + a], COMP[seq[start + a:start + a + d][::-1]], seq[start + a + d:start + L]])
        elif kind == "tra":
            c2 = int(rng.integers(0, len(ref)))
            s2 = ref[c2][1]
            b = int(rng.integers(0, len(s2) - L))
            piece = s2[b:b + L - a]
            if rng.integers(0, 2):
                piece = COMP[piece[::-1]]
            frag = np.concatenate([seq[start:start + a], piece])
        elif kind == "dup":
            d = int(rng.integers(1000, 3000))
            d = min(d, a)
            frag = np.concatenate([seq[start:start + a], seq[start + a - d:start + a], seq[start + a:start + L]])
        else:
            frag = seq[start:start + L]
        strand = int(rng.integers(0, 2))
        if strand:
            frag = COMP[frag[::-1]]
        out.append(("sv%d_%s_%s_%d_%s" % (i, kind, name, start, "-" if strand else "+"), mutate(frag, err, rng)))
    return out
