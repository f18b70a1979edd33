"""Small seeded inputs for the seeding (a1-a5) parity tests: a genome with a repeat and N runs, an index of it built with the
oracle's own StoreMinimizers + sort (a plain sorted minimizer list is a valid `genomemm` for CompareLists), and reads."""
import numpy as np
from oracle import pyoracle as po

B = np.frombuffer(b"ACGT", np.uint8)
COMP = np.zeros(256, np.uint8); COMP[[65, 67, 71, 84, 78]] = [84, 71, 67, 65, 78]


def make_case(seed, glen=120000, n_reads=24, k=17, w=10, index_w=10):
    rng = np.random.default_rng(seed)
    g = B[rng.integers(0, 4, glen)].copy()
    g[40000:46000] = g[2000:8000]                       # a 6 kb repeat
    g[70000:70050] = ord("N")
    g[90000:92000] = np.tile(B[rng.integers(0, 4, 5)], 400)   # low-complexity stretch: equal k-mers, ties everywhere
    gt, gp = po.sort_minimizers(*po.store_minimizers(g, k, index_w, "port"), "port")
    reads = []
    for i in range(n_reads):
        L = int(rng.choice([30, 300, 2000, 6000]))
        a = int(rng.integers(0, glen - L - 1))
        r = g[a:a + L].copy()
        m = rng.random(L) < float(rng.choice([0.0, 0.02, 0.1]))
        r[m] = B[rng.integers(0, 4, int(m.sum()))]
        if rng.random() < 0.5:
            r = COMP[r[::-1]]
        if i % 7 == 3:                                    # read carrying both strands of the same stretch (palindromic keys)
            r = np.concatenate([r, COMP[r[::-1]]])
        if i % 5 == 2:
            r[len(r) // 3: len(r) // 3 + 7] = ord("N")
        reads.append(r)
    read_len = np.array([len(r) for r in reads], np.uint32)
    read_off = np.zeros(n_reads, np.uint64); read_off[1:] = np.cumsum(read_len[:-1])
    arena = np.concatenate(reads + [np.full(16, ord("N"), np.uint8)])
    genome = np.concatenate([g, np.full(16, ord("N"), np.uint8)])
    return dict(genome=genome, idx_t=gt, idx_pos=gp, reads=reads, arena=arena, read_off=read_off, read_len=read_len, k=k, w=w)


def expected(case, max_freq, which="port"):
    out = []
    for r in case["reads"]:
        out.append(po.seed_read(r, case["genome"], case["idx_t"], case["idx_pos"], case["k"], case["w"], max_freq, which))
    return out
