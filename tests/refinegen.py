"""Seeded inputs for the local-refinement (a12 LocalIndex::IndexSeq, a13 REFINEclusters) parity tests: a genome of three contigs
(a repeat, an N run, a low-complexity stretch), its global minimizer index (oracle) and LocalIndex, reads drawn from it on both
strands with substitutions and indels, and per read the clusters a repeat-free run of the reference would hand to
REFINEclusters: the seeds of one strand on one contig (plus a few degenerate clusters: empty, across two contigs, at contig ends)."""
import sys, os
import numpy as np
from oracle import pyoracle as po
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
import synth

B = np.frombuffer(b"ACGT", np.uint8)
COMP = np.zeros(256, np.uint8); COMP[[65, 67, 71, 84, 78]] = [84, 71, 67, 65, 78]


def make_case(seed, n_reads=10, contig_lens=(50000, 30011, 41000), k=17, w=10, err=(0.0, 0.01, 0.08), max_freq=50, read_lens=(900, 3000, 7000, 12000)):
    rng = np.random.default_rng(seed)
    contigs = [B[rng.integers(0, 4, n)].copy() for n in contig_lens]
    c0 = contigs[0]
    n0 = len(c0)
    c0[2 * n0 // 5:2 * n0 // 5 + 3000] = c0[1000:4000]            # repeat inside contig 0
    c0[3 * n0 // 5:3 * n0 // 5 + 40] = ord("N")
    c0[7 * n0 // 10:7 * n0 // 10 + 1500] = np.tile(B[rng.integers(0, 4, 3)], 500)      # low complexity: equal local tuples, sort ties, frequency filter
    g = np.concatenate(contigs)
    hdr = np.zeros(len(contigs) + 1, np.uint64); hdr[1:] = np.cumsum(contig_lens)
    gt, gp = po.sort_minimizers(*po.store_minimizers(g, k, w, "port"), "port")
    reads, clusters = [], []
    for i in range(n_reads):
        ci = int(rng.integers(0, len(contigs)))
        L = int(rng.choice(read_lens)); L = min(L, contig_lens[ci] - 2)
        a = int(rng.integers(0, contig_lens[ci] - L)) if i % 4 else (0 if i % 8 else contig_lens[ci] - L)   # some reads at the contig ends
        r = synth.mutate(contigs[ci][a:a + L].copy(), float(rng.choice(err)), rng)
        r[r == ord("N")] = ord("A") if i % 3 else ord("N")
        strand = int(rng.random() < 0.5)
        if strand:
            r = COMP[r[::-1]]
        reads.append(r)
        qt, qpos, tt, tpos, st = po.seed_read(r, np.concatenate([g, np.full(16, ord("N"), np.uint8)]), gt, gp, k, w, max_freq, "port")
        for s in (0, 1):
            sel = st == s
            q, t = qpos[sel], tpos[sel]
            # one cluster per contig the seeds of this strand fall into
            for cj in range(len(contigs)):
                inc = (t >= hdr[cj]) & (t + k <= hdr[cj + 1])
                if inc.sum() == 0:
                    continue
                mq, mt = q[inc].astype(np.uint32), t[inc].astype(np.uint32)
                box = np.array([mq.min(), mq.max() + k, mt.min(), mt.max() + k], np.uint32)
                clusters.append(dict(read=i, strand=s, mq=mq, mt=mt, box=box))
        if i == 1:
            clusters.append(dict(read=i, strand=0, mq=np.zeros(0, np.uint32), mt=np.zeros(0, np.uint32), box=np.zeros(4, np.uint32)))
        if i == 2:   # a cluster across two contigs: dropped by CHROMIndex
            mq = np.array([5, 40], np.uint32); mt = np.array([hdr[1] - 100, hdr[1] + 30], np.uint32)
            clusters.append(dict(read=i, strand=0, mq=mq, mt=mt, box=np.array([5, 40 + k, mt[0], mt[1] + k], np.uint32)))
    return dict(contigs=contigs, genome=g, hdr=hdr, reads=reads, clusters=clusters, k=k)


def expected(case, which="port", small_k=10, window=100, local_max_freq=15):
    gl = po.local_index(case["contigs"], max_freq=local_max_freq, which="port")
    out = []
    handles = {}
    glh = po.RefLocalIndexHandle(case["contigs"], max_freq=local_max_freq) if which == "ref" else None
    for c in case["clusters"]:
        r = case["reads"][c["read"]]
        rc = COMP[r[::-1]]
        if which == "ref":
            if c["read"] not in handles:
                handles[c["read"]] = (po.RefLocalIndexHandle(r, max_freq=local_max_freq), po.RefLocalIndexHandle(rc, max_freq=local_max_freq))
            f, v = handles[c["read"]]
            out.append(po.refine_cluster(c["mq"], c["mt"], c["box"], c["strand"], len(r), case["hdr"], None, None, None, case["k"], small_k, window,
                                         local_max_freq, "ref", (glh.h, f.h, v.h)))
        else:
            rf = po.local_index(r, max_freq=local_max_freq); rr = po.local_index(rc, max_freq=local_max_freq)
            out.append(po.refine_cluster(c["mq"], c["mt"], c["box"], c["strand"], len(r), case["hdr"], gl, rf, rr, case["k"], small_k, window,
                                         local_max_freq, "port"))
    for f, v in handles.values():
        f.close(); v.close()
    if glh:
        glh.close()
    return out


def same(a, b):
    if a["status"] != b["status"]:
        return False
    if a["status"] != 0:
        return True
    return (a["chrom"] == b["chrom"] and (a["mq"] == b["mq"]).all() and (a["mt"] == b["mt"]).all() and (a["box"] == b["box"]).all()
            and (a["diag"] == b["diag"]).all() and len(a["rq"]) == len(b["rq"]) and (a["rq"] == b["rq"]).all() and (a["rt"] == b["rt"]).all()
            and (a["rtup"] == b["rtup"]).all() and (len(a["rq"]) == 0 or ((a["rbox"] == b["rbox"]).all() and a["eff"] == b["eff"])))


def pack_case(case, small_k=10, window=100, local_max_freq=15):
    """Batch layout of the C ABI: read arena (forward strands, 16 bytes of padding), per-read offsets, and the cluster batch."""
    reads = case["reads"]
    read_len = np.array([len(r) for r in reads], np.uint32)
    read_off = np.zeros(len(reads), np.uint64)
    pos = 0
    for i, r in enumerate(reads):       # reads do not need to be contiguous in the arena: leave odd gaps
        read_off[i] = pos; pos += len(r) + (i % 3) * 5
    arena = np.full(pos + 16, ord("N"), np.uint8)
    rc_arena = arena.copy()
    for i, r in enumerate(reads):
        arena[int(read_off[i]):int(read_off[i]) + len(r)] = r
        rc_arena[int(read_off[i]):int(read_off[i]) + len(r)] = COMP[r[::-1]]
    cls = case["clusters"]
    m_off = np.zeros(len(cls) + 1, np.uint64); m_off[1:] = np.cumsum([len(c["mq"]) for c in cls])
    cat = lambda key: np.concatenate([c[key] for c in cls]).astype(np.uint32) if len(cls) else np.zeros(0, np.uint32)
    cl = dict(m_q=cat("mq"), m_t=cat("mt"), m_off=m_off, box=np.array([c["box"] for c in cls], np.uint32).reshape(-1, 4),
              strand=np.array([c["strand"] for c in cls], np.uint8), read_id=np.array([c["read"] for c in cls], np.uint32), hdr_pos=case["hdr"],
              global_k=case["k"], small_k=small_k, window=window, local_max_freq=local_max_freq)
    genome = np.concatenate([case["genome"], np.full(16, ord("N"), np.uint8)])
    return dict(arena=arena, rc_arena=rc_arena, read_off=read_off, read_len=read_len, genome=genome, cl=cl)


def check_batch(o, cl, exp):
    """o: result dict of the emulator / C ABI; exp: list from expected()."""
    for c, e in enumerate(exp):
        assert o["status"][c] == e["status"], (c, o["status"][c], e["status"])
        if e["status"] != 0:
            assert o["r_off"][c + 1] == o["r_off"][c], c
            continue
        a, b = int(o["r_off"][c]), int(o["r_off"][c + 1])
        m0, m1 = int(cl["m_off"][c]), int(cl["m_off"][c + 1])
        assert o["chrom"][c] == e["chrom"], c
        assert (o["diag"][2 * c:2 * c + 2] == e["diag"]).all(), c
        assert (o["m_q_out"][m0:m1] == e["mq"]).all() and (o["m_t_out"][m0:m1] == e["mt"]).all(), c
        assert (o["box_out"].reshape(-1, 4)[c] == e["box"]).all(), c
        assert b - a == len(e["rq"]), (c, b - a, len(e["rq"]))
        assert (o["r_q"][a:b] == e["rq"]).all() and (o["r_t"][a:b] == e["rt"]).all() and (o["r_tup"][a:b] == e["rtup"]).all(), c
        if b > a:
            assert (o["rbox"].reshape(-1, 4)[c] == e["rbox"]).all(), c
            assert o["eff"][c] == e["eff"], (c, o["eff"][c], e["eff"])


# ---- Refine_splitchain (the low-accuracy pipeline's a13)

def make_chains(case, seed=0):
    """Split chains for the clusters of make_case(): the anchors of a cluster in chain order (ascending genome position; every fifth
    chain keeps an unsorted order -- the reference scans anchors sequentially), anchor lengths 17..60, boundaries from the anchors.
    A reverse chain (Strand 1) takes its anchors from strand-1 clusters; every seventh chain mixes in anchors of the other strand."""
    rng = np.random.default_rng(seed)
    hdr = case["hdr"]
    chains = []
    for ci, c in enumerate(case["clusters"]):
        n = len(c["mq"])
        if n == 0:
            chains.append(dict(read=c["read"], strand=0, chrom=0, mq=c["mq"], mt=c["mt"], mlen=np.zeros(0, np.uint32), mstrand=np.zeros(0, np.uint8), box=np.zeros(4, np.uint32)))
            continue
        chrom = int(np.searchsorted(hdr, c["mt"].min(), side="right") - 1)
        if int(np.searchsorted(hdr, c["mt"].max() + 60, side="right") - 1) != chrom:
            continue                                                  # a split chain lives on one contig
        order = np.argsort(c["mt"], kind="stable") if ci % 5 else rng.permutation(n)
        mq, mt = c["mq"][order], c["mt"][order]
        mlen = rng.integers(17, 61, n).astype(np.uint32)
        mstrand = np.full(n, c["strand"], np.uint8)
        if ci % 7 == 3:
            mstrand[rng.random(n) < 0.1] ^= 1
        L = len(case["reads"][c["read"]])
        mlen = np.minimum(mlen, np.maximum(17, L - mq.astype(np.int64) - 1)).astype(np.uint32)
        box = np.array([mq.min(), (mq + mlen).max(), mt.min(), (mt + mlen).max()], np.uint32)
        chains.append(dict(read=c["read"], strand=c["strand"], chrom=chrom, mq=mq, mt=mt, mlen=mlen, mstrand=mstrand, box=box))
    return chains


def expected_chains(case, chains, which="port", small_k=10, window=100, local_max_freq=15, limitrefine=1):
    gl = po.local_index(case["contigs"], max_freq=local_max_freq, which="port")
    glh = po.RefLocalIndexHandle(case["contigs"], max_freq=local_max_freq) if which == "ref" else None
    out, handles = [], {}
    for c in chains:
        r = case["reads"][c["read"]]; rc = COMP[r[::-1]]
        if which == "ref":
            if c["read"] not in handles:
                handles[c["read"]] = (po.RefLocalIndexHandle(r, max_freq=local_max_freq), po.RefLocalIndexHandle(rc, max_freq=local_max_freq))
            f, v = handles[c["read"]]
            out.append(po.refine_splitchain(c["mq"], c["mt"], c["mlen"], c["mstrand"], c["box"], c["chrom"], c["strand"], len(r), case["hdr"], None, None, None,
                                            case["k"], small_k, window, local_max_freq, limitrefine, "ref", (glh.h, f.h, v.h)))
        else:
            rf = po.local_index(r, max_freq=local_max_freq); rr = po.local_index(rc, max_freq=local_max_freq)
            out.append(po.refine_splitchain(c["mq"], c["mt"], c["mlen"], c["mstrand"], c["box"], c["chrom"], c["strand"], len(r), case["hdr"], gl, rf, rr,
                                            case["k"], small_k, window, local_max_freq, limitrefine, "port"))
    for f, v in handles.values():
        f.close(); v.close()
    if glh:
        glh.close()
    return out


def same_chain(a, b):
    if a["status"] != b["status"]:
        return False
    if a["status"] != 0:
        return True
    return (a["chrom"] == b["chrom"] and (a["diag"] == b["diag"]).all() and len(a["rq"]) == len(b["rq"]) and (a["rq"] == b["rq"]).all()
            and (a["rt"] == b["rt"]).all() and (a["rtup"] == b["rtup"]).all() and (len(a["rq"]) == 0 or ((a["rbox"] == b["rbox"]).all() and a["eff"] == b["eff"])))


def pack_chains(case, chains, small_k=10, window=100, local_max_freq=15, limitrefine=1):
    """The split-chain batch of the C ABI (lra_b200_splitchains) on the arenas of pack_case()."""
    pk = pack_case(case, small_k, window, local_max_freq)
    m_off = np.zeros(len(chains) + 1, np.uint64); m_off[1:] = np.cumsum([len(c["mq"]) for c in chains])
    cat = lambda key, dt: np.concatenate([c[key] for c in chains]).astype(dt) if len(chains) else np.zeros(0, dt)
    pk["cl"] = dict(m_q=cat("mq", np.uint32), m_t=cat("mt", np.uint32), m_len=cat("mlen", np.uint32), m_strand=cat("mstrand", np.uint8), m_off=m_off,
                    box=np.array([c["box"] for c in chains], np.uint32).reshape(-1, 4), strand=np.array([c["strand"] for c in chains], np.uint8),
                    chrom=np.array([c["chrom"] for c in chains], np.int32), read_id=np.array([c["read"] for c in chains], np.uint32), hdr_pos=case["hdr"],
                    global_k=case["k"], small_k=small_k, window=window, local_max_freq=local_max_freq, limitrefine=limitrefine)
    return pk


def check_chain_batch(o, exp):
    for c, e in enumerate(exp):
        assert o["status"][c] == e["status"], (c, o["status"][c], e["status"])
        a, b = int(o["r_off"][c]), int(o["r_off"][c + 1])
        if e["status"] != 0:
            assert a == b, c
            continue
        assert o["chrom"][c] == e["chrom"] and (o["diag"][2 * c:2 * c + 2] == e["diag"]).all(), c
        assert b - a == len(e["rq"]), (c, b - a, len(e["rq"]))
        assert (o["r_q"][a:b] == e["rq"]).all() and (o["r_t"][a:b] == e["rt"]).all() and (o["r_tup"][a:b] == e["rtup"]).all(), c
        if b > a:
            assert (o["rbox"].reshape(-1, 4)[c] == e["rbox"]).all() and o["eff"][c] == e["eff"], c
