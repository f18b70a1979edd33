"""SparseDP problems for the a10 parity tests (TEST INFRASTRUCTURE): parser of the $LRA_CAPTURE_SDP stream written by
oracle/_ref/lra_capture (inputs and outputs of every SparseDP / SparseDP_ForwardOnly call of a real `lra align` run),
seeded random problems, and the packing into the batch layout of lra_b200_sdp_batch."""
import numpy as np


def parse_capture(path, limit=None):
    a = np.fromfile(path, dtype=np.uint32)
    f = a.view(np.float32)
    i = 0
    out = []
    while i < len(a) and (limit is None or len(out) < limit):
        kind = int(a[i])
        if kind == 0:
            n_cl, nfrag = int(a[i + 1]), int(a[i + 2])
            rec = dict(kind=0, rate=float(f[i + 3]), alnthres=float(f[i + 4]), NumAln=int(a[i + 5]), read_len=int(a[i + 6]))
            i += 7
            rec["cl_off"] = a[i:i + n_cl + 1].astype(np.int32); i += n_cl + 1
            rec["cl_strand"] = a[i:i + n_cl].astype(np.uint8); i += n_cl
            rec["q"] = a[i:i + nfrag].copy(); i += nfrag
            rec["t"] = a[i:i + nfrag].copy(); i += nfrag
            rec["len"] = a[i:i + nfrag].astype(np.int32); i += nfrag
            nch = int(a[i]); i += 1
            chains = []
            for _ in range(nch):
                n = int(a[i]); val = f[i + 1]; b = a[i + 2:i + 6].copy(); i += 6
                ch = a[i:i + n].copy(); i += n
                lk = a[i:i + max(0, n - 1)].astype(np.uint8); i += max(0, n - 1)
                chains.append(dict(n=n, value=np.float32(val), bounds=b, chain=ch, link=lk))
            rec["chains"] = chains
        elif kind == 1:
            strand, nfrag = int(a[i + 1]), int(a[i + 2])
            rec = dict(kind=1, rate=float(f[i + 3]), cl_off=np.array([0, nfrag], np.int32), cl_strand=np.array([strand], np.uint8))
            i += 4
            rec["q"] = a[i:i + nfrag].copy(); i += nfrag
            rec["t"] = a[i:i + nfrag].copy(); i += nfrag
            rec["len"] = a[i:i + nfrag].astype(np.int32); i += nfrag
            n = int(a[i]); val = f[i + 1]; i += 2
            ch = a[i:i + n].copy(); i += n
            lk = a[i:i + max(0, n - 1)].astype(np.uint8); i += max(0, n - 1)
            rec["chains"] = [dict(n=n, value=np.float32(val), chain=ch, link=lk)]
        elif kind == 2:
            nfrag = int(a[i + 1])
            rec = dict(kind=2, irate=int(a[i + 2]), cl_off=np.array([0, nfrag], np.int32), cl_strand=np.array([0], np.uint8))
            i += 3
            rec["q"] = a[i:i + nfrag].copy(); i += nfrag
            rec["t"] = a[i:i + nfrag].copy(); i += nfrag
            rec["len"] = a[i:i + nfrag].astype(np.int32); i += nfrag
            n = int(a[i]); val = f[i + 1]; i += 2
            ch = a[i:i + n].copy(); i += n
            rec["chains"] = [dict(n=n, value=np.float32(val), chain=ch, link=np.zeros(0, np.uint8))]
        elif kind == 3:
            n_cl, nfrag = int(a[i + 1]), int(a[i + 2])
            rec = dict(kind=3, rate=float(f[i + 3]))
            i += 4
            rec["cl_off"] = a[i:i + n_cl + 1].astype(np.int32); i += n_cl + 1
            rec["cl_strand"] = a[i:i + n_cl].astype(np.uint8); i += n_cl
            rec["q"] = a[i:i + nfrag].copy(); i += nfrag
            rec["t"] = a[i:i + nfrag].copy(); i += nfrag
            rec["len"] = a[i:i + nfrag].astype(np.int32); i += nfrag
            n = int(a[i]); val = f[i + 1]; i += 2
            ch = a[i:i + n].copy(); i += n
            rec["chains"] = [dict(n=n, value=np.float32(val), chain=ch, link=np.zeros(0, np.uint8))]
        elif kind == 4:
            nfrag = int(a[i + 1])
            rec = dict(kind=4, rate=float(f[i + 2]), alnthres=float(f[i + 3]), globalK=int(a[i + 4]), NumAln=int(a[i + 5]), read_len=int(a[i + 6]),
                       cl_off=np.array([0, nfrag], np.int32), cl_strand=np.array([0], np.uint8))
            i += 7
            rec["q"] = a[i:i + nfrag].copy(); i += nfrag
            rec["qe"] = a[i:i + nfrag].copy(); i += nfrag
            rec["t"] = a[i:i + nfrag].copy(); i += nfrag
            rec["te"] = a[i:i + nfrag].copy(); i += nfrag
            rec["fstrand"] = a[i:i + nfrag].astype(np.uint8); i += nfrag
            rec["fval"] = f[i:i + nfrag].copy(); i += nfrag
            rec["fn0"] = a[i:i + nfrag].astype(np.int32); i += nfrag
            rec["len"] = np.zeros(nfrag, np.int32)
            nch = int(a[i]); i += 1
            chains = []
            for _ in range(nch):
                n = int(a[i]); val = f[i + 1]; b = a[i + 2:i + 6].copy(); n0 = int(a[i + 6]); i += 7
                ch = a[i:i + n].copy(); i += n
                lk = a[i:i + max(0, n - 1)].astype(np.uint8); i += max(0, n - 1)
                chains.append(dict(n=n, value=np.float32(val), bounds=b, n0=n0, chain=ch, link=lk))
            rec["chains"] = chains
        else:
            raise ValueError("bad capture stream at word %d" % i)
        out.append(rec)
    return out


def pack(recs):
    n = len(recs)
    frag_off = np.zeros(n + 1, np.uint64)
    cl_off_off = np.zeros(n + 1, np.uint64)
    for k, r in enumerate(recs):
        frag_off[k + 1] = frag_off[k] + len(r["q"])
        cl_off_off[k + 1] = cl_off_off[k] + len(r["cl_off"])
    cat = lambda key, dt: np.ascontiguousarray(np.concatenate([np.asarray(r[key], dt) for r in recs])) if n else np.zeros(0, dt)
    ext = {}
    if any(r["kind"] == 4 for r in recs):
        z = lambda r, key, dt: np.asarray(r[key], dt) if key in r else np.zeros(len(r["q"]), dt)
        for key, dt in (("qe", np.uint32), ("te", np.uint32), ("fstrand", np.uint8), ("fval", np.float32), ("fn0", np.int32)):
            ext[key] = np.ascontiguousarray(np.concatenate([z(r, key, dt) for r in recs]))
        ext["globalK"] = int([r for r in recs if r["kind"] == 4][0]["globalK"])
    return dict(ext, mode=np.array([r["kind"] for r in recs], np.int32), frag_off=frag_off, q=cat("q", np.uint32), t=cat("t", np.uint32), len=cat("len", np.int32),
                cl_off_off=cl_off_off, cl_off=cat("cl_off", np.int32),
                cl_strand=np.ascontiguousarray(np.concatenate([np.append(np.asarray(r["cl_strand"], np.uint8), 0) for r in recs]).astype(np.uint8)),
                only_cl=np.zeros(n, np.int32), rate=np.array([r.get("rate", 0.0) for r in recs], np.float32),
                irate=np.array([r.get("irate", 0) for r in recs], np.int32), read_len=np.array([r.get("read_len", 1000) for r in recs], np.int32))


def compare(recs, pb, out, max_aln):
    """Returns the list of (problem index, message) mismatches: chain indices, link bits, float value bits, bounds."""
    bad = []
    for k, r in enumerate(recs):
        exp = r["chains"]
        fo = int(pb["frag_off"][k]); nf = int(pb["frag_off"][k + 1]) - fo
        nch = int(out["n_chains"][k])
        if r["kind"] not in (0, 4):
            e = exp[0]
            n = int(out["chain_len"][k * max_aln])
            if n != e["n"]: bad.append((k, "len %d != %d" % (n, e["n"]))); continue
            base = max_aln * fo
            if not (out["chain"][base:base + n] == e["chain"]).all(): bad.append((k, "chain")); continue
            if r["kind"] == 1 and not (out["link"][base:base + n - 1] == e["link"]).all(): bad.append((k, "link")); continue
            if out["chain_val"][k * max_aln].view(np.uint32) != np.float32(e["value"]).view(np.uint32): bad.append((k, "value")); continue
            continue
        if nch != len(exp): bad.append((k, "n_chains %d != %d" % (nch, len(exp)))); continue
        for c, e in enumerate(exp):
            o = k * max_aln + c
            n = int(out["chain_len"][o])
            if n != e["n"]: bad.append((k, "chain %d len %d != %d" % (c, n, e["n"]))); break
            base = max_aln * fo + c * nf
            if not (out["chain"][base:base + n] == e["chain"]).all(): bad.append((k, "chain %d anchors" % c)); break
            if not (out["link"][base:base + n - 1] == e["link"]).all(): bad.append((k, "chain %d link" % c)); break
            if out["chain_val"][o].view(np.uint32) != np.float32(e["value"]).view(np.uint32): bad.append((k, "chain %d value" % c)); break
            if not (out["bounds"][4 * o:4 * o + 4] == e["bounds"]).all(): bad.append((k, "chain %d bounds" % c)); break
            if r["kind"] == 4 and int(out["n0"][o]) != e["n0"]: bad.append((k, "chain %d NumOfAnchors0" % c)); break
    return bad
