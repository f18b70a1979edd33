"""The MapRead_lowacc pipeline on the CPU for the parity tests (TEST INFRASTRUCTURE): the mapper worker kernel and the finalize kernel of
lra_b200/csrc/mp_*.cuh run under the SIMT emulator (tests/emu_mp.py); the stages between them that have their own kernels and their own
parity tests -- a12 LocalIndex::IndexSeq of the reads, a19 IndelRefineAlignment, a21 CalculateStatistics -- are stood in for by the
reference itself (oracle/_ref/libref_lra.so).  The SAM text comes from the product's host emitter (lra_b200_format_sam).
On a GPU the whole path is lra_b200_map_batch; this module exists so that the glue can be diffed against `lra_ref align` without one."""
import ctypes as C
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)
import emu_mp  # noqa: E402
import lra_b200  # noqa: E402
from lra_b200 import capi  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

COMP = np.full(256, ord("N"), np.uint8)
for a, b in zip(b"ACGTacgtn", b"TGCAtgcan"):
    COMP[a] = b

_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")

SEGREC = np.dtype([("read", "<i4"), ("chain", "<i4"), ("order_in_chain", "<i4"), ("strand", "<i4"), ("chrom", "<i4"), ("n0", "<i4"), ("n1", "<i4"),
                   ("supp", "<i4"), ("issec", "<i4"), ("first", "<f4"), ("blk_off", "<u8"), ("blk_cnt", "<i4"), ("_pad", "<i4")])


def _bind(L):
    if getattr(L, "_map_bound", False):
        return
    L.emu_map_reads.restype = C.c_int
    L.emu_map_reads.argtypes = [_u8p, _u8p, C.c_uint64, _u64p, _u32p, C.c_int, _u8p, C.c_uint64, _u64p, C.c_int, _u64p, _u32p, C.c_int64,
                                _u64p, _u64p, _u32p, C.c_int, _u32p, _u64p, _u64p, _u32p, _u32p, _u64p, _u64p, _u32p, C.c_void_p, _i64p, _f32p, _f32p, C.c_int, C.c_int,
                                _i32p, _i32p, _i32p, _i32p, C.c_void_p, C.c_int, _u32p, C.c_uint64, _u64p, C.c_uint64]
    L.emu_map_finalize.restype = C.c_int
    L.emu_map_finalize.argtypes = [C.c_int, C.c_void_p, _u64p, _u32p, _i32p, _i32p, _i32p, _i32p, C.c_void_p, _i32p, _u64p, _u32p, _i32p, _f32p, _u64p, _f32p,
                                   C.c_void_p, _i32p, _u64p, _f32p, C.c_void_p]
    L._map_bound = True


def load_inputs(w):
    """reads, genome, indexes and options of a tests/mapgen.workdir."""
    ref = w["ref_records"]; reads = w["read_records"]
    genome = np.concatenate([s for _, s in ref])
    hdr = np.zeros(len(ref) + 1, np.uint64); hdr[1:] = np.cumsum([len(s) for _, s in ref])
    mms = capi.read_mms(w["ref"] + ".mms")
    gli = capi.read_gli(w["ref"] + ".gli")
    opts = capi.map_opts_preset(w["preset"])
    opts.globalK = mms["k"]; opts.smallK = gli["k"]; opts.smallW = gli["w"]; opts.localIndexWindow = gli["window"]
    rl = np.array([len(s) for _, s in reads], np.uint32)
    ro = np.zeros(len(reads), np.uint64); ro[1:] = np.cumsum(rl[:-1].astype(np.uint64))
    fwd = np.concatenate([s for _, s in reads])
    return dict(genome=genome, hdr=hdr, contig_names=[n for n, _ in ref], mms=mms, gli=gli, opts=opts, reads=fwd, read_off=ro, read_len=rl, names=[n for n, _ in reads])


def read_local_indexes(inp):
    """LocalIndex::IndexSeq of every read, both strands, through the reference (stand-in for the a12 kernel)."""
    out = []
    rcs = []
    for strand in (0, 1):
        wf = np.zeros(len(inp["read_len"]) + 1, np.uint32); offs, bnds, mins = [], [], []
        nw = 0; nm = 0
        for r, (o, l) in enumerate(zip(inp["read_off"], inp["read_len"])):
            s = inp["reads"][int(o):int(o) + int(l)]
            if strand:
                s = COMP[s[::-1]]
                rcs.append(s)
            li = po.local_index(s, k=inp["opts"].smallK, w=inp["opts"].smallW, window=inp["opts"].localIndexWindow, max_freq=inp["opts"].localIndexMaxFreq, which="ref")
            n = len(li.seq_off) - 1
            offs.append(li.seq_off[:-1] + np.uint64(o)); bnds.append(li.bnd[:-1] + np.uint64(nm)); mins.append(li.mins)
            nw += n; nm += len(li.mins); wf[r + 1] = nw
        win_off = np.concatenate(offs + [np.array([inp["read_off"][-1] + np.uint64(inp["read_len"][-1])], np.uint64)])
        bnd = np.concatenate(bnds + [np.array([nm], np.uint64)])
        out.append(dict(win_first=wf, win_off=np.ascontiguousarray(win_off), bnd=np.ascontiguousarray(bnd), mins=np.ascontiguousarray(np.concatenate(mins + [np.zeros(1, np.uint32)]))))
    return out[0], out[1], np.concatenate(rcs)


def map_reads(inp, lanes=1, arena_bytes=1 << 30):
    L = emu_mp.lib(lanes)
    _bind(L)
    n = len(inp["read_len"])
    rf, rr, rc = read_local_indexes(inp)
    o = inp["opts"]
    pwl = lra_b200.init_pwl(o.gapopen, o.gapextend, o.gaproot, o.gapCeiling1, o.gapCeiling2)
    seg_cap = 8 * n + 64; blk_cap = int(inp["read_len"].sum()) + 64 * n + 1024
    out = dict(status=np.zeros(n, np.int32), n_chains=np.zeros(n, np.int32), chain_nseg=np.zeros(4 * n, np.int32), chain_seg0=np.zeros(4 * n, np.int32),
               seg=np.zeros(seg_cap, SEGREC), blocks=np.zeros(3 * blk_cap, np.uint32), counts=np.zeros(4, np.uint64))
    assert L.emu_sizeof_segrec() == SEGREC.itemsize
    gli = inp["gli"]
    err = L.emu_map_reads(inp["reads"], rc, len(inp["reads"]), inp["read_off"], inp["read_len"], n, inp["genome"], len(inp["genome"]), inp["hdr"], len(inp["hdr"]) - 1,
                          inp["mms"]["t"], inp["mms"]["pos"], len(inp["mms"]["t"]), gli["seq_offsets"], gli["tuple_boundaries"], gli["minimizers"], len(gli["seq_offsets"]) - 1,
                          rf["win_first"], rf["win_off"], rf["bnd"], rf["mins"], rr["win_first"], rr["win_off"], rr["bnd"], rr["mins"],
                          C.addressof(o), pwl[0], pwl[1], pwl[2], pwl[3], pwl[4], out["status"], out["n_chains"], out["chain_nseg"], out["chain_seg0"],
                          out["seg"].ctypes.data, seg_cap, out["blocks"], blk_cap, out["counts"], arena_bytes)
    out["err"] = err; out["rc"] = rc
    out["n_seg"] = int(out["counts"][0]); out["n_blk"] = int(out["counts"][1]); out["peak"] = int(out["counts"][3])
    return out


def refine_and_stats(inp, mo):
    """IndelRefineAlignment + CalculateStatistics of every segment through the reference (stand-ins for the a19 / a21 kernels)."""
    S = mo["n_seg"]; seg = mo["seg"][:S]; o = inp["opts"]
    N = len(inp["reads"])
    q_arena = np.concatenate([inp["reads"], mo["rc"], np.zeros(64, np.uint8)])
    hdr = inp["hdr"]
    sb = dict(q_arena=q_arena, blocks_in=mo["blocks"][:3 * mo["n_blk"]].reshape(-1, 3), blk_off=seg["blk_off"].astype(np.uint64), blk_cnt=seg["blk_cnt"].astype(np.int32),
              q_base=(inp["read_off"][seg["read"]] + np.uint64(N) * seg["strand"].astype(np.uint64)).astype(np.uint32),
              read_len=inp["read_len"][seg["read"]].astype(np.int32), contig_len=(hdr[seg["chrom"] + 1] - hdr[seg["chrom"]]).astype(np.int32),
              k=o.refineBand, match=o.localMatch, mismatch=o.localMismatch, indel=o.localIndel, end_align=1 if o.HighlyAccurate else 0)
    t_base = hdr[seg["chrom"]].astype(np.uint32)
    g_arena = np.concatenate([inp["genome"], np.zeros(64, np.uint8)])
    if S == 0:
        return dict(ir_n=np.zeros(1, np.int32), ir_off=np.zeros(1, np.uint64), ir_blocks=np.zeros(3, np.uint32), stats=np.zeros(16, np.int32), value=np.zeros(1, np.float32),
                    cigar_off=np.zeros(2, np.uint64), cigar=np.zeros(1, np.uint32))
    n, off, blocks = po.indel_refine_batch_ref(sb, g_arena, t_base, nthreads=4)
    bl = [blocks[int(off[s]):int(off[s]) + int(n[s])].copy() for s in range(S)]
    opmap = {"=": 7, "X": 8, "I": 1, "D": 2, "M": 0}
    import re

    def seqs(s):
        qb = int(sb["q_base"][s]); rl = int(sb["read_len"][s]); tb = int(t_base[s]); cl = int(sb["contig_len"][s])
        return q_arena[qb:qb + rl], g_arena[tb:tb + cl], rl

    def all_stats():
        stats = np.zeros(16 * S, np.int32); value = np.zeros(S, np.float32); cig_off = np.zeros(S + 1, np.uint64); cigs = []
        for s in range(S):
            rd, gn, _ = seqs(s)
            st, v, cig = po.calc_stats_ref(rd.tobytes(), gn.tobytes(), bl[s])
            stats[16 * s:16 * s + 16] = st; value[s] = v
            ops = np.array([(int(l) << 4) | opmap[c] for l, c in re.findall(r"(\d+)([=XIDM])", cig)], np.uint32)
            cigs.append(ops); cig_off[s + 1] = cig_off[s] + np.uint64(len(ops))
        return stats, value, cig_off, cigs

    stats, value, cig_off, cigs = all_stats()
    extra = {}
    if o.HighlyAccurate:
        # Map_highacc.h:722-731: RefineBreakpoint between consecutive segments of every alignment, then CalculateStatistics again (stand-in: the reference)
        extra["stats_first"] = stats
        nrd = len(inp["read_len"])
        for r in range(nrd):
            for a in range(int(mo["n_chains"][r])):
                ns = int(mo["chain_nseg"][4 * r + a]); s0 = int(mo["chain_seg0"][4 * r + a])
                for k in range(1, ns):
                    li, ri = s0 + k, s0 + k - 1
                    lrd, lgn, rl = seqs(li); rrd, rgn, _ = seqs(ri)
                    if len(bl[li]) == 0 or len(bl[ri]) == 0:
                        continue
                    bl[li], bl[ri] = po.refine_breakpoint_ref(lrd, rrd, rl, lgn, rgn, bl[li], int(seg["strand"][li]), bl[ri], int(seg["strand"][ri]))
        stats, value, cig_off, cigs = all_stats()
    n = np.array([len(b) for b in bl], np.int32); off = np.zeros(S, np.uint64); off[1:] = np.cumsum(n[:-1].astype(np.uint64))
    blocks = np.concatenate([b.reshape(-1, 3) for b in bl] + [np.zeros((1, 3), np.uint32)])
    return dict(ir_n=n, ir_off=off, ir_blocks=np.ascontiguousarray(blocks.reshape(-1)), stats=stats, value=value, cigar_off=cig_off,
                cigar=np.ascontiguousarray(np.concatenate(cigs + [np.zeros(1, np.uint32)])), **extra)


def finalize(inp, mo, rs, lanes=1):
    L = emu_mp.lib(lanes)
    _bind(L)
    n = len(inp["read_len"]); S = mo["n_seg"]
    rec = np.zeros(max(S, 1), capi.RECORD)
    assert L.emu_sizeof_record() == capi.RECORD.itemsize
    rank = np.zeros(4 * n, np.int32); ab = np.zeros(1, np.uint64)
    logf_len = np.array([0.0] + [np.log(np.float32(i)) for i in range(1, 8)], np.float32)
    libm = C.CDLL("libm.so.6"); libm.logf.restype = C.c_float; libm.logf.argtypes = [C.c_float]
    K = np.float32(inp["opts"].globalK)
    seg_l = np.array([libm.logf(np.float32(v) / K) if v > 3 else 0.0 for v in rs["value"][:max(S, 1)]], np.float32)
    L.emu_map_finalize(n, C.addressof(inp["opts"]), inp["read_off"], inp["read_len"], mo["status"], mo["n_chains"], mo["chain_nseg"], mo["chain_seg0"], mo["seg"].ctypes.data,
                       rs["ir_n"], rs["ir_off"], rs["ir_blocks"], rs["stats"], rs["value"], rs["cigar_off"], logf_len, rec.ctypes.data, rank, ab, seg_l,
                       rs.get("stats_first", rs["stats"]).ctypes.data if inp["opts"].HighlyAccurate else None)
    return dict(status=mo["status"], n_aln=mo["n_chains"], aln_nseg=mo["chain_nseg"], aln_seg0=mo["chain_seg0"], aln_rank=rank, records=rec, n_records=S, cigar=rs["cigar"],
                n_cigar=int(rs["cigar_off"][-1]), aligned_bases=int(ab[0]))


def sam_text(inp, res, runtime=0):
    return capi.format_sam(inp["opts"], res, inp["names"], inp["reads"], inp["read_off"], inp["read_len"], inp["contig_names"], runtime)


def run(w, lanes=1):
    inp = load_inputs(w)
    mo = map_reads(inp, lanes)
    rs = refine_and_stats(inp, mo)
    res = finalize(inp, mo, rs, lanes)
    return inp, mo, res, sam_text(inp, res)
