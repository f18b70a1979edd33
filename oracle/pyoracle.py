"""ctypes bindings for the oracle libraries (TEST INFRASTRUCTURE ONLY).

  liboracle_lra.so  -- the plain-C restatement in oracle/*.c ("port")
  libref_lra.so     -- extern "C" wrappers around the UNMODIFIED reference headers ("reference");
                       exists only where /root/reference was present at build time (or was shipped
                       prebuilt in oracle/_ref/ to the GPU box).
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_SRC = "/root/reference"

_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def build(want_ref=None):
    """(Re)build the oracle libraries.  The reference-backed targets are built only when the reference
    sources are present; otherwise the prebuilt files in oracle/_ref/ are used as they are."""
    if want_ref is None:
        want_ref = os.path.isdir(REFERENCE_SRC)
    targets = ["restatement"] + (["ref"] if want_ref else [])
    subprocess.run(["make", "-C", HERE, "-j4"] + targets, check=True, stdout=subprocess.DEVNULL)


def _load(name):
    path = os.path.join(REF_DIR, name)
    if not os.path.exists(path):
        return None
    return C.CDLL(path)


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        if not os.path.exists(os.path.join(REF_DIR, "liboracle_lra.so")):
            build(want_ref=False)
        _port = _load("liboracle_lra.so")
        L = _port
        L.lra_oracle_aog.restype = C.c_int
        L.lra_oracle_aog.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     _u32p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.lra_oracle_aog_batch.restype = C.c_int
        L.lra_oracle_aog_batch.argtypes = [_u8p, _u8p, _u32p, _u32p, _i32p, _i32p, _i32p, C.c_int, C.c_int, C.c_int,
                                           C.c_int, _i32p, _i32p, _i64p, _i32p, _u32p, _i32p]
    return _port


def ref():
    """The reference-backed library, or None if it was never built."""
    global _ref
    if _ref is None:
        L = _load("libref_lra.so")
        if L is None:
            return None
        L.ref_aog.restype = C.c_int
        L.ref_aog.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                              _u32p, C.c_int, C.POINTER(C.c_int)]
        L.ref_aog_batch.restype = C.c_int
        L.ref_aog_batch.argtypes = [_u8p, _u8p, _u32p, _u32p, _i32p, _i32p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    _i32p, _i32p, _i64p, C.c_void_p, C.c_int]
        _ref = L
    return _ref


# ---------------------------------------------------------------- a18 AffineOneGapAlign

def aog_port(q, t, m, mm, indel, k):
    """One job through the C restatement.  Returns (score, blocks[n,3] uint32, status)."""
    cap = min(len(q), len(t)) + 2
    blocks = np.zeros(cap * 3, dtype=np.uint32)
    nb, st = C.c_int(0), C.c_int(0)
    s = port().lra_oracle_aog(bytes(q), len(q), bytes(t), len(t), m, mm, indel, k, blocks, cap, C.byref(nb), C.byref(st))
    return s, blocks[: 3 * min(nb.value, cap)].reshape(-1, 3).copy(), st.value


def aog_ref(q, t, m, mm, indel, k):
    """One job through the real reference header.  Returns (score, blocks[n,3])."""
    cap = min(len(q), len(t)) + 2
    blocks = np.zeros(cap * 3, dtype=np.uint32)
    nb = C.c_int(0)
    s = ref().ref_aog(bytes(q), len(q), bytes(t), len(t), m, mm, indel, k, blocks, cap, C.byref(nb))
    return s, blocks[: 3 * min(nb.value, cap)].reshape(-1, 3).copy()


def block_layout(q_len, t_len):
    """Per-job worst-case block capacity and exclusive prefix (shared by all batch back-ends)."""
    cap = (np.minimum(q_len, t_len) + 1).astype(np.int32)
    off = np.zeros(len(cap), dtype=np.int64)
    np.cumsum(cap[:-1], out=off[1:])
    return cap, off, int(cap.sum())


def aog_batch_port(q_arena, t_arena, q_off, t_off, q_len, t_len, k, m, mm, indel, want_blocks=True):
    n = len(q_off)
    cap, off, total = block_layout(q_len, t_len)
    score = np.zeros(n, np.int32); nb = np.zeros(n, np.int32); st = np.zeros(n, np.int32)
    blocks = np.zeros(max(1, total) * 3, np.uint32)
    port().lra_oracle_aog_batch(q_arena, t_arena, q_off, t_off, q_len, t_len, k, n, m, mm, indel, score, nb, off, cap,
                                blocks, st)
    return score, nb, off, blocks.reshape(-1, 3), st


def aog_batch_ref(q_arena, t_arena, q_off, t_off, q_len, t_len, k, m, mm, indel, nthreads=1, want_blocks=True):
    n = len(q_off)
    cap, off, total = block_layout(q_len, t_len)
    score = np.zeros(n, np.int32); nb = np.zeros(n, np.int32)
    blocks = np.zeros(max(1, total) * 3, np.uint32) if want_blocks else None
    ref().ref_aog_batch(q_arena, t_arena, q_off, t_off, q_len, t_len, k, n, m, mm, indel, score, nb, off,
                        blocks.ctypes.data if want_blocks else None, nthreads)
    return score, nb, off, (blocks.reshape(-1, 3) if want_blocks else None)


# ---------------------------------------------------------------- capture files (oracle/lra_capture.cpp)

def read_aog_capture(path, limit=None):
    """Parse an LRA_CAPTURE_AOG file -> list of dict(q,t,m,mm,indel,k,score,blocks)."""
    data = open(path, "rb").read()
    out, p = [], 0
    while p < len(data) and (limit is None or len(out) < limit):
        h = np.frombuffer(data, dtype=np.int32, count=8, offset=p); p += 32
        ql, tl, m, mm, indel, k, score, nb = (int(x) for x in h)
        q = data[p:p + ql]; p += ql
        t = data[p:p + tl]; p += tl
        b = np.frombuffer(data, dtype=np.uint32, count=3 * nb, offset=p).reshape(-1, 3).copy(); p += 12 * nb
        out.append(dict(q=q, t=t, m=m, mm=mm, indel=indel, k=k, score=score, blocks=b))
    return out


# ---------------------------------------------------------------- a19 IndelRefineAlignment

def _bind_ir(L):
    if getattr(L, "_ir_bound", False):
        return
    L.lra_oracle_indel_refine.restype = C.c_int
    L.lra_oracle_indel_refine.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_long, C.c_long, _u32p, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_int, _u32p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.POINTER(C.c_long)]
    L._ir_bound = True


def indel_refine_port(read, twin, t_win_off, contig_len, blocks_in, k, match, mismatch, indel, end_align):
    """One segment through the C restatement.  Returns (blocks_out[n,3], status, cells)."""
    L = port(); _bind_ir(L)
    bi = np.ascontiguousarray(blocks_in, dtype=np.uint32).reshape(-1)
    cap = len(read) + 8
    out = np.zeros(cap * 3, np.uint32)
    n, st, cells = C.c_int(0), C.c_int(0), C.c_long(0)
    L.lra_oracle_indel_refine(bytes(read), len(read), bytes(twin), t_win_off, contig_len, bi, len(bi) // 3, k, match, mismatch,
                              indel, 1 if end_align else 0, out, cap, C.byref(n), C.byref(st), C.byref(cells))
    return out[: 3 * n.value].reshape(-1, 3).copy(), st.value, cells.value


def read_ir_capture(path, limit=None):
    """Parse an LRA_CAPTURE_IR file (oracle/lra_capture.cpp) -> list of dicts."""
    data = open(path, "rb").read()
    out, p = [], 0
    while p < len(data) and (limit is None or len(out) < limit):
        h = np.frombuffer(data, dtype=np.int32, count=12, offset=p); p += 48
        rl, cl, k, m, mm, indel, ea, nin, nout, woff, wlen, strand = (int(x) for x in h)
        read = data[p:p + rl]; p += rl
        twin = data[p:p + wlen]; p += wlen
        bi = np.frombuffer(data, dtype=np.uint32, count=3 * nin, offset=p).reshape(-1, 3).copy(); p += 12 * nin
        bo = np.frombuffer(data, dtype=np.uint32, count=3 * nout, offset=p).reshape(-1, 3).copy(); p += 12 * nout
        out.append(dict(read=read, twin=twin, t_win_off=woff, contig_len=cl, k=k, match=m, mismatch=mm, indel=indel,
                        end_align=ea, blocks_in=bi, blocks_out=bo, strand=strand))
    return out


def indel_refine_groups_port(read, twin, t_win_off, contig_len, blocks_in, k, match, mismatch, indel, end_align, max_groups=4096):
    """Like indel_refine_port, also returning the banded DP groups: list of dict(qStart,tStart,tLen,qSeqLen,tSeqLen,qS,qE,blocks)."""
    L = port(); _bind_ir(L)
    if not getattr(L, "_irg_bound", False):
        L.lra_oracle_indel_refine_groups.restype = C.c_int
        L.lra_oracle_indel_refine_groups.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_long, C.c_long, _u32p, C.c_int, C.c_int,
                                                     C.c_int, C.c_int, C.c_int, C.c_int, _u32p, C.c_int, C.POINTER(C.c_int),
                                                     C.POINTER(C.c_int), _i32p, C.c_int, _i32p, C.c_long]
        L._irg_bound = True
    bi = np.ascontiguousarray(blocks_in, dtype=np.uint32).reshape(-1)
    cap = len(read) + 8
    out = np.zeros(cap * 3, np.uint32)
    meta = np.zeros(max_groups * 8, np.int32)
    band_cap = 4 * (len(read) + len(twin)) + 1024
    band = np.zeros(band_cap, np.int32)
    n, st = C.c_int(0), C.c_int(0)
    ng = L.lra_oracle_indel_refine_groups(bytes(read), len(read), bytes(twin), t_win_off, contig_len, bi, len(bi) // 3, k, match,
                                          mismatch, indel, 1 if end_align else 0, out, cap, C.byref(n), C.byref(st), meta, max_groups,
                                          band, band_cap)
    blocks = out[: 3 * n.value].reshape(-1, 3).copy()
    groups = []
    for g in range(ng):
        qs, ts, tl, qsl, tsl, boff, fo, no = (int(x) for x in meta[8 * g: 8 * g + 8])
        groups.append(dict(qStart=qs, tStart=ts, tLen=tl, qSeqLen=qsl, tSeqLen=tsl, qS=band[boff:boff + tl].copy(),
                           qE=band[boff + tl:boff + 2 * tl].copy(), blocks=blocks[fo:fo + no].copy()))
    return blocks, st.value, groups


def indel_refine_batch_ref(sb, t_arena, t_base, nthreads=1, want_blocks=True):
    """Whole-function batch through the UNMODIFIED reference (libref_lra.so).  sb as for lra_b200_indel_refine_batch but with the
    target arena / bases given explicitly (CPU arms use compact window arenas).  Returns (n_blocks, out_off, blocks)."""
    L = ref()
    if not getattr(L, "_irb_bound", False):
        L.ref_indel_refine_batch.restype = C.c_int
        L.ref_indel_refine_batch.argtypes = [_u8p, _u8p, _u32p, _u64p, _i32p, _u32p, _u32p, _i32p, _i32p, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_int, _i32p, _u64p, C.c_void_p, C.c_int]
        L._irb_bound = True
    S = len(sb["blk_cnt"])
    cap = (sb["read_len"].astype(np.int64) + sb["contig_len"]).astype(np.uint64)
    off = np.zeros(S, np.uint64); off[1:] = np.cumsum(cap[:-1])
    n = np.zeros(S, np.int32)
    blocks = np.zeros((int(cap.sum()) + 1) * 3, np.uint32) if want_blocks else None
    L.ref_indel_refine_batch(sb["q_arena"], t_arena, np.ascontiguousarray(sb["blocks_in"], np.uint32).reshape(-1),
                             np.ascontiguousarray(sb["blk_off"], np.uint64), sb["blk_cnt"], sb["q_base"], t_base, sb["read_len"],
                             sb["contig_len"], S, sb["k"], sb["match"], sb["mismatch"], sb["indel"], sb["end_align"], n, off,
                             blocks.ctypes.data if want_blocks else None, nthreads)
    return n, off, (blocks.reshape(-1, 3) if want_blocks else None)


def indel_refine_batch_port(sb, t_arena, t_base):
    """Same through the C restatement (single thread)."""
    S = len(sb["blk_cnt"])
    outs = []
    for s in range(S):
        o, c = int(sb["blk_off"][s]), int(sb["blk_cnt"][s])
        qb, tb = int(sb["q_base"][s]), int(t_base[s])
        bo, st, cells = indel_refine_port(sb["q_arena"][qb:qb + sb["read_len"][s]].tobytes(), t_arena[tb:tb + sb["contig_len"][s]].tobytes(), 0,
                                          int(sb["contig_len"][s]), sb["blocks_in"][o:o + c], sb["k"], sb["match"], sb["mismatch"], sb["indel"],
                                          sb["end_align"])
        assert st == 0
        outs.append(bo)
    return outs


# ---------------------------------------------------------------- a1-a5 seeding prefix of MapRead

def _bind_seed(L, prefix):
    if getattr(L, "_seed_bound", False):
        return
    f = getattr(L, prefix + "store_minimizers"); f.restype = C.c_long
    f.argtypes = [_u8p, C.c_uint32, C.c_int, C.c_int, _u64p, _u32p, C.c_long]
    f = getattr(L, prefix + "sort_minimizers"); f.restype = None
    f.argtypes = [_u64p, _u32p, C.c_long]
    f = getattr(L, prefix + "compare_lists"); f.restype = C.c_long
    f.argtypes = [_u64p, _u32p, C.c_long, _u64p, _u32p, C.c_long, C.c_long if prefix.startswith("lra_oracle") else C.c_int, _u64p, _u32p, _u64p, _u32p, C.c_long]
    f = getattr(L, prefix + "seed_read"); f.restype = C.c_long
    f.argtypes = [_u8p, C.c_uint32, _u8p, _u64p, _u32p, C.c_long, C.c_int, C.c_int, C.c_long if prefix.startswith("lra_oracle") else C.c_int,
                  _u64p, _u32p, _u64p, _u32p, _u8p, C.c_long]
    L._seed_bound = True


def _seed_lib(which):
    if which == "ref":
        L = ref(); _bind_seed(L, "ref_"); return L, "ref_"
    L = port(); _bind_seed(L, "lra_oracle_"); return L, "lra_oracle_"


def _u8(a):
    return np.frombuffer(a, dtype=np.uint8).copy() if isinstance(a, (bytes, bytearray)) else np.ascontiguousarray(a, np.uint8)


def store_minimizers(seq, k, w, which="port"):
    L, p = _seed_lib(which)
    s = _u8(seq); n = len(s)
    s = np.concatenate([s, np.zeros(8, np.uint8)])
    t = np.zeros(n + 16, np.uint64); pos = np.zeros(n + 16, np.uint32)
    m = getattr(L, p + "store_minimizers")(s, n, k, w, t, pos, n + 16)
    return t[:m].copy(), pos[:m].copy()


def sort_minimizers(t, pos, which="port"):
    L, p = _seed_lib(which)
    t = np.ascontiguousarray(t, np.uint64).copy(); pos = np.ascontiguousarray(pos, np.uint32).copy()
    getattr(L, p + "sort_minimizers")(t, pos, len(t))
    return t, pos


def compare_lists(qt, qpos, tt, tpos, max_freq, which="port"):
    L, p = _seed_lib(which)
    cap = 4 * (len(qt) + 16) * 8 + 1024
    for _ in range(3):
        r = [np.zeros(cap, np.uint64), np.zeros(cap, np.uint32), np.zeros(cap, np.uint64), np.zeros(cap, np.uint32)]
        n = getattr(L, p + "compare_lists")(np.ascontiguousarray(qt, np.uint64), np.ascontiguousarray(qpos, np.uint32), len(qt),
                                            np.ascontiguousarray(tt, np.uint64), np.ascontiguousarray(tpos, np.uint32), len(tt), max_freq,
                                            r[0], r[1], r[2], r[3], cap)
        if n <= cap:
            return [x[:n].copy() for x in r]
        cap = n + 16
    raise RuntimeError("compare_lists capacity")


def seed_read(read, genome_concat, tt, tpos, k, w, max_freq, which="port"):
    """a2..a5 for one read.  Returns (q_t, q_pos, t_t, t_pos, strand) in the reference's allMatches order."""
    L, p = _seed_lib(which)
    rd = _u8(read); n = len(rd)
    rd = np.concatenate([rd, np.zeros(8, np.uint8)])
    cap = 8 * n + 1024
    for _ in range(3):
        r = [np.zeros(cap, np.uint64), np.zeros(cap, np.uint32), np.zeros(cap, np.uint64), np.zeros(cap, np.uint32), np.zeros(cap, np.uint8)]
        m = getattr(L, p + "seed_read")(rd, n, _u8(genome_concat) if not isinstance(genome_concat, np.ndarray) else genome_concat,
                                        np.ascontiguousarray(tt, np.uint64), np.ascontiguousarray(tpos, np.uint32), len(tt), k, w, max_freq,
                                        r[0], r[1], r[2], r[3], r[4], cap)
        if m <= cap:
            return [x[:m].copy() for x in r]
        cap = m + 16
    raise RuntimeError("seed_read capacity")


def read_mms(path):
    """Parse a reference global index file (<ref>.mms, MMIndex.h:416-424).
    Returns dict(k, names, pos (cumulative contig offsets), t (uint64), tpos (uint32))."""
    data = open(path, "rb").read()
    n = int(np.frombuffer(data, np.int64, 1, 0)[0]); k = int(np.frombuffer(data, np.int32, 1, 8)[0])
    p = 12
    nc = int(np.frombuffer(data, np.int32, 1, p)[0]); p += 4
    names = []
    for _ in range(nc):
        ln = int(np.frombuffer(data, np.int32, 1, p)[0]); p += 4
        names.append(data[p:p + ln].decode()); p += ln
    pos = np.frombuffer(data, np.uint64, nc + 1, p).copy(); p += 8 * (nc + 1)
    rec = np.frombuffer(data, np.dtype([("t", "<u8"), ("pos", "<u4"), ("pad", "<u4")]), n, p)
    return dict(k=k, names=names, pos=pos, t=rec["t"].copy(), tpos=rec["pos"].copy())


# ---------------------------------------------------------------- a21 CalculateStatistics

def log_lut():
    """The reference's LookUpTable: logf(i) for i = 1, 6, ..., 10001 (LogLookUpTable.h:9-15), built with the HOST libm."""
    L = port()
    if getattr(L, "_lut", None) is None:
        libm = C.CDLL("libm.so.6"); libm.logf.restype = C.c_float; libm.logf.argtypes = [C.c_float]
        L._lut = np.array([libm.logf(float(i)) for i in range(1, 10002, 5)], np.float32)
    return L._lut


_CIG = {0: "M", 1: "I", 2: "D", 7: "=", 8: "X"}


def cigar_string(ops):
    return "".join("%d%s" % (int(o) >> 4, _CIG[int(o) & 15]) for o in ops)


def calc_stats_port(read, text, t_win_off, blocks):
    L = port()
    if not getattr(L, "_st_bound", False):
        L.lra_oracle_calc_stats.restype = C.c_long
        L.lra_oracle_calc_stats.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_long, _u32p, C.c_int, np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS"),
                                            _u32p, C.c_long, _i32p, C.POINTER(C.c_float)]
        L._st_bound = True
    b = np.ascontiguousarray(blocks, np.uint32).reshape(-1)
    cap = 2 * (len(read) + len(text)) + 16
    cig = np.zeros(cap, np.uint32); st = np.zeros(16, np.int32); v = C.c_float(0)
    n = L.lra_oracle_calc_stats(bytes(read), len(read), bytes(text), t_win_off, b, len(b) // 3, log_lut(), cig, cap, st, C.byref(v))
    return st, np.float32(v.value), cig[:n].copy()


def calc_stats_ref(read, text, blocks):
    """Through the unmodified reference; blocks' tPos index `text` directly.  Returns (stats, value, cigar string)."""
    L = ref()
    if not getattr(L, "_st_bound", False):
        L.ref_calc_stats.restype = C.c_int
        L.ref_calc_stats.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, _u32p, C.c_int, _i32p, C.POINTER(C.c_float), C.c_char_p, C.c_int]
        L._st_bound = True
    b = np.ascontiguousarray(blocks, np.uint32).reshape(-1)
    st = np.zeros(16, np.int32); v = C.c_float(0)
    buf = C.create_string_buffer(12 * (len(read) + len(text)) + 64)
    L.ref_calc_stats(bytes(read), len(read), bytes(text), len(text), b, len(b) // 3, st, C.byref(v), buf, len(buf))
    return st, np.float32(v.value), buf.value.decode()


# ---------------------------------------------------------------- a12 LocalIndex::IndexSeq, a13 REFINEclusters

import threading
_bind_lock = threading.Lock()


def _bind_once(L, name, restype, argtypes):
    """ctypes prototypes are set once per library object: bench.py calls these wrappers from a thread pool, and re-assigning argtypes
    while another thread is inside a call is a race."""
    f = getattr(L, name)
    if not getattr(f, "_lra_bound", False):
        with _bind_lock:
            if not getattr(f, "_lra_bound", False):
                f.restype = restype; f.argtypes = argtypes; f._lra_bound = True
    return f


class LocalIndexData:
    """A LocalIndex as three arrays: seq_off (uint64, leading 0), bnd (uint64, leading 0), mins (uint32: tuple | pos << 20)."""
    def __init__(self, seq_off, bnd, mins, k=10, w=5, window=2048, max_freq=15):
        self.seq_off, self.bnd, self.mins = seq_off, bnd, mins
        self.k, self.w, self.window, self.max_freq = k, w, window, max_freq


def local_index(seqs, k=10, w=5, window=2048, max_freq=15, which="port"):
    """LocalIndex of one sequence (a read strand) or of a list of contigs (the genome: IndexFile calls IndexSeq per contig)."""
    if isinstance(seqs, (bytes, bytearray, np.ndarray)):
        seqs = [seqs]
    seqs = [np.frombuffer(s, np.uint8) if isinstance(s, (bytes, bytearray)) else np.ascontiguousarray(s, np.uint8) for s in seqs]
    if which == "ref":
        L = ref()
        _bind_once(L, "ref_lidx_new", C.c_void_p, [C.c_int] * 4)
        _bind_once(L, "ref_lidx_index_seq", None, [C.c_void_p, _u8p, C.c_int])
        _bind_once(L, "ref_lidx_sizes", None, [C.c_void_p, C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_long)])
        _bind_once(L, "ref_lidx_copy", None, [C.c_void_p, _u64p, _u64p, _u32p])
        _bind_once(L, "ref_lidx_free", None, [C.c_void_p])
        h = C.c_void_p(L.ref_lidx_new(k, w, window, max_freq))
        for s in seqs:
            L.ref_lidx_index_seq(h, np.concatenate([s, np.zeros(8, np.uint8)]), len(s))
        a, b, c = C.c_long(), C.c_long(), C.c_long()
        L.ref_lidx_sizes(h, C.byref(a), C.byref(b), C.byref(c))
        off = np.zeros(a.value, np.uint64); bnd = np.zeros(b.value, np.uint64); mins = np.zeros(max(c.value, 1), np.uint32)
        L.ref_lidx_copy(h, off, bnd, mins)
        L.ref_lidx_free(h)
        return LocalIndexData(off, bnd, mins[:c.value], k, w, window, max_freq)
    L = port()
    _bind_once(L, "lra_oracle_index_seq", C.c_long, [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u64p, _u64p, _u32p, C.c_long, C.POINTER(C.c_int)])
    offs, bnds, minss = [np.zeros(1, np.uint64)], [np.zeros(1, np.uint64)], []
    base, nmin = 0, 0
    for s in seqs:
        nwin = (len(s) + window - 1) // window
        off = np.zeros(nwin + 1, np.uint64); bnd = np.zeros(nwin + 1, np.uint64); mins = np.zeros(len(s) + 16, np.uint32)
        ni = C.c_int()
        n = L.lra_oracle_index_seq(np.concatenate([s, np.zeros(8, np.uint8)]), len(s), k, w, window, max_freq, off, bnd, mins, len(mins), C.byref(ni))
        assert n >= 0 and ni.value == nwin
        offs.append(off[1:] + np.uint64(base)); bnds.append(bnd[1:] + np.uint64(nmin)); minss.append(mins[:n])
        base += len(s); nmin += n
    return LocalIndexData(np.concatenate(offs), np.concatenate(bnds), np.concatenate(minss) if minss else np.zeros(0, np.uint32), k, w, window, max_freq)


def local_minimizers_port(seq, k=10, w=5):
    L = port()
    L.lra_oracle_local_minimizers.restype = C.c_long
    L.lra_oracle_local_minimizers.argtypes = [_u8p, C.c_uint32, C.c_int, C.c_int, _u32p, C.c_long]
    s = np.frombuffer(seq, np.uint8) if isinstance(seq, (bytes, bytearray)) else seq
    out = np.zeros(len(s) + 8, np.uint32)
    n = L.lra_oracle_local_minimizers(np.concatenate([s, np.zeros(8, np.uint8)]), len(s), k, w, out, len(out))
    return out[:n]


def compare_lists_local_port(q, t, max_freq):
    L = port()
    L.lra_oracle_compare_lists_local.restype = C.c_long
    L.lra_oracle_compare_lists_local.argtypes = [_u32p, C.c_long, _u32p, C.c_long, C.c_long, _u32p, _u32p, C.c_long]
    q = np.ascontiguousarray(q, np.uint32); t = np.ascontiguousarray(t, np.uint32)
    cap = 4 * (len(q) + len(t)) + 64
    while True:
        rq = np.zeros(cap, np.uint32); rt = np.zeros(cap, np.uint32)
        n = L.lra_oracle_compare_lists_local(q if len(q) else np.zeros(1, np.uint32), len(q), t if len(t) else np.zeros(1, np.uint32), len(t), max_freq, rq, rt, cap)
        if n <= cap:
            return rq[:n], rt[:n]
        cap = n


def refine_cluster(mq, mt, box, strand, read_len, hdr_pos, gl, rd_fwd, rd_rev, global_k, small_k=10, window=100, local_max_freq=15, which="port",
                   ref_handles=None):
    """One cluster through REFINEclusters.  Returns a dict: status, chrom, mq/mt/box (as the reference leaves them), diag (min,max),
    rq/rt/rtup (refined anchors), rbox (qStart,qEnd,tStart,tEnd of the refined cluster), eff (float32)."""
    mq = np.array(mq, np.uint32); mt = np.array(mt, np.uint32); box = np.array(box, np.uint32)
    hdr = np.ascontiguousarray(hdr_pos, np.uint64)
    info = np.zeros(8, np.int32); diag = np.zeros(2, np.int64); eff = C.c_float()
    cap = 1 << 16
    mq1 = mq if len(mq) else np.zeros(1, np.uint32); mt1 = mt if len(mt) else np.zeros(1, np.uint32)
    while True:
        a, b, bx = mq1.copy(), mt1.copy(), box.copy()
        rq = np.zeros(cap, np.uint32); rt = np.zeros(cap, np.uint32); ru = np.zeros(cap, np.uint32)
        if which == "ref":
            L = ref()
            _bind_once(L, "ref_refine_cluster", C.c_long, [_u32p, _u32p, C.c_long, _u32p, C.c_int, C.c_uint32, _u64p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                            C.c_int, C.c_int, C.c_int, C.c_long, _u32p, _u32p, _u32p, C.c_long, _i32p, _i64p, C.POINTER(C.c_float)])
            n = L.ref_refine_cluster(a, b, len(mq), bx, strand, read_len, hdr, len(hdr), ref_handles[0], ref_handles[1], ref_handles[2],
                                     global_k, small_k, window, local_max_freq, rq, rt, ru, cap, info, diag, C.byref(eff))
        else:
            L = port()
            _bind_once(L, "lra_oracle_refine_cluster", C.c_long, [_u32p, _u32p, C.c_long, _u32p, C.c_int, C.c_uint32, _u64p, C.c_int,
                                                                   _u64p, C.c_long, _u64p, _u32p, _u64p, C.c_long, _u64p, _u32p,
                                                                   C.c_int, C.c_int, C.c_int, C.c_long, _u32p, _u32p, _u32p, C.c_long, _i32p, _i64p, C.POINTER(C.c_float)])
            rd = rd_rev if strand else rd_fwd
            pad = lambda m: m if len(m) else np.zeros(1, np.uint32)
            n = L.lra_oracle_refine_cluster(a, b, len(mq), bx, strand, read_len, hdr, len(hdr), gl.seq_off, len(gl.seq_off), gl.bnd, pad(gl.mins),
                                            rd.seq_off, len(rd.seq_off), rd.bnd, pad(rd.mins), global_k, small_k, window, local_max_freq,
                                            rq, rt, ru, cap, info, diag, C.byref(eff))
        if n <= cap:
            break
        cap = n
    return dict(status=int(info[0]), chrom=int(info[1]), mq=a[:len(mq)], mt=b[:len(mq)], box=bx, diag=diag.copy(), rq=rq[:n], rt=rt[:n], rtup=ru[:n],
                rbox=info[2:6].astype(np.uint32), eff=np.float32(eff.value))


class RefLocalIndexHandle:
    """A live reference LocalIndex (for ref_refine_cluster)."""
    def __init__(self, seqs, k=10, w=5, window=2048, max_freq=15):
        L = ref()
        _bind_once(L, "ref_lidx_new", C.c_void_p, [C.c_int] * 4)
        _bind_once(L, "ref_lidx_index_seq", None, [C.c_void_p, _u8p, C.c_int])
        _bind_once(L, "ref_lidx_free", None, [C.c_void_p])
        self.L = L
        self.h = C.c_void_p(L.ref_lidx_new(k, w, window, max_freq))
        if isinstance(seqs, (bytes, bytearray, np.ndarray)):
            seqs = [seqs]
        for s in seqs:
            s = np.frombuffer(s, np.uint8) if isinstance(s, (bytes, bytearray)) else np.ascontiguousarray(s, np.uint8)
            L.ref_lidx_index_seq(self.h, np.concatenate([s, np.zeros(8, np.uint8)]), len(s))

    def close(self):
        if self.h:
            self.L.ref_lidx_free(self.h); self.h = None


def refine_splitchain(mq, mt, mlen, mstrand, box, chrom, strand, read_len, hdr_pos, gl, rd_fwd, rd_rev, global_k, small_k=10, window=100, local_max_freq=15,
                      limitrefine=1, which="port", ref_handles=None):
    """One split chain through Refine_splitchain (ChainRefine.h:383-576).  mstrand: strand of the cluster each anchor comes from.
    Returns dict(status, chrom, diag, rq, rt, rtup, rbox, eff)."""
    mq = np.ascontiguousarray(mq, np.uint32); mt = np.ascontiguousarray(mt, np.uint32); mlen = np.ascontiguousarray(mlen, np.uint32)
    mstrand = np.ascontiguousarray(mstrand, np.uint8); box = np.ascontiguousarray(box, np.uint32)
    hdr = np.ascontiguousarray(hdr_pos, np.uint64)
    info = np.zeros(8, np.int32); diag = np.zeros(2, np.int64); eff = C.c_float()
    n = len(mq)
    pad = lambda a, dt: a if len(a) else np.zeros(1, dt)
    cap = 1 << 16
    while True:
        rq = np.zeros(cap, np.uint32); rt = np.zeros(cap, np.uint32); ru = np.zeros(cap, np.uint32)
        if which == "ref":
            L = ref()
            _bind_once(L, "ref_refine_splitchain", C.c_long,
                       [_u32p, _u32p, _u32p, _i32p, C.c_long, _u8p, C.c_int, _u32p, C.c_int, C.c_int, C.c_uint32, _u64p, C.c_int,
                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_long, C.c_int, _u32p, _u32p, _u32p, C.c_long,
                        _i32p, _i64p, C.POINTER(C.c_float)])
            # anchors of strand s live in cluster s: two clusters, each flipped once by the reference
            cluster_of = mstrand.astype(np.int32)
            m = L.ref_refine_splitchain(pad(mq, np.uint32), pad(mt, np.uint32), pad(mlen, np.uint32), pad(cluster_of, np.int32), n, np.array([0, 1], np.uint8), 2,
                                        box, chrom, strand, read_len, hdr, len(hdr), ref_handles[0], ref_handles[1], ref_handles[2], global_k, small_k, window,
                                        local_max_freq, limitrefine, rq, rt, ru, cap, info, diag, C.byref(eff))
            assert m >= 0, "the reference did not restore the clusters"
        else:
            L = port()
            _bind_once(L, "lra_oracle_refine_splitchain", C.c_long,
                       [_u32p, _u32p, _u32p, _u8p, C.c_long, _u32p, C.c_int, C.c_int, C.c_uint32, _u64p, C.c_int,
                        _u64p, C.c_long, _u64p, _u32p, _u64p, C.c_long, _u64p, _u32p, C.c_int, C.c_int, C.c_int, C.c_long, C.c_int,
                        _u32p, _u32p, _u32p, C.c_long, _i32p, _i64p, C.POINTER(C.c_float)])
            rd = rd_rev if strand else rd_fwd
            m = L.lra_oracle_refine_splitchain(pad(mq, np.uint32), pad(mt, np.uint32), pad(mlen, np.uint32), pad(mstrand, np.uint8), n, box, chrom, strand, read_len,
                                               hdr, len(hdr), gl.seq_off, len(gl.seq_off), gl.bnd, pad(gl.mins, np.uint32), rd.seq_off, len(rd.seq_off), rd.bnd,
                                               pad(rd.mins, np.uint32), global_k, small_k, window, local_max_freq, limitrefine, rq, rt, ru, cap, info, diag, C.byref(eff))
        if m <= cap:
            break
        cap = m
    return dict(status=int(info[0]), chrom=int(info[1]), diag=diag.copy(), rq=rq[:m], rt=rt[:m], rtup=ru[:m], rbox=info[2:6].astype(np.uint32), eff=np.float32(eff.value))


# ---------------------------------------------------------------- a6 anchor sorts (Sorting.h)

def sort_matches(mode, q, t, which="port"):
    """mode 0 DiagonalSort, 1 AntiDiagonalSort, 2 CartesianSort, 3 CartesianTargetSort on one anchor list.  Returns (q, t[, perm])."""
    q = np.array(q, np.uint32); t = np.array(t, np.uint32)
    n = len(q)
    if n == 0:
        return q, t, np.zeros(0, np.uint32)
    if which == "ref":
        L = ref()
        L.ref_sort_matches.argtypes = [C.c_int, _u32p, _u32p, C.c_long]; L.ref_sort_matches.restype = None
        L.ref_sort_matches(mode, q, t, n)
        return q, t, None
    L = port()
    L.lra_oracle_sort_matches.argtypes = [C.c_int, _u32p, _u32p, C.c_long, _u32p]; L.lra_oracle_sort_matches.restype = None
    perm = np.zeros(n, np.uint32)
    L.lra_oracle_sort_matches(mode, q, t, n, perm)
    return q, t, perm


# ---------------------------------------------------------------- a24 GlobalChain / PrioritySearchTree

def global_chain(frag, score, which="port"):
    """frag[n,4] = xl, yl, xh, yh (int32), score[n].  Returns (chain indices, final scores, prev)."""
    frag = np.ascontiguousarray(frag, np.int32).reshape(-1, 4); n = len(frag)
    sc = np.array(score, np.int32); prev = np.full(max(n, 1), -1, np.int32); chain = np.zeros(max(n, 1), np.int32)
    L = ref() if which == "ref" else port()
    f = L.ref_global_chain if which == "ref" else L.lra_oracle_global_chain
    f.restype = C.c_long; f.argtypes = [_i32p, _i32p, _i32p, C.c_long, _i32p]
    m = f(frag.reshape(-1) if n else np.zeros(4, np.int32), sc if n else np.zeros(1, np.int32), prev, n, chain)
    return chain[:m].copy(), sc, prev[:n]


# ---------------------------------------------------------------- a20 RefineBreakpoint

def refine_breakpoint_ref(lread, rread, read_len, lchrom, rchrom, lblocks, lstrand, rblocks, rstrand):
    """Through the unmodified reference.  Returns the two updated block lists."""
    L = ref()
    _bind_once(L, "ref_refine_breakpoint", C.c_int, [_u8p, _u8p, C.c_int, _u8p, C.c_int, _u8p, C.c_int, _u32p, C.POINTER(C.c_int), C.c_int, _u32p,
                                                     C.POINTER(C.c_int), C.c_int, C.c_int])
    cap = len(lblocks) + len(rblocks) + 1200
    lb = np.zeros((cap, 3), np.uint32); rb = np.zeros((cap, 3), np.uint32)
    lb[:len(lblocks)] = lblocks; rb[:len(rblocks)] = rblocks
    ln, rn = C.c_int(len(lblocks)), C.c_int(len(rblocks))
    pad = lambda a: np.concatenate([np.ascontiguousarray(a, np.uint8), np.zeros(8, np.uint8)])
    rc = L.ref_refine_breakpoint(pad(lread), pad(rread), read_len, pad(lchrom), len(lchrom), pad(rchrom), len(rchrom), lb.reshape(-1), C.byref(ln), lstrand,
                                 rb.reshape(-1), C.byref(rn), rstrand, cap)
    assert rc == 0
    return lb[:ln.value].copy(), rb[:rn.value].copy()


def splice_blocks(blocks, mode, bound, new):
    """Apply one side of a refine-breakpoint result (mode 1 append / 2 prepend, boundary block after the merge, blocks to splice in)."""
    b = np.array(blocks, np.uint32).reshape(-1, 3)
    if mode == 1:
        b[-1] = bound
        return np.concatenate([b, np.asarray(new, np.uint32).reshape(-1, 3)])
    if mode == 2:
        b[0] = bound
        return np.concatenate([np.asarray(new, np.uint32).reshape(-1, 3), b])
    return b


def refine_breakpoint_port(lread, rread, read_len, lchrom, rchrom, lblocks, lstrand, rblocks, rstrand):
    """Through the C restatement.  Returns (refined?, left blocks, right blocks)."""
    L = port()
    _bind_once(L, "lra_oracle_refine_breakpoint", C.c_int, [_u8p, _u8p, C.c_int, _u8p, C.c_int, _u8p, C.c_int, _u32p, _u32p, C.c_int, _u32p, _u32p, C.c_int,
                                                            _i32p, _i32p, _u32p, _u32p, C.c_int])
    cap = 600
    mode = np.zeros(2, np.int32); n_out = np.zeros(2, np.int32); bound = np.zeros(6, np.uint32); out = np.zeros(2 * cap * 3, np.uint32)
    pad = lambda a: np.concatenate([np.ascontiguousarray(a, np.uint8), np.zeros(8, np.uint8)])
    lb = np.ascontiguousarray(lblocks, np.uint32).reshape(-1, 3); rb = np.ascontiguousarray(rblocks, np.uint32).reshape(-1, 3)
    r = L.lra_oracle_refine_breakpoint(pad(lread), pad(rread), read_len, pad(lchrom), len(lchrom), pad(rchrom), len(rchrom), lb[0].copy(), lb[-1].copy(), lstrand,
                                       rb[0].copy(), rb[-1].copy(), rstrand, mode, n_out, bound, out, cap)
    o = out.reshape(2, cap, 3)
    return r, splice_blocks(lb, mode[0], bound[:3], o[0, :n_out[0]]), splice_blocks(rb, mode[1], bound[3:], o[1, :n_out[1]]), (mode.copy(), n_out.copy(), bound.copy(), o)


# ---------------------------------------------------------------- a16 chain filters (Chain.h:546-960)

def chain_filter(mode, q, t, length, strand, which="port"):
    """keep mask of one chain (anchors in chain order).  mode 0 RemoveSmallPairedIndels, 1 RemovePairedIndels(refineEnds), 2 RemovePairedIndels(no
    refineEnds), 3 RemovePairedIndels(matches, chain, lengths), 4 RemoveSpuriousAnchors, 5 RemoveSpuriousJump."""
    q = np.ascontiguousarray(q, np.uint32); t = np.ascontiguousarray(t, np.uint32); length = np.ascontiguousarray(length, np.uint32)
    strand = np.ascontiguousarray(strand, np.uint8)
    n = len(q)
    keep = np.zeros(max(n, 1), np.uint8)
    L = ref() if which == "ref" else port()
    f = _bind_once(L, "ref_chain_filter" if which == "ref" else "lra_oracle_chain_filter", None, [C.c_int, _u32p, _u32p, _u32p, _u8p, C.c_long, _u8p])
    pad = lambda a, dt: a if n else np.zeros(1, dt)
    f(mode, pad(q, np.uint32), pad(t, np.uint32), pad(length, np.uint32), pad(strand, np.uint8), n, keep)
    return keep[:n]


# ---------------------------------------------------------------- a7 CleanOffDiagonal (Clustering.h:565-868)

COD_FIELDS = ["cleanMaxDiag", "minDiagCluster", "bypassClustering", "cleanClustersize", "SecondCleanMinDiagCluster", "punish_anchorfreq", "anchorPerlength",
              "SecondCleanMaxDiag", "ExtractDiagonalFromClean", "globalK"]
COD_PRESETS = {      # lra.cpp:268-431 (align presets) / Options.h:123-240
    "ccs": dict(cleanMaxDiag=150, minDiagCluster=30, bypassClustering=0, cleanClustersize=100, SecondCleanMinDiagCluster=30, punish_anchorfreq=10, anchorPerlength=10,
                SecondCleanMaxDiag=100, ExtractDiagonalFromClean=1, globalK=17),
    "clr": dict(cleanMaxDiag=150, minDiagCluster=10, bypassClustering=1, cleanClustersize=100, SecondCleanMinDiagCluster=30, punish_anchorfreq=10, anchorPerlength=10,
                SecondCleanMaxDiag=100, ExtractDiagonalFromClean=1, globalK=15),
    "ont": dict(cleanMaxDiag=200, minDiagCluster=3, bypassClustering=1, cleanClustersize=100, SecondCleanMinDiagCluster=10, punish_anchorfreq=5, anchorPerlength=5,
                SecondCleanMaxDiag=120, ExtractDiagonalFromClean=1, globalK=17),
    "noextract": dict(cleanMaxDiag=100, minDiagCluster=10, bypassClustering=0, cleanClustersize=100, SecondCleanMinDiagCluster=40, punish_anchorfreq=10, anchorPerlength=10,
                      SecondCleanMaxDiag=10, ExtractDiagonalFromClean=0, globalK=17),
}


def clean_off_diagonal(q, t, qt, strand, opts, hdr_pos, which="port"):
    """One anchor list (sorted by DiagonalSort / AntiDiagonalSort) through CleanOffDiagonal.  Returns dict(kq, kt, kfreq (the surviving anchors),
    cl[ncl,7] (start, end, qStart, qEnd, tStart, tEnd, chromIndex), cl_freq) and, for the port, keep / freq / cnt per input anchor."""
    q = np.array(q, np.uint32); t = np.array(t, np.uint32); qt = np.ascontiguousarray(qt, np.uint64)
    n = len(q)
    hdr = np.ascontiguousarray(hdr_pos, np.uint64)
    ov = np.array([opts[k] for k in COD_FIELDS], np.int32)
    cl = np.zeros(7 * (n + 1), np.int32); clf = np.zeros(n + 1, np.float32)
    pad = lambda a, dt: a if n else np.zeros(1, dt)
    if which == "ref":
        L = ref()
        f = _bind_once(L, "ref_clean_off_diagonal", C.c_long, [_u32p, _u32p, _u64p, C.c_long, C.c_int, _i32p, _u64p, C.c_int, np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS"),
                                                               C.POINTER(C.c_long), _i32p, np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")])
        freq = np.zeros(n + 1, np.float32); nk = C.c_long(0)
        qq, tt = pad(q, np.uint32), pad(t, np.uint32)
        ncl = f(qq, tt, pad(qt, np.uint64), n, strand, ov, hdr, len(hdr), freq, C.byref(nk), cl, clf)
        k = nk.value
        return dict(kq=qq[:k].copy(), kt=tt[:k].copy(), kfreq=freq[:k].copy(), cl=cl[:7 * ncl].reshape(-1, 7).copy(), cl_freq=clf[:ncl].copy())
    L = port()
    f = _bind_once(L, "lra_oracle_clean_off_diagonal", C.c_long, [_u32p, _u32p, _u64p, C.c_long, C.c_int, _i32p, _u64p, C.c_int, _u8p,
                                                                  np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS"), _i32p, _i32p,
                                                                  np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")])
    keep = np.zeros(n + 1, np.uint8); freq = np.zeros(n + 1, np.float32); cnt = np.zeros(n + 1, np.int32)
    ncl = f(pad(q, np.uint32), pad(t, np.uint32), pad(qt, np.uint64), n, strand, ov, hdr, len(hdr), keep, freq, cnt, cl, clf)
    m = keep[:n] == 1
    return dict(kq=q[m], kt=t[m], kfreq=freq[:n][m], cl=cl[:7 * ncl].reshape(-1, 7).copy(), cl_freq=clf[:ncl].copy(), keep=keep[:n].copy(), freq=freq[:n].copy(), cnt=cnt[:n].copy())


# ---------------------------------------------------------------- a9 SplitClusters + DecideSplitClustersValue (SplitClusters.h)

def split_clusters(box, strand, freq, contig, mq, m_off, global_k, which="port"):
    """The clusters of one read (box[n,4] = qStart,qEnd,tStart,tEnd; anchors' read positions mq per cluster, CartesianSort order).
    Returns dict(split, val_cluster, sp[k,6] (qStart,qEnd,tStart,tEnd,strand,coarse), sp_val, sp_n0)."""
    box = np.ascontiguousarray(box, np.uint32).reshape(-1); strand = np.ascontiguousarray(strand, np.uint8); freq = np.ascontiguousarray(freq, np.float32)
    mq = np.ascontiguousarray(mq, np.uint32); m_off = np.ascontiguousarray(m_off, np.uint64)
    n = len(strand)
    L = ref() if which == "ref" else port()
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    f = _bind_once(L, "ref_split_clusters" if which == "ref" else "lra_oracle_split_clusters", C.c_long,
                   [_u32p, _u8p, f32p, C.c_long, C.c_int, _u32p, _u64p, C.c_int, _u8p, _i32p, _u32p, _i32p, _i32p, C.c_long])
    cap = 64 * (n + 1)
    pad = lambda a, dt: a if len(a) else np.zeros(1, dt)
    while True:
        split = np.zeros(max(n, 1), np.uint8); vc = np.zeros(max(n, 1), np.int32); sp = np.zeros(6 * cap, np.uint32); sv = np.zeros(cap, np.int32); s0 = np.zeros(cap, np.int32)
        ns = f(pad(box, np.uint32), pad(strand, np.uint8), pad(freq, np.float32), n, contig, pad(mq, np.uint32), m_off, global_k, split, vc, sp, sv, s0, cap)
        if ns <= cap:
            return dict(split=split[:n], val_cluster=vc[:n], sp=sp[:6 * ns].reshape(-1, 6).copy(), sp_val=sv[:ns].copy(), sp_n0=s0[:ns].copy())
        cap = ns


# ---------------------------------------------------------------- a22 SetFromSegAlignment / AlignmentsOrder / SimpleMapQV

def mapq(rd, bypass, read_type, K, which="port"):
    """One read: rd = dict(seg_off, value, n0, n1, nm, nmm, ndel, nins, strand, flag, typeofaln, issec, supp, update_at).
    Returns dict(flag, typeofaln, issec, supp, mapq per segment; g_issec, g_value, g_n0, g_n1, g_nm[g,4], order per group)."""
    so = np.ascontiguousarray(rd["seg_off"], np.int32); G = len(so) - 1; S = int(so[-1])
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    L = ref() if which == "ref" else port()
    f = _bind_once(L, "ref_mapq" if which == "ref" else "lra_oracle_mapq", None,
                   [C.c_int, _i32p, f32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _u8p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int,
                    _i32p, _i32p, _u8p, _u8p, _i32p, _u8p, f32p, _i32p, _i32p, _i32p, _i32p])
    pad = lambda a, dt: np.ascontiguousarray(a, dt).copy() if len(a) else np.zeros(1, dt)
    i32 = lambda k: pad(rd[k], np.int32)
    o = dict(flag=i32("flag"), typeofaln=i32("typeofaln"), issec=pad(rd["issec"], np.uint8), supp=pad(rd["supp"], np.uint8), mapq=np.zeros(max(S, 1), np.int32),
             g_issec=np.zeros(max(G, 1), np.uint8), g_value=np.zeros(max(G, 1), np.float32), g_n0=np.zeros(max(G, 1), np.int32), g_n1=np.zeros(max(G, 1), np.int32),
             g_nm=np.zeros(4 * max(G, 1), np.int32), order=np.full(max(G, 1), -1, np.int32))
    ua = pad(rd["update_at"], np.int32)
    f(G, so, pad(rd["value"], np.float32), i32("n0"), i32("n1"), i32("nm"), i32("nmm"), i32("ndel"), i32("nins"), pad(rd["strand"], np.uint8), ua, len(rd["update_at"]),
      bypass, read_type, K, o["flag"], o["typeofaln"], o["issec"], o["supp"], o["mapq"], o["g_issec"], o["g_value"], o["g_n0"], o["g_n1"], o["g_nm"], o["order"])
    for k in ("flag", "typeofaln", "issec", "supp", "mapq"):
        o[k] = o[k][:S]
    for k in ("g_issec", "g_value", "g_n0", "g_n1", "order"):
        o[k] = o[k][:G]
    o["g_nm"] = o["g_nm"][:4 * G].reshape(-1, 4)
    return o


# ---------------------------------------------------------------- a15 LinearExtend (pairs) / DecideCoordinates / TrimOverlappedAnchors (clusters)

def linear_extend(read, genome, rd, K, skipsorting, trim, which="port"):
    """One read (ASCII bytes / uint8 array) against `genome` (uint8 arena).  rd = dict(g_off[n_groups+1], p_off[n_parts+1], p_strand, chrom_off,
    chrom_len (per part), q, t).  Returns dict(e_off, q, t, len, box[g,4], sorted_q, sorted_t)."""
    go = np.ascontiguousarray(rd["g_off"], np.int32); po_ = np.ascontiguousarray(rd["p_off"], np.int32)
    G = len(go) - 1; N = int(po_[-1])
    L = ref() if which == "ref" else port()
    f = _bind_once(L, "ref_linear_extend" if which == "ref" else "lra_oracle_linear_extend", C.c_long,
                   [_u8p, C.c_int, _u8p, _u64p, _i32p, C.c_int, _i32p, _i32p, _u8p, _u32p, _u32p, C.c_int, C.c_int, C.c_int, _i32p, _u32p, _u32p, _i32p, _u32p])
    pad = lambda a, dt: np.ascontiguousarray(a, dt).copy() if len(a) else np.zeros(1, dt)
    q = pad(rd["q"], np.uint32); t = pad(rd["t"], np.uint32)
    r = np.frombuffer(read, np.uint8) if isinstance(read, (bytes, bytearray)) else np.ascontiguousarray(read, np.uint8)
    o = dict(e_off=np.zeros(G + 1, np.int32), q=np.zeros(max(N, 1), np.uint32), t=np.zeros(max(N, 1), np.uint32), len=np.zeros(max(N, 1), np.int32),
             box=np.zeros(4 * max(G, 1), np.uint32))
    n = f(r, len(r), np.ascontiguousarray(genome, np.uint8), pad(rd["chrom_off"], np.uint64), pad(rd["chrom_len"], np.int32), G, go, po_, pad(rd["p_strand"], np.uint8),
          q, t, K, int(skipsorting), int(trim), o["e_off"], o["q"], o["t"], o["len"], o["box"])
    for k in ("q", "t", "len"):
        o[k] = o[k][:n]
    o["box"] = o["box"][:4 * G].reshape(-1, 4)
    o["sorted_q"], o["sorted_t"] = q[:N], t[:N]
    return o


def linear_extend_chain(read, genome, cd, chain, K, skiprepetitive=1, trim=1, merge_dist=100, which="port"):
    """High-accuracy overload for one chain.  cd = dict(cl_off, q, t, box[n,4], strand, chrom_off, chrom_len, freq); chain = cluster indices.
    Returns dict(e_off, q, t, len, ovp, box[e,4], overlap, md_off, md_start, md_end, sorted_q, sorted_t)."""
    co = np.ascontiguousarray(cd["cl_off"], np.int32); ncl = len(co) - 1
    ch = np.ascontiguousarray(chain, np.int32); E = len(ch)
    cap = int(sum(int(co[c + 1] - co[c]) for c in ch)) + 1
    L = ref() if which == "ref" else port()
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    f = _bind_once(L, "ref_linear_extend_chain" if which == "ref" else "lra_oracle_linear_extend_chain", C.c_long,
                   [_u8p, C.c_int, _u8p, C.c_int, _i32p, _u32p, _u32p, _u32p, _u8p, _u64p, _i32p, f32p, C.c_int, _i32p, C.c_int, C.c_int, C.c_int, C.c_int,
                    _i32p, _u32p, _u32p, _i32p, _u8p, _u32p, _i32p, _i32p, _i32p, _i32p])
    pad = lambda a, dt: np.ascontiguousarray(a, dt).reshape(-1).copy() if len(a) else np.zeros(1, dt)
    q = pad(cd["q"], np.uint32); t = pad(cd["t"], np.uint32)
    r = np.frombuffer(read, np.uint8) if isinstance(read, (bytes, bytearray)) else np.ascontiguousarray(read, np.uint8)
    o = dict(e_off=np.zeros(E + 1, np.int32), q=np.zeros(cap, np.uint32), t=np.zeros(cap, np.uint32), len=np.zeros(cap, np.int32), ovp=np.zeros(cap, np.uint8),
             box=np.zeros(4 * max(E, 1), np.uint32), md_off=np.zeros(E + 1, np.int32), md_start=np.zeros(cap + E, np.int32), md_end=np.zeros(cap + E, np.int32))
    ov = np.zeros(1, np.int32)
    n = f(r, len(r), np.ascontiguousarray(genome, np.uint8), ncl, co, q, t, pad(cd["box"], np.uint32), pad(cd["strand"], np.uint8), pad(cd["chrom_off"], np.uint64),
          pad(cd["chrom_len"], np.int32), pad(cd["freq"], np.float32), E, pad(ch, np.int32), K, int(skiprepetitive), int(trim), int(merge_dist),
          o["e_off"], o["q"], o["t"], o["len"], o["ovp"], o["box"], ov, o["md_off"], o["md_start"], o["md_end"])
    for k in ("q", "t", "len", "ovp"):
        o[k] = o[k][:n]
    nm = int(o["md_off"][E])
    o["md_start"] = o["md_start"][:nm]; o["md_end"] = o["md_end"][:nm]
    o["box"] = o["box"][:4 * E].reshape(-1, 4)
    o["overlap"] = int(ov[0])
    o["sorted_q"], o["sorted_t"] = q[:len(cd["q"])], t[:len(cd["t"])]
    return o


# ---------------------------------------------------------------- a11 SPLITChain (UltimateChain) / MergeSplitchainINS / RemoveSpuriousSplitChain

def split_chain(ch, hdr_pos, splitdist=50000, bypass=0, which="port"):
    """One chain: ch = dict(q, t, len, strand, cnum, link[n-1]).  Returns dict(sp_off, sptc, sp_lk, ci_off, ci, box[k,4], chrom, type, strand, link)."""
    n = len(ch["q"])
    L = ref() if which == "ref" else port()
    f = _bind_once(L, "ref_split_chain" if which == "ref" else "lra_oracle_split_chain", C.c_long,
                   [_u32p, _u32p, _i32p, _u8p, _i32p, _u8p, C.c_int, _u64p, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _u8p, _i32p, _i32p, _u32p, _i32p, _u8p, _u8p, _u8p, _i32p])
    pad = lambda a, dt: np.ascontiguousarray(a, dt).copy() if len(a) else np.zeros(1, dt)
    hdr = np.ascontiguousarray(hdr_pos, np.uint64)
    o = dict(sp_off=np.zeros(n + 2, np.int32), sptc=np.zeros(n + 1, np.int32), sp_lk=np.zeros(n + 1, np.uint8), ci_off=np.zeros(n + 2, np.int32), ci=np.zeros(n + 1, np.int32),
             box=np.zeros(4 * (n + 1), np.uint32), chrom=np.zeros(n + 1, np.int32), type=np.zeros(n + 1, np.uint8), strand=np.zeros(n + 1, np.uint8),
             link=np.zeros(n + 2, np.uint8))
    nl = np.zeros(1, np.int32)
    k = f(pad(ch["q"], np.uint32), pad(ch["t"], np.uint32), pad(ch["len"], np.int32), pad(ch["strand"], np.uint8), pad(ch["cnum"], np.int32), pad(ch["link"], np.uint8), n,
          hdr, len(hdr), int(splitdist), int(bypass), o["sp_off"], o["sptc"], o["sp_lk"], o["ci_off"], o["ci"], o["box"], o["chrom"], o["type"], o["strand"], o["link"], nl)
    o["sp_off"] = o["sp_off"][:k + 1]; o["ci_off"] = o["ci_off"][:k + 1]
    m = int(o["sp_off"][k]); o["sptc"] = o["sptc"][:m]; o["sp_lk"] = o["sp_lk"][:m]; o["ci"] = o["ci"][:int(o["ci_off"][k])]
    o["box"] = o["box"][:4 * k].reshape(-1, 4)
    for key in ("chrom", "type", "strand"):
        o[key] = o[key][:k]
    o["link"] = o["link"][:int(nl[0])]
    return o


# ---------------------------------------------------------------- a11 MergeChain / switchindex

def merge_chain(sp, chrom, strand, box, which="port"):
    """One split chain over clusters.  Returns head flags (1 = the entry starts a new Merge_SplitChain)."""
    n = len(sp)
    L = ref() if which == "ref" else port()
    f = _bind_once(L, "ref_merge_chain" if which == "ref" else "lra_oracle_merge_chain", C.c_long, [_i32p, C.c_int, _i32p, _u8p, _u32p, _u8p])
    head = np.zeros(max(n, 1), np.uint8)
    g = f(np.ascontiguousarray(sp, np.int32) if n else np.zeros(1, np.int32), n, np.ascontiguousarray(chrom, np.int32), np.ascontiguousarray(strand, np.uint8),
          np.ascontiguousarray(box, np.uint32).reshape(-1), head)
    assert g == int(head[:n].sum())
    return head[:n]


def switchindex(ch, link, coarse, cq, which="port"):
    """One chain over split clusters (link: len(ch) - 1 bits).  Returns (chain over clusters, links)."""
    n = len(ch)
    c = np.ascontiguousarray(ch, np.int32).copy() if n else np.zeros(1, np.int32)
    l = np.zeros(max(n, 1), np.uint8); l[:len(link)] = link
    nl = np.zeros(1, np.int32)
    coarse = np.ascontiguousarray(coarse, np.int32); cq = np.ascontiguousarray(cq, np.uint32).reshape(-1)
    if which == "ref":
        f = _bind_once(ref(), "ref_switchindex", C.c_long, [_i32p, C.c_int, _u8p, C.c_int, _i32p, C.c_int, _u32p, C.c_int, _i32p])
        m = f(c, n, l, len(link), coarse, len(coarse), cq, len(cq) // 2, nl)
    else:
        f = _bind_once(port(), "lra_oracle_switchindex", C.c_long, [_i32p, C.c_int, _u8p, C.c_int, _i32p, _u32p, _i32p])
        m = f(c, n, l, len(link), coarse, cq, nl)
    return c[:m].copy(), l[:int(nl[0])].copy()


# ---------------------------------------------------------------- a17 (leaf) RefineByLinearAlignment

def refine_linear(read, contig, qs, qe, ts, te, m, mm, indel, local_band, which="port"):
    """One gap: read[qs, qe) against contig[ts, te).  Returns the blocks [n,3] in read / contig coordinates.
    port: SetMatchAndGaps / Matched (LocalRefineAlignment.h:89-99) in uint32 arithmetic, AlignSubstrings' band (:101-129), the a18 oracle, the shift of
    RefineSubstrings (:131-142)."""
    r = np.ascontiguousarray(read, np.uint8); c = np.ascontiguousarray(contig, np.uint8)
    if which == "ref":
        f = _bind_once(ref(), "ref_refine_linear", C.c_long, [_u8p, C.c_int, _u8p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int,
                                                              _u32p, C.c_long])
        cap = max(int(qe) - int(qs), 1) + 16
        out = np.zeros(3 * cap, np.uint32)
        n = f(r, len(r), c, len(c), int(qs), int(qe), int(ts), int(te), m, mm, indel, local_band, out, cap)
        return out[:3 * n].reshape(-1, 3)
    a = (int(qe) - int(qs) + 1) & 0xFFFFFFFF; b = (int(te) - int(ts) + 1) & 0xFFFFFFFF
    mt = min(a, b)
    if mt >= 1 << 31:
        mt -= 1 << 32
    if mt <= 0:
        return np.zeros((0, 3), np.uint32)
    ql, tl = int(qe) - int(qs), int(te) - int(ts)
    k = min(abs(ql - tl) * 2 + 1, local_band)
    score, blocks, st = aog_port(bytes(r[qs:qe]), bytes(c[ts:te]), m, mm, indel, k)
    assert st == 0
    blocks[:, 0] += np.uint32(qs); blocks[:, 1] += np.uint32(ts)
    return blocks


# ---------------------------------------------------------------- a14 (core, small spaces) RefineSpace

def _refine_space_large(r, q, t, t0, K, W, diag, local_max_freq, qs, qe, ts, te, st, consider_str, lrts):
    """The branch of RefineSpace for a space of 1000 bases or more on either axis (ClusterRefine.h:296-305): non-canonical minimizers of both windows,
    std::sort, CompareLists(Global = false) inside the diagonal band [min(0, d2) - diag, max(0, d2) + diag], d2 = (te - (ts - lrts)) - (qe - qs); identity = -1."""
    P = port()
    f = _bind_once(P, "lra_oracle_store_minimizers_nc", C.c_long, [_u8p, C.c_uint32, C.c_int, C.c_int, _u64p, _u32p, C.c_long])
    g = _bind_once(P, "lra_oracle_compare_lists_band", C.c_long, [_u64p, _u32p, C.c_long, _u64p, _u32p, C.c_long, C.c_long, C.c_long, C.c_long, _u64p, _u32p, _u64p, _u32p, C.c_long])
    srt = _bind_once(P, "lra_oracle_sort_minimizers", None, [_u64p, _u32p, C.c_long])

    def mins(seq):
        buf = np.concatenate([np.ascontiguousarray(seq, np.uint8), np.zeros(8, np.uint8)])
        cap = len(seq) + 8
        tt = np.zeros(cap, np.uint64); pp = np.zeros(cap, np.uint32)
        n = f(buf, len(seq), K, W, tt, pp, cap)
        tt, pp = tt[:n].copy(), pp[:n].copy()
        if n:
            srt(tt, pp, n)
        return tt, pp
    gt, gp = mins(t); qt, qp = mins(q)
    d2 = (int(te) - (int(ts) - int(lrts))) - (int(qe) - int(qs))
    mn, mx = min(0, d2) - diag, max(0, d2) + diag
    pq, pt = np.zeros(0, np.uint32), np.zeros(0, np.uint32)
    if len(qt) and len(gt):
        cap = 4 * (len(qt) + len(gt)) + 1024
        while True:
            r4 = [np.zeros(cap, np.uint64), np.zeros(cap, np.uint32), np.zeros(cap, np.uint64), np.zeros(cap, np.uint32)]
            n = g(qt, qp, len(qt), gt, gp, len(gt), local_max_freq, mx, mn, r4[0], r4[1], r4[2], r4[3], cap)
            if n <= cap:
                break
            cap = n + 16
        pq = r4[1][:n].astype(np.int64) + int(qs); pt = r4[3][:n].astype(np.int64) + t0
        if consider_str and st == 1:
            pq = len(r) - pq - K
        pq = (pq & 0xFFFFFFFF).astype(np.uint32); pt = (pt & 0xFFFFFFFF).astype(np.uint32)
    return pq, pt, np.float32(-1.0)


def refine_space(strandseq, contig, K, qs, qe, ts, te, st, consider_str, lrts, lrlength, m, mm, indel, which="port", W=10, diag=100, local_max_freq=30):
    """One space.  Returns (pq, pt, identity).  port: the AffineOneGapAlign branch of RefineSpace (ClusterRefine.h:262-294) on the a18 oracle, the K-mer
    harvest, identity in binary32, the coordinate shift (:314-325); spaces of 1000 or more are outside the restatement."""
    r = np.ascontiguousarray(strandseq, np.uint8); c = np.ascontiguousarray(contig, np.uint8)
    if which == "ref":
        f = _bind_once(ref(), "ref_refine_space", C.c_long, [_u8p, C.c_int, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                             C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, _u32p, _u32p, C.c_long, C.POINTER(C.c_float), C.c_int])
        cap = 8 * (int(qe) - int(qs)) + 1024
        pq = np.zeros(cap, np.uint32); pt = np.zeros(cap, np.uint32); idn = C.c_float(0)
        n = f(r, len(r), c, len(c), K, W, diag, int(consider_str), int(qe), int(qs), int(te), int(ts), int(st), int(lrts), int(lrlength), m, mm, indel, pq, pt, cap, C.byref(idn), local_max_freq)
        return pq[:n], pt[:n], np.float32(idn.value)
    q = r[qs:qe]; t0 = int(ts) - int(lrts); t = c[t0:t0 + int(te) - int(ts) + int(lrlength)]
    if not (len(q) < 1000 and len(t) < 1000):
        return _refine_space_large(r, q, t, t0, K, W, diag, local_max_freq, qs, qe, ts, te, st, consider_str, lrts)
    score, blocks, stt = aog_port(bytes(q), bytes(t), m, mm, indel, 30)
    assert stt == 0
    n_match = 0; pq, pt = [], []
    for bq, bt, ln in blocks.tolist():
        n_match += int((q[bq:bq + ln] == t[bt:bt + ln]).sum())
        if ln > K:
            for bp in range(0, ln - K, K):              # bp + K < length
                if (q[bq + bp:bq + bp + K] == t[bt + bp:bt + bp + K]).all():
                    fq = bq + bp + int(qs)
                    if consider_str and st == 1:
                        fq = (len(r) - fq - K) & 0xFFFFFFFF
                    pq.append(fq); pt.append(bt + bp + t0)
    mn = min(len(q), len(t))
    idn = np.float32(n_match) / np.float32(mn) if mn else np.frombuffer(np.uint32(0xFFC00000).tobytes(), np.float32)[0]
    return np.array(pq, np.uint32), np.array(pt, np.uint32), np.float32(idn)


# ---------------------------------------------------------------- a17 SwitchToOriginalAnchors

def switch_to_original(cnum, k, run_off, start, end, coarse, which="port"):
    """One FinalChain: entry i names run k[i] of cluster cnum[i] (runs of cluster c: start/end[run_off[c] .. run_off[c+1])).  Returns (chain, ClusterIndex).
    port: LocalRefineAlignment.h:187-198 restated -- every run expands to its anchors end-1 .. start, ClusterIndex = the cluster's coarse."""
    cnum = np.ascontiguousarray(cnum, np.int32); k = np.ascontiguousarray(k, np.int32); run_off = np.ascontiguousarray(run_off, np.int32)
    start = np.ascontiguousarray(start, np.int32); end = np.ascontiguousarray(end, np.int32); coarse = np.ascontiguousarray(coarse, np.int32)
    if which == "ref":
        f = _bind_once(ref(), "ref_switch_to_original", C.c_long, [_i32p, _i32p, C.c_int, _i32p, _i32p, _i32p, _i32p, C.c_int, _u32p, _i32p])
        cap = int(sum(int(end[run_off[c] + r] - start[run_off[c] + r]) for c, r in zip(cnum, k))) + 1
        ch = np.zeros(cap, np.uint32); ci = np.zeros(cap, np.int32)
        n = f(cnum, k, len(cnum), run_off, start, end, coarse, len(coarse), ch, ci)
        return ch[:n], ci[:n]
    ch, ci = [], []
    for c, r in zip(cnum, k):
        for j in range(int(end[run_off[c] + r]) - 1, int(start[run_off[c] + r]) - 1, -1):
            ch.append(j); ci.append(int(coarse[c]))
    return np.array(ch, np.uint32), np.array(ci, np.int32)


# ---------------------------------------------------------------- a8 (first half) SplitRoughClustersWithGaps

def split_rough(q, t, rc, globalK, maxGap, minClusterSize, maxDiag, which="port"):
    """One anchor list (Cartesian-sorted inside every rough cluster).  rc = dict(start, end, box[n,4], strand, freq, chrom).
    Returns dict(start, end, box, strand, coarse, freq, chrom, smi = [splitmatchindex of every split cluster])."""
    q = np.ascontiguousarray(q, np.uint32); t = np.ascontiguousarray(t, np.uint32)
    n = len(rc["start"]); N = len(q)
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    L = ref() if which == "ref" else port()
    f = _bind_once(L, "ref_split_rough" if which == "ref" else "lra_oracle_split_rough", C.c_long,
                   [_u32p, _u32p, C.c_int, _i32p, _i32p, _u32p, _u8p, f32p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int,
                    _i32p, _i32p, _u32p, _u8p, _i32p, f32p, _i32p, _i32p, _i32p, _i32p, _i32p])
    pad = lambda a, dt: np.ascontiguousarray(a, dt).reshape(-1).copy() if len(a) else np.zeros(1, dt)
    cap = N + n + 1
    o = dict(start=np.zeros(cap, np.int32), end=np.zeros(cap, np.int32), box=np.zeros(4 * cap, np.uint32), strand=np.zeros(cap, np.uint8), coarse=np.zeros(cap, np.int32),
             freq=np.zeros(cap, np.float32), chrom=np.zeros(cap, np.int32))
    pc = np.zeros(cap, np.int32); ps = np.zeros(cap, np.int32); pe = np.zeros(cap, np.int32); npiece = np.zeros(1, np.int32)
    ns = f(pad(q, np.uint32), pad(t, np.uint32), n, pad(rc["start"], np.int32), pad(rc["end"], np.int32), pad(rc["box"], np.uint32), pad(rc["strand"], np.uint8),
           pad(rc["freq"], np.float32), pad(rc["chrom"], np.int32), globalK, maxGap, minClusterSize, maxDiag,
           o["start"], o["end"], o["box"], o["strand"], o["coarse"], o["freq"], o["chrom"], pc, ps, pe, npiece)
    for k in ("start", "end", "strand", "coarse", "freq", "chrom"):
        o[k] = o[k][:ns]
    o["box"] = o["box"][:4 * ns].reshape(-1, 4)
    smi = [[] for _ in range(ns)]
    for j in range(int(npiece[0])):
        smi[pc[j]] += list(range(int(ps[j]), int(pe[j])))
    o["smi"] = smi
    o["n_piece"] = int(npiece[0])
    return o


# ---------------------------------------------------------------- StoreDiagonalClusters (CleanMatches without ExtractDiagonalFromClean)

def store_diagonal(q, t, qt, freq, strand, hdr_pos, globalK, maxDiag, minClusterSize, minClusterLength, bypass, which="port"):
    """One cleaned, diagonal-sorted anchor list.  Returns dict(start, end, box[k,4], freq, chrom)."""
    n = len(q)
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    L = ref() if which == "ref" else port()
    f = _bind_once(L, "ref_store_diagonal" if which == "ref" else "lra_oracle_store_diagonal", C.c_long,
                   [_u32p, _u32p, _u64p, f32p, C.c_int, C.c_int, _u64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _u32p, f32p, _i32p])
    pad = lambda a, dt: np.ascontiguousarray(a, dt) if n else np.zeros(1, dt)
    hdr = np.ascontiguousarray(hdr_pos, np.uint64)
    o = dict(start=np.zeros(n + 1, np.int32), end=np.zeros(n + 1, np.int32), box=np.zeros(4 * (n + 1), np.uint32), freq=np.zeros(n + 1, np.float32), chrom=np.zeros(n + 1, np.int32))
    k = f(pad(q, np.uint32), pad(t, np.uint32), pad(qt, np.uint64), pad(freq, np.float32), n, int(strand), hdr, len(hdr), globalK, maxDiag, minClusterSize, minClusterLength,
          int(bypass), o["start"], o["end"], o["box"], o["freq"], o["chrom"])
    for key in ("start", "end", "freq", "chrom"):
        o[key] = o[key][:k]
    o["box"] = o["box"][:4 * k].reshape(-1, 4)
    return o


# ---------------------------------------------------------------- TrimSplitChainDiagonal

def trim_splitchain(cq, ct, strand, q, t, which="port"):
    """One split chain (chain anchors cq / ct in sptc order) and its refined anchors.  Returns (kept q, kept t, nRemoved)."""
    cq = np.ascontiguousarray(cq, np.uint32); ct = np.ascontiguousarray(ct, np.uint32)
    n = len(q)
    qq = np.ascontiguousarray(q, np.uint32).copy() if n else np.zeros(1, np.uint32); tt = np.ascontiguousarray(t, np.uint32).copy() if n else np.zeros(1, np.uint32)
    if which == "ref":
        f = _bind_once(ref(), "ref_trim_splitchain", C.c_long, [_u32p, _u32p, C.c_int, C.c_int, _u32p, _u32p, C.c_int, C.POINTER(C.c_long)])
        rem = C.c_long(0)
        k = f(cq, ct, len(cq), int(strand), qq, tt, n, C.byref(rem))
        return qq[:k].copy(), tt[:k].copy(), int(rem.value)
    f = _bind_once(port(), "lra_oracle_trim_splitchain", C.c_long, [_u32p, _u32p, C.c_int, C.c_int, _u32p, _u32p, C.c_int, _u8p])
    keep = np.zeros(max(n, 1), np.uint8)
    rem = f(cq, ct, len(cq), int(strand), qq, tt, n, keep)
    m = keep[:n].astype(bool)
    return qq[:n][m], tt[:n][m], int(rem)
