"""Builds and binds tests/simt/libemu_lra.so: the device code of lra_b200/csrc/*.cuh compiled for the CPU through the
lock-step SIMT emulator (TEST INFRASTRUCTURE; lets the CPU suite check kernel logic where no GPU exists)."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SIMT = os.path.join(HERE, "simt")
CSRC = os.path.join(os.path.dirname(HERE), "lra_b200", "csrc")
_lib = None
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def set_lane_order(mode):
    lib().emu_set_lane_order(C.c_long(mode))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    so = os.path.join(SIMT, "libemu_lra.so")
    srcs = [os.path.join(SIMT, f) for f in os.listdir(SIMT) if f.endswith((".cpp", ".h"))]
    srcs += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        cpps = [s for s in srcs if s.endswith(".cpp")]
        subprocess.run(["g++", "-std=c++17", "-O1", "-DLRA_EMU", "-I" + SIMT, "-I" + CSRC, "-fPIC", "-shared"] + cpps +
                       ["-o", so], check=True)
    L = C.CDLL(so)
    L.emu_aog_batch.restype = C.c_int
    L.emu_aog_batch.argtypes = [_u8p, C.c_uint64, _u8p, C.c_uint64, _u32p, _u32p, _i32p, _i32p, _i32p, C.c_int, C.c_int,
                                C.c_int, C.c_int, _i32p, _i32p, _u64p, _u32p, C.c_uint64, C.c_int, C.c_int,
                                C.POINTER(C.c_uint64)]
    L.emu_seq_pack.restype = C.c_int
    L.emu_seq_pack.argtypes = [_u8p, C.c_uint64, _u32p, _u32p]
    L.emu_ir_dp_batch.restype = C.c_int
    L.emu_ir_dp_batch.argtypes = [_u8p, C.c_uint64, _u8p, C.c_uint64, _u32p, _u32p, _i32p, _i32p, _i32p, _i32p, _i32p, _u32p, _i32p,
                                  C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _u64p, _u32p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
    L.emu_ir_segments.restype = C.c_int
    L.emu_ir_segments.argtypes = [_u8p, C.c_uint64, _u8p, C.c_uint64, _u32p, _u64p, _i32p, _u32p, _u32p, _i32p, _i32p, C.c_uint64, C.c_int,
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _u64p, _u32p, C.c_uint64, _u64p]
    L.emu_seed_batch.restype = C.c_long
    L.emu_seed_batch.argtypes = [_u8p, C.c_uint64, _u64p, _u32p, C.c_int, _u8p, C.c_uint64, _u64p, _u32p, C.c_long, C.c_int, C.c_int, C.c_long,
                                 _u64p, _u64p, _u32p, _u64p, _u32p, _u8p, C.c_uint64, _u32p, _u64p, _u32p]
    L.emu_seq_revcomp.restype = C.c_int
    L.emu_seq_revcomp.argtypes = [_u8p, C.c_uint64, _u64p, _u32p, C.c_int, _u32p, _u32p]
    L.emu_calc_stats.restype = C.c_long
    L.emu_calc_stats.argtypes = [_u8p, C.c_uint64, _u8p, C.c_uint64, _u32p, _u64p, _i32p, _u32p, _u32p, _i32p, C.c_int,
                                 np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS"), _i32p, np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS"),
                                 _u64p, _u32p, C.c_uint64, C.c_int]
    _lib = L
    return L


def calc_stats(sb, t_arena, t_base, lut, thread_kernels=0):
    S = len(sb["blk_cnt"])
    bi = np.ascontiguousarray(sb["blocks_in"], np.uint32)
    cap = 4 * (bi.size // 3) + 16 * S + 16
    o = dict(stats=np.zeros((S, 16), np.int32), value=np.zeros(S, np.float32), cigar_off=np.zeros(S + 1, np.uint64), cigar=np.zeros(cap, np.uint32))
    lib().emu_calc_stats(sb["q_arena"], len(sb["q_arena"]) - 16, t_arena, len(t_arena) - 16, bi.reshape(-1), np.ascontiguousarray(sb["blk_off"], np.uint64),
                         sb["blk_cnt"], sb["q_base"], np.ascontiguousarray(t_base, np.uint32), sb["read_len"], S, np.ascontiguousarray(lut, np.float32),
                         o["stats"].reshape(-1), o["value"], o["cigar_off"], o["cigar"], cap, thread_kernels)
    return o


def seed_batch(reads_arena, read_off, read_len, genome, idx_t, idx_pos, k, w, max_freq, cap=None):
    """Returns dict(match_off, q_t, q_pos, t_t, t_pos, strand, n_mm, mm_t, mm_pos)."""
    R = len(read_off)
    rn = len(reads_arena) - 16
    cap = cap or (8 * rn + 1024)
    o = dict(match_off=np.zeros(R + 1, np.uint64), q_t=np.zeros(cap, np.uint64), q_pos=np.zeros(cap, np.uint32), t_t=np.zeros(cap, np.uint64),
             t_pos=np.zeros(cap, np.uint32), strand=np.zeros(cap, np.uint8), n_mm=np.zeros(R, np.uint32), mm_t=np.zeros(rn + 8, np.uint64),
             mm_pos=np.zeros(rn + 8, np.uint32))
    n = lib().emu_seed_batch(reads_arena, rn, np.ascontiguousarray(read_off, np.uint64), np.ascontiguousarray(read_len, np.uint32), R, genome,
                             len(genome) - 16, np.ascontiguousarray(idx_t, np.uint64), np.ascontiguousarray(idx_pos, np.uint32), len(idx_t), k, w,
                             max_freq, o["match_off"], o["q_t"], o["q_pos"], o["t_t"], o["t_pos"], o["strand"], cap, o["n_mm"], o["mm_t"], o["mm_pos"])
    o["n"] = n
    return o


def seq_revcomp(arena, read_off, read_len):
    n = len(arena) - 16
    b2 = np.zeros((n + 15) // 16, np.uint32); nm = np.zeros((n + 31) // 32, np.uint32)
    lib().emu_seq_revcomp(arena, n, np.ascontiguousarray(read_off, np.uint64), np.ascontiguousarray(read_len, np.uint32), len(read_off), b2, nm)
    return b2, nm


def ir_segments(sb, out_cap=None):
    """sb: dict from irgen.pack_segments.  Returns (err, n_blocks, block_off, blocks, info[n_aog,n_groups,cells])."""
    S = len(sb["blk_cnt"])
    if out_cap is None:
        out_cap = int(sb["read_len"].sum()) + 16
    n = np.zeros(S, np.int32); off = np.zeros(S, np.uint64); blocks = np.zeros(out_cap * 3, np.uint32); info = np.zeros(3, np.uint64)
    err = lib().emu_ir_segments(sb["q_arena"], len(sb["q_arena"]) - 16, sb["t_arena"], len(sb["t_arena"]) - 16, sb["blocks_in"].reshape(-1),
                                sb["blk_off"], sb["blk_cnt"], sb["q_base"], sb["t_base"], sb["read_len"], sb["contig_len"],
                                len(sb["blocks_in"]), S, sb["k"], sb["match"], sb["mismatch"], sb["indel"], sb["end_align"], n, off,
                                blocks, out_cap, info)
    return err, n, off, blocks.reshape(-1, 3), info


def ir_dp_batch(gb, force_generic=0, block_cap=None):
    """gb: dict from irgen.pack_groups.  Returns (err, n_blocks, block_off, blocks, cells)."""
    n = len(gb["t_len"])
    if block_cap is None:
        block_cap = int(gb["q_seq_len"].sum() + gb["t_seq_len"].sum()) + 8
    nb = np.zeros(n, np.int32); off = np.zeros(n, np.uint64); blocks = np.zeros(block_cap * 3, np.uint32)
    cells = C.c_uint64(0)
    err = lib().emu_ir_dp_batch(gb["q_arena"], len(gb["q_arena"]) - 16, gb["t_arena"], len(gb["t_arena"]) - 16, gb["q_base"],
                                gb["t_base"], gb["q_start"], gb["t_start"], gb["t_len"], gb["q_seq_len"], gb["t_seq_len"],
                                gb["band_off"], gb["band"], n, gb["match"], gb["mismatch"], gb["indel"], nb, off, blocks, block_cap,
                                force_generic, C.byref(cells))
    return err, nb, off, blocks.reshape(-1, 3), cells.value


def aog_batch(qa, ta, qo, to, ql, tl, k, m, mm, indel, use_band=1, force_literal=0, block_cap=None):
    n = len(qo)
    if block_cap is None:
        block_cap = int((np.minimum(ql, tl) + 1).sum())
    score = np.zeros(n, np.int32); nb = np.zeros(n, np.int32); off = np.zeros(n, np.uint64)
    blocks = np.zeros(max(1, block_cap) * 3, np.uint32)
    cells = C.c_uint64(0)
    err = lib().emu_aog_batch(qa, len(qa) - 16, ta, len(ta) - 16, qo, to, ql, tl, k, n, m, mm, indel, score, nb, off, blocks,
                              block_cap, use_band, force_literal, C.byref(cells))
    return err, score, nb, off, blocks.reshape(-1, 3), cells.value


# ---------------------------------------------------------------- a12 / a13

class EmuLidx(C.Structure):
    _fields_ = [("win_off", C.c_void_p), ("win_len", C.c_void_p), ("bnd", C.c_void_p), ("mins", C.c_void_p), ("win_first", C.c_void_p),
                ("seq_start", C.c_void_p), ("seq_len", C.c_void_p), ("n_win", C.c_int32), ("n_seq", C.c_int32)]


def lindex_layout(seq_start, seq_len, window=2048):
    """The window table of lra_b200/csrc/lref_host.cuh: lindex_layout."""
    win_off, win_len, win_first = [], [], []
    for s, L in zip(seq_start, seq_len):
        win_first.append(len(win_off))
        for p in range(0, int(L), window):
            win_off.append(int(s) + p); win_len.append(min(window, int(L) - p))
    win_first.append(len(win_off))
    win_off.append(int(seq_start[-1]) + int(seq_len[-1]) if len(seq_start) else 0)
    return np.array(win_off, np.uint64), np.array(win_len + [0], np.uint32), np.array(win_first, np.uint32)


def lindex_build(arena, seq_start, seq_len, k=10, w=5, window=2048, max_freq=15):
    """arena: ASCII + 16 bytes of padding.  Returns a dict image (win_off, win_len, bnd, mins, win_first, seq_start, seq_len)."""
    L = lib()
    L.emu_lindex_build.restype = C.c_long
    L.emu_lindex_build.argtypes = [_u8p, C.c_uint64, _u64p, _u32p, C.c_int, C.c_int, C.c_int, C.c_int, _u64p, _u32p]
    seq_start = np.ascontiguousarray(seq_start, np.uint64); seq_len = np.ascontiguousarray(seq_len, np.uint32)
    win_off, win_len, win_first = lindex_layout(seq_start, seq_len, window)
    nw = len(win_off) - 1
    bnd = np.zeros(nw + 2, np.uint64); mins = np.zeros(len(arena) + 16, np.uint32)
    n = L.emu_lindex_build(arena, len(arena) - 16, win_off, win_len, nw, k, w, max_freq, bnd, mins)
    return dict(win_off=win_off, win_len=win_len, bnd=bnd[:nw + 1].copy(), mins=mins[:max(n, 1)].copy(), n_mins=n, win_first=win_first, seq_start=seq_start,
                seq_len=seq_len, n_win=nw, n_seq=len(seq_start))


def _emu_lidx(img):
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    return EmuLidx(p(img["win_off"]), p(img["win_len"]), p(img["bnd"]), p(img["mins"]), p(img["win_first"]), p(img["seq_start"]), p(img["seq_len"]),
                   img["n_win"], img["n_seq"])


def refine_clusters(gl, rf, rr, cl, cap=None, literal=0):
    """cl: dict(m_q, m_t, m_off, box[n,4], strand, read_id, hdr_pos, global_k, small_k, window, local_max_freq).  Returns a result dict."""
    L = lib()
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS"); i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
    L.emu_refine_clusters.restype = C.c_long
    L.emu_refine_clusters.argtypes = [C.POINTER(EmuLidx)] * 3 + [C.c_int, _u32p, _u32p, _u64p, _u32p, _u8p, _u32p, _u64p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                                 C.c_long, _i32p, _i32p, i64p, _u64p, _u32p, _u32p, _u32p, C.c_uint64, _u32p, f32p, _u32p, _u32p,
                                                                 _u32p, _u64p, C.c_int, C.c_int, _u32p, _u8p, _i32p, C.c_int]
    n = len(cl["strand"])
    M = int(cl["m_off"][n])
    cap = cap or 1 << 18
    pad = lambda a, dt: np.ascontiguousarray(a, dt) if len(a) else np.zeros(1, dt)
    while True:
        o = dict(status=np.zeros(n, np.int32), chrom=np.zeros(n, np.int32), diag=np.zeros(2 * n, np.int64), r_off=np.zeros(n + 1, np.uint64),
                 r_q=np.zeros(cap, np.uint32), r_t=np.zeros(cap, np.uint32), r_tup=np.zeros(cap, np.uint32), rbox=np.zeros(4 * n, np.uint32),
                 eff=np.zeros(n, np.float32), m_q_out=np.zeros(M + 1, np.uint32), m_t_out=np.zeros(M + 1, np.uint32), box_out=np.zeros(4 * n, np.uint32))
        counts = np.zeros(2, np.uint64)
        a, b_, c = _emu_lidx(gl), _emu_lidx(rf), _emu_lidx(rr)
        tot = L.emu_refine_clusters(C.byref(a), C.byref(b_), C.byref(c), n, pad(cl["m_q"], np.uint32), pad(cl["m_t"], np.uint32),
                                    np.ascontiguousarray(cl["m_off"], np.uint64), np.ascontiguousarray(cl["box"], np.uint32).reshape(-1),
                                    np.ascontiguousarray(cl["strand"], np.uint8), np.ascontiguousarray(cl["read_id"], np.uint32),
                                    np.ascontiguousarray(cl["hdr_pos"], np.uint64), len(cl["hdr_pos"]), cl["global_k"], cl["small_k"], cl["window"],
                                    cl["local_max_freq"], o["status"], o["chrom"], o["diag"], o["r_off"], o["r_q"], o["r_t"], o["r_tup"], cap,
                                    o["rbox"], o["eff"], o["m_q_out"], o["m_t_out"], o["box_out"], counts, literal,
                                    1 if "m_len" in cl else 0, pad(cl.get("m_len", []), np.uint32), pad(cl.get("m_strand", []), np.uint8),
                                    pad(cl.get("chrom", []), np.int32), cl.get("limitrefine", 1))
        if tot <= cap:
            o["n_anchors"] = tot; o["n_units"], o["n_tasks"] = int(counts[0]), int(counts[1])
            return o
        cap = tot


def sort_matches(mode, q, t, seg_off):
    L = lib()
    L.emu_sort_matches.argtypes = [C.c_int, _u32p, _u32p, _u64p, C.c_int, _u32p]
    q = np.array(q, np.uint32); t = np.array(t, np.uint32); perm = np.zeros(max(len(q), 1), np.uint32)
    seg_off = np.ascontiguousarray(seg_off, np.uint64)
    L.emu_sort_matches(mode, q if len(q) else np.zeros(1, np.uint32), t if len(t) else np.zeros(1, np.uint32), seg_off, len(seg_off) - 1, perm)
    return q, t, perm[:len(q)]


def global_chain(frag, frag_off, score):
    L = lib()
    L.emu_global_chain.argtypes = [_i32p, _u64p, C.c_int, _i32p, _i32p, _i32p, _i32p]
    f = np.ascontiguousarray(frag, np.int32).reshape(-1); fo = np.ascontiguousarray(frag_off, np.uint64)
    n = len(f) // 4
    sc = np.array(score, np.int32); prev = np.zeros(max(n, 1), np.int32); chain = np.zeros(max(n, 1), np.int32); cl = np.zeros(max(len(fo) - 1, 1), np.int32)
    L.emu_global_chain(f if n else np.zeros(4, np.int32), fo, len(fo) - 1, sc if n else np.zeros(1, np.int32), prev, chain, cl)
    return dict(score=sc, prev=prev[:n], chain=chain, chain_len=cl[:len(fo) - 1])


def refine_breakpoint(fwd, rcs, genome, bp):
    """fwd / rcs / genome: ASCII arenas with 16 bytes of padding; bp as for Context.refine_breakpoint_batch."""
    L = lib()
    L.emu_refine_breakpoint.argtypes = [_u8p, _u8p, C.c_uint64, _u8p, C.c_uint64, C.c_int, _u32p, _u32p, _u32p, _u32p, _u8p, _u8p, _u64p, _u32p, _u64p, _u64p, _u32p,
                                        _u32p, _i32p, _i32p, _u32p, _u32p, _i32p]
    n = len(bp["lstrand"])
    u32 = lambda k: np.ascontiguousarray(bp[k], np.uint32).reshape(-1)
    u64 = lambda k: np.ascontiguousarray(bp[k], np.uint64)
    o = dict(mode=np.zeros((n, 2), np.int32), n_out=np.zeros((n, 2), np.int32), bound=np.zeros((n, 2, 3), np.uint32), out=np.zeros((n, 2, 512, 3), np.uint32),
             refined=np.zeros(n, np.int32))
    L.emu_refine_breakpoint(fwd, rcs, len(fwd) - 16, genome, len(genome) - 16, n, u32("lf"), u32("ll"), u32("rf"), u32("rl"), np.ascontiguousarray(bp["lstrand"], np.uint8),
                            np.ascontiguousarray(bp["rstrand"], np.uint8), u64("read_off"), u32("read_len"), u64("lchrom_off"), u64("rchrom_off"), u32("lchrom_len"),
                            u32("rchrom_len"), o["mode"].reshape(-1), o["n_out"].reshape(-1), o["bound"].reshape(-1), o["out"].reshape(-1), o["refined"])
    return o


def chain_filter(mode, q, t, length, strand, chain_off):
    L = lib()
    L.emu_chain_filter.argtypes = [C.c_int, _u32p, _u32p, _u32p, _u8p, _u64p, C.c_int, _u8p]
    n = len(q)
    pad = lambda a, dt: np.ascontiguousarray(a, dt) if n else np.zeros(1, dt)
    keep = np.zeros(max(n, 1), np.uint8)
    co = np.ascontiguousarray(chain_off, np.uint64)
    L.emu_chain_filter(mode, pad(q, np.uint32), pad(t, np.uint32), pad(length, np.uint32), pad(strand, np.uint8), co, len(co) - 1, keep)
    return keep[:n]


def clean_off_diagonal(q, t, qt, list_off, strand, opt_values, hdr_pos):
    """opt_values: the ten Options fields in the order of lra_b200_clean_opts."""
    L = lib()
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    L.emu_clean_off_diagonal.argtypes = [_u32p, _u32p, _u64p, _u64p, _u8p, C.c_int, _i32p, _u64p, C.c_int, _u8p, f32p, _i32p, _i32p, f32p, _i32p]
    N, n = len(q), len(list_off) - 1
    pad = lambda a, dt: np.ascontiguousarray(a, dt) if len(a) else np.zeros(1, dt)
    o = dict(keep=np.zeros(max(N, 1), np.uint8), freq=np.zeros(max(N, 1), np.float32), cnt=np.zeros(max(N, 1), np.int32), cl=np.zeros(7 * max(N, 1), np.int32),
             cl_freq=np.zeros(max(N, 1), np.float32), n_cl=np.zeros(max(n, 1), np.int32))
    hdr = np.ascontiguousarray(hdr_pos, np.uint64)
    L.emu_clean_off_diagonal(pad(q, np.uint32), pad(t, np.uint32), pad(qt, np.uint64), np.ascontiguousarray(list_off, np.uint64), pad(strand, np.uint8), n,
                             np.ascontiguousarray(opt_values, np.int32), hdr, len(hdr), o["keep"], o["freq"], o["cnt"], o["cl"], o["cl_freq"], o["n_cl"])
    o["cl"] = o["cl"].reshape(-1, 7)
    return {k: (v[:n] if k == "n_cl" else v[:N]) for k, v in o.items()}


def split_clusters(cl_off, box, strand, freq, m_off, mq, contig, global_k, cap=1 << 16):
    L = lib()
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    L.emu_split_clusters.restype = C.c_long
    L.emu_split_clusters.argtypes = [C.c_int, _u64p, _u32p, _u8p, f32p, _u64p, _u32p, C.c_int, C.c_int, _u8p, _i32p, _u64p, _u32p, _i32p, _i32p, C.c_uint64]
    co = np.ascontiguousarray(cl_off, np.uint64); R, Cn = len(co) - 1, int(co[-1])
    pad = lambda a, dt: np.ascontiguousarray(a, dt).reshape(-1) if len(a) else np.zeros(1, dt)
    o = dict(split=np.zeros(max(Cn, 1), np.uint8), val_cluster=np.zeros(max(Cn, 1), np.int32), sp_off=np.zeros(R + 2, np.uint64), sp=np.zeros(6 * cap, np.uint32),
             sp_val=np.zeros(cap, np.int32), sp_n0=np.zeros(cap, np.int32))
    n = L.emu_split_clusters(R, co, pad(box, np.uint32), pad(strand, np.uint8), pad(freq, np.float32), np.ascontiguousarray(m_off, np.uint64), pad(mq, np.uint32), contig, global_k,
                             o["split"], o["val_cluster"], o["sp_off"], o["sp"], o["sp_val"], o["sp_n0"], cap)
    assert n <= cap
    return dict(split=o["split"][:Cn], val_cluster=o["val_cluster"][:Cn], sp_off=o["sp_off"][:R + 1], sp=o["sp"][:6 * n].reshape(-1, 6), sp_val=o["sp_val"][:n], sp_n0=o["sp_n0"][:n])


def mapq(ag, logv, lenpen, bypass, read_type):
    L = lib()
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    L.emu_mapq.argtypes = [C.c_int, _i32p, _i32p, _i32p, _i32p, f32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _u8p, f32p, _i32p, C.c_int, C.c_int, _i32p, _i32p, _u8p, _u8p,
                           _i32p, _u8p, f32p, _i32p, _i32p, _i32p, _i32p]
    i32 = lambda k: np.ascontiguousarray(ag[k], np.int32)
    G = int(ag["grp_off"][-1]); S = int(ag["seg_off"][G]); R = len(ag["grp_off"]) - 1
    o = dict(flag=np.array(ag["flag"], np.int32), typeofaln=np.array(ag["typeofaln"], np.int32), issec=np.array(ag["issec"], np.uint8), supp=np.array(ag["supp"], np.uint8),
             mapq=np.zeros(S, np.int32), g_issec=np.zeros(G, np.uint8), g_value=np.zeros(G, np.float32), g_n0=np.zeros(G, np.int32), g_n1=np.zeros(G, np.int32),
             g_nm=np.zeros(4 * G, np.int32), order=np.zeros(G, np.int32))
    L.emu_mapq(R, i32("grp_off"), i32("seg_off"), i32("upd_off"), i32("update_at"), np.ascontiguousarray(ag["value"], np.float32), i32("n0"), i32("n1"), i32("nm"), i32("nmm"),
               i32("ndel"), i32("nins"), np.ascontiguousarray(ag["strand"], np.uint8), np.ascontiguousarray(logv, np.float32), np.ascontiguousarray(lenpen, np.int32), bypass,
               read_type, o["flag"], o["typeofaln"], o["issec"], o["supp"], o["mapq"], o["g_issec"], o["g_value"], o["g_n0"], o["g_n1"], o["g_nm"], o["order"])
    o["g_nm"] = o["g_nm"].reshape(-1, 4)
    return o


def linear_extend(read_arena, genome, ep, K, skipsorting, trim):
    """read_arena / genome: ASCII arenas with 16 bytes of padding; ep as for Context.linear_extend_batch."""
    L = lib()
    L.emu_linear_extend.argtypes = [_u8p, C.c_uint64, _u8p, C.c_uint64, C.c_int, _u64p, _u64p, _u8p, _u64p, _u32p, _u64p, _u32p, _u32p, _u32p, C.c_int, C.c_int, C.c_int,
                                    _u64p, _u32p, _u32p, _i32p, _u32p]
    G = len(ep["g_off"]) - 1; N = len(ep["q"])
    pad = lambda a, dt: np.ascontiguousarray(a, dt).copy() if len(a) else np.zeros(1, dt)
    o = dict(e_off=np.zeros(G + 1, np.uint64), q=np.zeros(max(N, 1), np.uint32), t=np.zeros(max(N, 1), np.uint32), len=np.zeros(max(N, 1), np.int32),
             box=np.zeros(4 * max(G, 1), np.uint32))
    L.emu_linear_extend(read_arena, len(read_arena) - 16, genome, len(genome) - 16, G, np.ascontiguousarray(ep["g_off"], np.uint64), pad(ep["p_off"], np.uint64),
                        pad(ep["p_strand"], np.uint8), pad(ep["chrom_off"], np.uint64), pad(ep["chrom_len"], np.uint32), pad(ep["read_off"], np.uint64),
                        pad(ep["read_len"], np.uint32), pad(ep["q"], np.uint32), pad(ep["t"], np.uint32), K, int(skipsorting), int(trim), o["e_off"], o["q"], o["t"],
                        o["len"], o["box"])
    n = int(o["e_off"][G])
    for k in ("q", "t", "len"):
        o[k] = o[k][:n]
    o["box"] = o["box"][:4 * G].reshape(-1, 4)
    return o


def linear_extend_chains(read_arena, genome, cd, K, skiprepetitive=1, trim=1, merge_dist=100):
    """cd as for Context.linear_extend_chains_batch."""
    L = lib()
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    L.emu_linear_extend_chains.argtypes = [_u8p, C.c_uint64, _u8p, C.c_uint64, C.c_int, _u64p, _u32p, C.c_int, _u64p, _u32p, _u32p, _u32p, _u8p, f32p, _u64p, _u32p, _u64p,
                                           _u32p, C.c_int, C.c_int, C.c_int, C.c_int, _u64p, _u32p, _u32p, _i32p, _u8p, _u8p, _u32p, _i32p]
    pad = lambda a, dt: np.ascontiguousarray(a, dt).reshape(-1).copy() if len(a) else np.zeros(1, dt)
    cho = np.ascontiguousarray(cd["ch_off"], np.uint64); clo = np.ascontiguousarray(cd["cl_off"], np.uint64)
    ch = pad(cd["ch"], np.uint32); U = int(cho[-1])
    cap = int(np.diff(clo.astype(np.int64))[ch[:U]].sum()) if U else 0
    o = dict(e_off=np.zeros(U + 1, np.uint64), q=np.zeros(max(cap, 1), np.uint32), t=np.zeros(max(cap, 1), np.uint32), len=np.zeros(max(cap, 1), np.int32),
             ovp=np.zeros(max(cap, 1), np.uint8), md_head=np.zeros(max(cap, 1), np.uint8), box=np.zeros(4 * max(U, 1), np.uint32), overlap=np.zeros(max(U, 1), np.int32))
    L.emu_linear_extend_chains(read_arena, len(read_arena) - 16, genome, len(genome) - 16, len(cho) - 1, cho, ch, len(clo) - 1, clo, pad(cd["q"], np.uint32),
                               pad(cd["t"], np.uint32), pad(cd["box"], np.uint32), pad(cd["strand"], np.uint8), pad(cd["freq"], np.float32), pad(cd["chrom_off"], np.uint64),
                               pad(cd["chrom_len"], np.uint32), pad(cd["read_off"], np.uint64), pad(cd["read_len"], np.uint32), K, int(skiprepetitive), int(trim),
                               int(merge_dist), o["e_off"], o["q"], o["t"], o["len"], o["ovp"], o["md_head"], o["box"], o["overlap"])
    n = int(o["e_off"][U])
    for k in ("q", "t", "len", "ovp", "md_head"):
        o[k] = o[k][:n]
    o["box"] = o["box"][:4 * U].reshape(-1, 4); o["overlap"] = o["overlap"][:U]
    return o


def split_chains(ac, hdr_pos, splitdist=50000, bypass=0):
    """ac as for Context.split_chains_batch; returns the same slot-layout dict."""
    L = lib()
    L.emu_split_chains.argtypes = [C.c_int, _u64p, _u32p, _u32p, _i32p, _u8p, _i32p, _u8p, _u64p, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _u8p,
                                   _u32p, _i32p, _u8p, _u8p, _u8p]
    co = np.ascontiguousarray(ac["c_off"], np.uint64); NC = len(co) - 1; N = int(co[-1]); Np = max(N, 1)
    pad = lambda a, dt: np.ascontiguousarray(a, dt).copy() if len(a) else np.zeros(1, dt)
    o = dict(n_sp=np.zeros(max(NC, 1), np.int32), n_link=np.zeros(max(NC, 1), np.int32), sp_off=np.zeros(Np + NC + 1, np.int32), ci_off=np.zeros(Np + NC + 1, np.int32),
             sptc=np.zeros(Np, np.int32), ci=np.zeros(Np, np.int32), sp_lk=np.zeros(Np, np.uint8), sp_box=np.zeros(4 * Np, np.uint32), sp_chrom=np.zeros(Np, np.int32),
             sp_type=np.zeros(Np, np.uint8), sp_strand=np.zeros(Np, np.uint8), sp_link=np.zeros(Np, np.uint8))
    hdr = np.ascontiguousarray(hdr_pos, np.uint64)
    L.emu_split_chains(NC, co, pad(ac["q"], np.uint32), pad(ac["t"], np.uint32), pad(ac["len"], np.int32), pad(ac["strand"], np.uint8), pad(ac["cnum"], np.int32),
                       pad(ac["link"], np.uint8), hdr, len(hdr), int(splitdist), int(bypass), o["n_sp"], o["n_link"], o["sp_off"], o["ci_off"], o["sptc"], o["ci"], o["sp_lk"],
                       o["sp_box"], o["sp_chrom"], o["sp_type"], o["sp_strand"], o["sp_link"])
    o["sp_box"] = o["sp_box"].reshape(-1, 4)
    return o


def merge_chain(sp, first, chrom, strand, box):
    L = lib()
    L.emu_merge_chain.argtypes = [C.c_uint64, _i32p, _u8p, _i32p, _u8p, _u32p, _u8p]
    n = len(sp)
    head = np.zeros(max(n, 1), np.uint8)
    pad = lambda a, dt: np.ascontiguousarray(a, dt).reshape(-1) if len(a) else np.zeros(1, dt)
    L.emu_merge_chain(n, pad(sp, np.int32), pad(first, np.uint8), pad(chrom, np.int32), pad(strand, np.uint8), pad(box, np.uint32), head)
    return head[:n]


def switchindex(ch, link, c_off, coarse, cq):
    L = lib()
    L.emu_switchindex.argtypes = [C.c_int, _u64p, _i32p, _u8p, _i32p, _u32p, _i32p, _i32p]
    co = np.ascontiguousarray(c_off, np.uint64); NC = len(co) - 1
    ch = np.array(ch, np.int32) if len(ch) else np.zeros(1, np.int32); link = np.array(link, np.uint8) if len(link) else np.zeros(1, np.uint8)
    n_out = np.zeros(max(NC, 1), np.int32); nl_out = np.zeros(max(NC, 1), np.int32)
    L.emu_switchindex(NC, co, ch, link, np.ascontiguousarray(coarse, np.int32), np.ascontiguousarray(cq, np.uint32).reshape(-1), n_out, nl_out)
    return ch, link, n_out[:NC], nl_out[:NC]


def refine_linear(read_arena, genome, g, m, mm, indel, local_band):
    """g as for Context.refine_linear_batch; arenas with 16 bytes of padding."""
    L = lib()
    L.emu_refine_linear.argtypes = [_u8p, C.c_uint64, _u8p, C.c_uint64, C.c_int, _u32p, _u32p, _u32p, _u32p, _u32p, _u32p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p,
                                    _u64p, _u32p, C.c_uint64]
    a = {k: np.ascontiguousarray(g[k], np.uint32) for k in ["cur_read_end", "next_read_start", "cur_genome_end", "next_genome_start", "read_off", "chrom_off"]}
    n = len(a["read_off"])
    ql = (a["next_read_start"] - a["cur_read_end"]).astype(np.int32).astype(np.int64); tl = (a["next_genome_start"] - a["cur_genome_end"]).astype(np.int32).astype(np.int64)
    cap = int((np.minimum(ql, tl).clip(min=0) + 1).sum()) + 1
    o = dict(score=np.zeros(n, np.int32), n_blocks=np.zeros(n, np.int32), block_off=np.zeros(n, np.uint64), blocks=np.zeros(3 * cap, np.uint32))
    err = L.emu_refine_linear(read_arena, len(read_arena) - 16, genome, len(genome) - 16, n, a["cur_read_end"], a["next_read_start"], a["cur_genome_end"], a["next_genome_start"],
                              a["read_off"], a["chrom_off"], m, mm, indel, local_band, o["score"], o["n_blocks"], o["block_off"], o["blocks"], cap)
    assert err == 0
    o["blocks"] = o["blocks"].reshape(-1, 3)
    return o


def refine_space(read_arena, genome, sp, K, m, mm, indel):
    """sp as for Context.refine_space_batch."""
    L = lib()
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    L.emu_refine_space.argtypes = [_u8p, C.c_uint64, _u8p, C.c_uint64, C.c_int, C.c_int] + [_u32p] * 9 + [_u8p, C.c_int, C.c_int, C.c_int, _u64p, _u32p, _u32p, _i32p, f32p, C.c_uint64]
    a = {k: np.ascontiguousarray(sp[k], np.uint32) for k in ["qs", "qe", "ts", "te", "lrts", "lrlength", "read_off", "read_len", "chrom_off"]}
    flip = np.ascontiguousarray(sp["flip"], np.uint8)
    n = len(flip)
    ql = a["qe"].astype(np.int64) - a["qs"]; tl = a["te"].astype(np.int64) - a["ts"] + a["lrlength"]
    mn = np.minimum(ql, tl).clip(min=0)
    pair_off = np.zeros(n + 1, np.uint64); pair_off[1:] = np.cumsum(mn // K + 1)
    cap = int(pair_off[-1]) + 1
    o = dict(pair_off=pair_off, n_pairs=np.zeros(n, np.int32), identity=np.zeros(n, np.float32), pq=np.zeros(cap, np.uint32), pt=np.zeros(cap, np.uint32))
    err = L.emu_refine_space(read_arena, len(read_arena) - 16, genome, len(genome) - 16, n, K, *[a[k] for k in ["qs", "qe", "ts", "te", "lrts", "lrlength", "read_off", "read_len",
                             "chrom_off"]], flip, m, mm, indel, pair_off, o["pq"], o["pt"], o["n_pairs"], o["identity"], int((mn + 1).sum()) + 1)
    assert err == 0
    return o


def split_rough(rl, globalK, max_gap, min_cluster_size, max_diag):
    """rl as for Context.split_rough_batch; returns the same slot-layout dict."""
    L = lib()
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    L.emu_split_rough.argtypes = [C.c_int, _u64p, _u64p, _u32p, _u32p, _i32p, _i32p, _u32p, _u8p, f32p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p,
                                  _i32p, _i32p, _u32p, _u8p, f32p, _i32p, _i32p, _i32p]
    lo = np.ascontiguousarray(rl["l_off"], np.uint64); lro = np.ascontiguousarray(rl["lr_off"], np.uint64)
    NL = len(lo) - 1; T = int(lo[-1]) + int(lro[-1]) + 1
    pad = lambda a, dt: np.ascontiguousarray(a, dt).reshape(-1) if len(a) else np.zeros(1, dt)
    o = dict(n_split=np.zeros(max(NL, 1), np.int32), n_piece=np.zeros(max(NL, 1), np.int32), s_start=np.zeros(T, np.int32), s_end=np.zeros(T, np.int32),
             s_coarse=np.zeros(T, np.int32), s_chrom=np.zeros(T, np.int32), s_box=np.zeros(4 * T, np.uint32), s_strand=np.zeros(T, np.uint8), s_freq=np.zeros(T, np.float32),
             p_cluster=np.zeros(T, np.int32), p_start=np.zeros(T, np.int32), p_end=np.zeros(T, np.int32))
    L.emu_split_rough(NL, lo, lro, pad(rl["q"], np.uint32), pad(rl["t"], np.uint32), pad(rl["r_start"], np.int32), pad(rl["r_end"], np.int32), pad(rl["r_box"], np.uint32),
                      pad(rl["r_strand"], np.uint8), pad(rl["r_freq"], np.float32), pad(rl["r_chrom"], np.int32), globalK, max_gap, min_cluster_size, max_diag,
                      o["n_split"], o["n_piece"], o["s_start"], o["s_end"], o["s_coarse"], o["s_chrom"], o["s_box"], o["s_strand"], o["s_freq"], o["p_cluster"], o["p_start"], o["p_end"])
    o["s_box"] = o["s_box"].reshape(-1, 4)
    return o


def refine_space_large(read_arena, genome, sp, K, W, max_freq):
    """The minimizer branch of RefineSpace for every space of sp (as Context.refine_space_batch, with diag)."""
    L = lib()
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    L.emu_refine_space_large.restype = C.c_long
    L.emu_refine_space_large.argtypes = [_u8p, C.c_uint64, _u8p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_longlong] + [_u32p] * 9 + [_u8p, _i32p, _u64p, _u32p, _u32p, C.c_uint64,
                                         _i32p, f32p]
    a = {k: np.ascontiguousarray(sp[k], np.uint32) for k in ["qs", "qe", "ts", "te", "lrts", "lrlength", "read_off", "read_len", "chrom_off"]}
    flip = np.ascontiguousarray(sp["flip"], np.uint8); diag = np.ascontiguousarray(sp["diag"], np.int32)
    n = len(flip)
    cap = 1 << 16
    while True:
        o = dict(pair_off=np.zeros(n + 1, np.uint64), n_pairs=np.zeros(n, np.int32), identity=np.zeros(n, np.float32), pq=np.zeros(cap, np.uint32), pt=np.zeros(cap, np.uint32))
        r = L.emu_refine_space_large(read_arena, len(read_arena) - 16, genome, len(genome) - 16, n, K, W, max_freq, *[a[k] for k in ["qs", "qe", "ts", "te", "lrts", "lrlength",
                                     "read_off", "read_len", "chrom_off"]], flip, diag, o["pair_off"], o["pq"], o["pt"], cap, o["n_pairs"], o["identity"])
        if r >= 0:
            return o
        cap = -r + 16


def store_diagonal(cl, hdr_pos, globalK, max_diag, min_cluster_size, min_cluster_length, bypass):
    L = lib()
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    L.emu_store_diagonal.argtypes = [C.c_int, _u64p, _u32p, _u32p, _u64p, f32p, _u8p, _u64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p,
                                     _u32p, f32p]
    lo = np.ascontiguousarray(cl["l_off"], np.uint64); NL = len(lo) - 1; N = max(int(lo[-1]), 1)
    pad = lambda a, dt: np.ascontiguousarray(a, dt) if len(a) else np.zeros(1, dt)
    hdr = np.ascontiguousarray(hdr_pos, np.uint64)
    o = dict(n_cl=np.zeros(max(NL, 1), np.int32), c_start=np.zeros(N, np.int32), c_end=np.zeros(N, np.int32), c_chrom=np.zeros(N, np.int32), c_box=np.zeros(4 * N, np.uint32),
             c_freq=np.zeros(N, np.float32))
    L.emu_store_diagonal(NL, lo, pad(cl["q"], np.uint32), pad(cl["t"], np.uint32), pad(cl["qt"], np.uint64), pad(cl["freq"], np.float32), pad(cl["strand"], np.uint8), hdr, len(hdr),
                         globalK, max_diag, min_cluster_size, min_cluster_length, int(bypass), o["n_cl"], o["c_start"], o["c_end"], o["c_chrom"], o["c_box"], o["c_freq"])
    o["c_box"] = o["c_box"].reshape(-1, 4)
    return o


def trim_splitchains(cq, ct, c_off, strand, q, t, m_off):
    L = lib()
    L.emu_trim_splitchains.argtypes = [C.c_int, _u64p, _u32p, _u32p, _u8p, _u64p, _u32p, _u32p, _u8p, _i32p]
    co = np.ascontiguousarray(c_off, np.uint64); mo = np.ascontiguousarray(m_off, np.uint64); n = len(co) - 1
    pad = lambda a, dt: np.array(a, dt) if len(a) else np.zeros(1, dt)
    q2 = pad(q, np.uint32); t2 = pad(t, np.uint32)
    keep = np.zeros(max(len(q), 1), np.uint8); removed = np.zeros(max(n, 1), np.int32)
    L.emu_trim_splitchains(n, co, pad(cq, np.uint32), pad(ct, np.uint32), pad(strand, np.uint8), mo, q2, t2, keep, removed)
    return q2[:len(q)], t2[:len(t)], keep[:len(q)], removed[:n]
