"""Builds and binds tests/simt/mp/libemu_mp{32,1}.so: the mapper worker code (lra_b200/csrc/mp_*.cuh) compiled for the CPU
through the SIMT emulator (TEST INFRASTRUCTURE).  lanes=32 is the product's lane count; lanes=1 runs the same source with
identity collectives at CPU speed."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SIMT = os.path.join(HERE, "simt")
MPD = os.path.join(SIMT, "mp")
CSRC = os.path.join(os.path.dirname(HERE), "lra_b200", "csrc")
_libs = {}
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def lib(lanes=32):
    if lanes in _libs:
        return _libs[lanes]
    so = os.path.join(MPD, "libemu_mp%d.so" % lanes)
    srcs = [os.path.join(MPD, f) for f in os.listdir(MPD) if f.endswith((".cpp", ".h"))]
    srcs += [os.path.join(SIMT, "cuda_emu.h")]
    srcs += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        cpps = [s for s in srcs if s.endswith(".cpp")]
        if lanes == 1:      # the 1-lane build binds the in-worker AffineOneGapAlign to the pinned C restatement (see mp_aog.cuh)
            obj = os.path.join(MPD, "aog_port.o")
            subprocess.run(["gcc", "-std=c11", "-O2", "-fPIC", "-c", os.path.join(os.path.dirname(HERE), "oracle", "aog.c"), "-o", obj], check=True)
            cpps.append(obj)
        subprocess.run(["g++", "-std=c++17", "-O2" if lanes == 1 else "-O1", "-DLRA_EMU", "-DMP_LANES=%d" % lanes] + (["-DMP_DEBUG", "-g"] if os.environ.get("MP_DEBUG") else []) + [ "-I" + SIMT, "-I" + CSRC, "-fPIC", "-shared"] + cpps +
                       ["-o", so], check=True)
    L = C.CDLL(so)
    L.emu_sdp_batch.restype = C.c_int
    L.emu_sdp_batch.argtypes = [C.c_int, C.c_int, _i32p, _u64p, _u32p, _u32p, _i32p, _u64p, _i32p, _u8p, _i32p, _f32p, _i32p, _i32p, C.c_float, C.c_int,
                                _i64p, _f32p, _f32p, C.c_int, C.c_int, _i32p, _i32p, _f32p, _u32p, _u32p, _u8p, _i32p, C.c_uint64, _u64p]
    L.emu_sdp_batch_ext.restype = C.c_int
    L.emu_sdp_batch_ext.argtypes = L.emu_sdp_batch.argtypes + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, _i32p]
    _libs[lanes] = L
    return L


def sdp_batch(pb, pwl, alnthres, NumAln, max_aln=2, lanes=32, arena_bytes=256 << 20):
    """pb: packed problems from sdpgen.pack(); pwl: (stops, slope, inter, ceil1, ceil2)."""
    L = lib(lanes)
    n = len(pb["mode"])
    nf = int(pb["frag_off"][-1])
    out = dict(n_chains=np.zeros(n, np.int32), chain_len=np.zeros(n * max_aln, np.int32), chain_val=np.zeros(n * max_aln, np.float32),
               bounds=np.zeros(4 * n * max_aln, np.uint32), chain=np.zeros(max(1, nf * max_aln), np.uint32), link=np.zeros(max(1, nf * max_aln), np.uint8),
               cl_of_frag=np.zeros(max(1, nf), np.int32))
    peak = np.zeros(1, np.uint64)
    if "qe" in pb:
        out["n0"] = np.zeros(n * max_aln, np.int32)
        err = L.emu_sdp_batch_ext(n, max_aln, pb["mode"], pb["frag_off"], pb["q"], pb["t"], pb["len"], pb["cl_off_off"], pb["cl_off"], pb["cl_strand"], pb["only_cl"],
                                  pb["rate"], pb["irate"], pb["read_len"], alnthres, NumAln, pwl[0], pwl[1], pwl[2], pwl[3], pwl[4],
                                  out["n_chains"], out["chain_len"], out["chain_val"], out["bounds"], out["chain"], out["link"], out["cl_of_frag"], arena_bytes, peak,
                                  pb["qe"].ctypes.data, pb["te"].ctypes.data, pb["fstrand"].ctypes.data, pb["fval"].ctypes.data, pb["fn0"].ctypes.data, pb["globalK"], out["n0"])
        out["err"] = err; out["peak"] = int(peak[0])
        return out
    err = L.emu_sdp_batch(n, max_aln, pb["mode"], pb["frag_off"], pb["q"], pb["t"], pb["len"], pb["cl_off_off"], pb["cl_off"], pb["cl_strand"], pb["only_cl"],
                          pb["rate"], pb["irate"], pb["read_len"], alnthres, NumAln, pwl[0], pwl[1], pwl[2], pwl[3], pwl[4],
                          out["n_chains"], out["chain_len"], out["chain_val"], out["bounds"], out["chain"], out["link"], out["cl_of_frag"], arena_bytes, peak)
    out["err"] = err
    out["peak"] = int(peak[0])
    return out
