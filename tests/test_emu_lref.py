"""Kernel logic of a12 (lidx_kernels.cuh) and a13 (lref_kernels.cuh) through the SIMT emulator against the oracle."""
import numpy as np
import pytest

from oracle import pyoracle as po
import emu_lib
import refinegen

B = np.frombuffer(b"ACGT", np.uint8)


def _img_equals(img, li):
    nw = img["n_win"]
    return (len(li.seq_off) == nw + 1 and (img["bnd"][:nw + 1] == li.bnd).all() and img["n_mins"] == len(li.mins)
            and (img["mins"][:img["n_mins"]] == li.mins).all())


def test_emu_lindex_contigs():
    rng = np.random.default_rng(5)
    contigs = []
    for L in (13, 14, 300, 2048, 5000, 4099):
        s = B[rng.integers(0, 4, L)].copy()
        if L == 5000:
            s[1000:1012] = ord("N"); s[3000:4500] = np.tile(B[rng.integers(0, 4, 2)], 750); s[-2:] = ord("N")
        if L == 4099:
            s[:2100] = ord("A")
        contigs.append(s)
    # N handling on the lane-parallel path (no tied minima): sporadic N's, a valid run that starts exactly at seqLen - windowSpan
    # (never used by the reference), one that starts one base earlier (used), N as the very last base, a sequence of windowSpan bases
    for L, ns in ((3000, None), (500, [485]), (500, [484]), (500, [499]), (14, []), (15, []), (15, [0]), (2048 + 14, [2047]), (2048 + 15, [2048])):
        s = B[rng.integers(0, 4, L)].copy()
        if ns is None:
            s[rng.random(L) < 0.04] = ord("N")
        else:
            s[ns] = ord("N")
        contigs.append(s)
    lens = np.array([len(c) for c in contigs], np.uint32)
    start = np.zeros(len(contigs), np.uint64); start[1:] = np.cumsum(lens[:-1])
    arena = np.concatenate(contigs + [np.full(16, ord("N"), np.uint8)])
    for mf in (5, 15):
        img = emu_lib.lindex_build(arena, start, lens, max_freq=mf)
        li = po.local_index(contigs, max_freq=mf, which="port")
        assert (img["win_off"] == li.seq_off).all()
        assert _img_equals(img, li)


def test_emu_refine_clusters():
    case = refinegen.make_case(21, n_reads=5, contig_lens=(30000, 9000, 12000), read_lens=(700, 2500, 5000))
    pk = refinegen.pack_case(case)
    hdr = case["hdr"]
    gl = emu_lib.lindex_build(pk["genome"], hdr[:-1], np.diff(hdr).astype(np.uint32))
    rf = emu_lib.lindex_build(pk["arena"], pk["read_off"], pk["read_len"])
    rr = emu_lib.lindex_build(pk["rc_arena"], pk["read_off"], pk["read_len"])
    # the read images equal the per-read LocalIndex of the oracle
    for i, r in enumerate(case["reads"]):
        li = po.local_index(r, which="port")
        w0, w1 = int(rf["win_first"][i]), int(rf["win_first"][i + 1])
        a, b = int(rf["bnd"][w0]), int(rf["bnd"][w1])
        assert w1 - w0 == len(li.seq_off) - 1 and (rf["mins"][a:b] == li.mins).all()
    exp = refinegen.expected(case, "port")
    for literal in (0, 1):       # the warp kernel and the statement-by-statement thread kernel
        o = emu_lib.refine_clusters(gl, rf, rr, pk["cl"], literal=literal)
        assert o["n_anchors"] == sum(len(e["rq"]) for e in exp) and o["n_anchors"] > 200
        refinegen.check_batch(o, pk["cl"], exp)


def test_emu_refine_splitchains():
    case = refinegen.make_case(22, n_reads=5, contig_lens=(30000, 9000, 12000), read_lens=(700, 2500, 5000))
    chains = refinegen.make_chains(case, 22)
    hdr = case["hdr"]
    for limit in (1, 0):
        pk = refinegen.pack_chains(case, chains, limitrefine=limit)
        gl = emu_lib.lindex_build(pk["genome"], hdr[:-1], np.diff(hdr).astype(np.uint32))
        rf = emu_lib.lindex_build(pk["arena"], pk["read_off"], pk["read_len"])
        rr = emu_lib.lindex_build(pk["rc_arena"], pk["read_off"], pk["read_len"])
        o = emu_lib.refine_clusters(gl, rf, rr, pk["cl"], literal=1)
        exp = refinegen.expected_chains(case, chains, "port", limitrefine=limit)
        assert o["n_anchors"] == sum(len(e["rq"]) for e in exp) and o["n_anchors"] > 200
        refinegen.check_chain_batch(o, exp)
