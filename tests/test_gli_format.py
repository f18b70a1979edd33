"""f2 (local half): the LocalIndex built by this repository, written with lra_b200.write_gli, is byte-identical to the <ref>.gli that the
reference binary (`lra index`, oracle/_ref/lra_ref) writes for the same FASTA."""
import os
import subprocess
import numpy as np
import pytest

from oracle import pyoracle as po
import lra_b200

LRA_REF = os.path.join(os.path.dirname(os.path.abspath(po.__file__)), "_ref", "lra_ref")
needs_bin = pytest.mark.skipif(not os.path.exists(LRA_REF), reason="oracle/_ref/lra_ref not built (no /root/reference at build time)")
B = np.frombuffer(b"ACGT", np.uint8)


def genome(seed):
    rng = np.random.default_rng(seed)
    contigs = []
    for L in (70001, 4096, 123457, 2047):
        s = B[rng.integers(0, 4, L)].copy()
        if L == 123457:
            s[5000:5100] = ord("N"); s[40000:52000] = np.tile(B[rng.integers(0, 4, 3)], 4000); s[90000:96000] = ord("A"); s[-1] = ord("N")
            s[60000:60500] = np.char.lower(s[60000:60500].view("S1")).view(np.uint8)          # soft-masked stretch
        contigs.append(s)
    return contigs


def reference_gli(tmp_path, contigs):
    fa = tmp_path / "g.fa"
    with open(fa, "w") as f:
        for i, c in enumerate(contigs):
            f.write(">chr%d some description\n" % (i + 1))
            s = c.tobytes().decode()
            for a in range(0, len(s), 80):
                f.write(s[a:a + 80] + "\n")
    subprocess.run([LRA_REF, "index", "-ONT", str(fa)], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=str(tmp_path))
    return open(str(fa) + ".gli", "rb").read()


@needs_bin
def test_oracle_gli_is_byte_identical(tmp_path):
    contigs = genome(3)
    want = reference_gli(tmp_path, contigs)
    li = po.local_index(contigs, k=10, w=5, window=2048, max_freq=15, which="port")
    out = tmp_path / "port.gli"
    lra_b200.write_gli(str(out), 10, 5, 2048, li.seq_off, li.bnd, li.mins)
    assert open(out, "rb").read() == want
    back = lra_b200.read_gli(str(out))
    assert back["k"] == 10 and (back["minimizers"] == li.mins).all() and (back["seq_offsets"] == li.seq_off).all()


@needs_bin
@pytest.mark.gpu
def test_gpu_gli_is_byte_identical(tmp_path):
    contigs = genome(4)
    want = reference_gli(tmp_path, contigs)
    ctx = lra_b200.Context(0)
    lens = np.array([len(c) for c in contigs], np.uint32)
    start = np.zeros(len(contigs), np.uint64); start[1:] = np.cumsum(lens[:-1])
    arena = ctx.seq_upload(np.concatenate(contigs))
    img = ctx.lindex_build(arena, start, lens, k=10, w=5, window=2048, max_freq=15)
    wo, bd, mn = img.download()
    out = tmp_path / "gpu.gli"
    lra_b200.write_gli(str(out), 10, 5, 2048, wo, bd, mn)
    assert open(out, "rb").read() == want
    img.free(); arena.free(); ctx.close()
