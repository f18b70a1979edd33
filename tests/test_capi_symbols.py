"""CPU check: liblra_b200.so builds for sm_100a, loads, exports every symbol include/lra_b200.h declares, and fails
loudly (no CPU fallback) when no CUDA device exists."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from lra_b200 import build, capi
    build.build()
    return capi.load_library()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "lra_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(lra_b200_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), s


def test_version(lib):
    assert lib.lra_b200_version() >= 100


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lra_b200 import capi
    with pytest.raises(capi.LraB200Error) as e:
        capi.Context(0)
    assert e.value.code == capi.ECUDA
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """The shipped package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "lra_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "pyoracle" not in src and "liboracle" not in src and "libref_lra" not in src, f


def test_every_batch_entry_point_has_a_python_binding():
    """Every lra_b200_*_batch[_device] the header declares is called by the ctypes layer (a binding dropped by an edit shows up here, without a GPU)."""
    src = open(os.path.join(ROOT, "lra_b200", "capi.py")).read()
    for s in declared_symbols():
        if s.endswith("_batch") or s.endswith("_batch_device"):
            assert re.search(r"\.lib\.%s\(" % s, src), s


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/lra_b200.h compiles as C11 (no C++ in the boundary) and a C program that takes the address of every declared entry point links against
    liblra_b200.so and fails loudly, without a GPU, at lra_b200_create (the reference-side binding of INTEGRATION.md is C++, a cgo / FFI binding would be C)."""
    import subprocess
    from lra_b200 import capi
    names = declared_symbols()
    src = tmp_path / "abi.c"
    body = "\n".join("  p[%d] = (void *)%s;" % (i, n) for i, n in enumerate(names))
    src.write_text('#include <stdio.h>\n#include "lra_b200.h"\nint main(void) {\n  void *p[%d];\n%s\n  lra_b200_ctx *ctx = NULL;\n  int rc = lra_b200_create(&ctx, 0);\n'
                   '  printf("%%d %%p\\n", rc, p[0]);\n  return rc == LRA_B200_OK ? 0 : 3;\n}\n' % (len(names) + 1, body))
    exe = tmp_path / "abi"
    so = capi.library_path()
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), so, "-Wl,-rpath," + os.path.dirname(so)], check=True)
    import torch
    p = subprocess.run([str(exe)], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert p.returncode == 0
    else:
        assert p.returncode == 3, (p.returncode, p.stdout, p.stderr)      # LRA_B200_ECUDA: no CPU fallback
