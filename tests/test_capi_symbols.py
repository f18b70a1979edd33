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
