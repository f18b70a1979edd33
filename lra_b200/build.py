"""Builds lra_b200/liblra_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "liblra_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared"]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "lra_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


CLI_SRC = os.path.join(HERE, "cli", "lra_b200_cli.cpp")
CLI_OUT = os.path.join(HERE, "lra_b200")


def build_cli(force=False):
    """The C++ host program above the C ABI (`lra_b200 index|align`, lra_b200/cli/): plain g++, linked against liblra_b200.so."""
    if not force and os.path.exists(CLI_OUT) and os.path.getmtime(CLI_OUT) > max(os.path.getmtime(CLI_SRC), os.path.getmtime(OUT)):
        return CLI_OUT
    subprocess.run(["g++", "-std=c++17", "-O2", CLI_SRC, "-o", CLI_OUT, "-L" + HERE, "-llra_b200", "-lz", "-pthread", "-Wl,-rpath,$ORIGIN"], check=True)
    return CLI_OUT


def build(force=False, verbose=False):
    if force or needs_build():
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + sources() + ["-o", OUT]
        subprocess.run(cmd, check=True)
    build_cli(force)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
