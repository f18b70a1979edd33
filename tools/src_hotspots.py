#!/usr/bin/env python
"""Attribute the per-SASS-instruction columns of an ncu source page (reduced CSV of tools/map_profile.sh) to CUDA source lines / functions, using
the line table of the cubin (nvdisasm -g).  Usage: src_hotspots.py <reduced src csv> <kernel mangled name> [column]"""
import csv, re, subprocess, sys, os, collections, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def line_table(kernel):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "lra_b200", "liblra_b200.so")], cwd=d, stdout=subprocess.DEVNULL, check=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], stdout=subprocess.PIPE, text=True).stdout
    tab = {}; on = False; cur = None
    for l in txt.split("\n"):
        if l.startswith("\t.section\t.text."):
            on = kernel in l
            continue
        if not on: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/", l)
        if m: tab[int(m.group(1), 16)] = cur
    return tab

def functions_of(path):
    """(start line, name) of every function definition in a .cuh (rough)."""
    out = []
    for i, l in enumerate(open(path), 1):
        m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__|static|inline|extern)[^;(]*?([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
        if m and not l.strip().endswith(";"): out.append((i, m.group(1)))
    return out

def main():
    src, kernel = sys.argv[1], sys.argv[2]
    col = sys.argv[3] if len(sys.argv) > 3 else "Warp Stall Sampling (All Samples)"
    tab = line_table(kernel)
    rows = list(csv.reader(open(src)))
    on = False; ci = None; byline = collections.Counter(); tot = 0; base = None
    for r in rows:
        if r and r[0] == "Kernel Name": on = kernel_short(kernel) in r[1]; base = None; continue
        if not on: continue
        if r and r[0] == "Address": ci = r.index(col) if col in r else None; continue
        if ci is None: continue
        try: a = int(r[0], 16); v = float(r[ci] or 0)
        except Exception: continue
        if base is None: base = a
        a -= base
        byline[tab.get(a)] += v; tot += v
    fn_cache = {}
    byfn = collections.Counter()
    for (k, v) in byline.items():
        if k is None: byfn[("?", "?")] += v; continue
        f, ln = k
        p = os.path.join(ROOT, "lra_b200", "csrc", f)
        if f not in fn_cache: fn_cache[f] = functions_of(p) if os.path.exists(p) else []
        name = "?"
        for s, n in fn_cache[f]:
            if s <= ln: name = n
            else: break
        byfn[(f, name)] += v
    print("total %s = %.0f" % (col, tot))
    print("--- by function (of the source line: inlined code counts where it was written)")
    for (f, n), v in byfn.most_common(40): print("%6.2f %%  %-22s %s" % (100 * v / max(tot, 1), f, n))
    print("--- top lines")
    for k, v in byline.most_common(40): print("%6.2f %%  %s" % (100 * v / max(tot, 1), k))

def kernel_short(k):
    return "map_reads_kernel" if "map_reads" in k else k

main()
