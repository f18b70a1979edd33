#!/usr/bin/env python
"""Summarise ncu outputs brought back from the GPU box into profiles/ (markdown + traffic.json).

  python tools/ncu_summary.py --launches gpurun_out/launches.csv --raw gpurun_out/prof_r1_raw.csv --tag r01
`--raw` is the output of  ncu -i <rep> --page raw --csv.
"""
import argparse, csv, json, os, re
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("lra::", "")


def num(x):
    try:
        v = float(x.replace(",", ""))
        return v if v == v else None
    except Exception:
        return None


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    i_name = hdr.index("Kernel Name"); i_val = hdr.index("Metric Value"); i_unit = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= i_val:
            continue
        v = num(r[i_val])
        if v is None:
            continue
        u = r[i_unit]
        v_us = v / 1000.0 if u.startswith("n") else (v * 1000.0 if u.startswith("m") else v)
        a = agg.setdefault(short(r[i_name]), [0, 0.0])
        a[0] += 1; a[1] += v_us
    return agg


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        g = lambda k: num(r[ix[k]]) if k in ix else None
        def mb(k):
            v = g(k)
            if v is None: return None
            u = units[ix[k]].lower()
            return v * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1.0)
        out.append(dict(name=short(r[ix["Kernel Name"]]), us=g("gpu__time_duration.sum"), regs=g("launch__registers_per_thread"),
                        grid=g("launch__grid_size"), occ=g("sm__warps_active.avg.pct_of_peak_sustained_active"),
                        ipc=g("sm__inst_executed.avg.per_cycle_elapsed"), dram_r=mb("dram__bytes_read.sum"), dram_w=mb("dram__bytes_write.sum"),
                        dram_pct=g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), sm_pct=g("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                        l1_hit=g("l1tex__t_sector_hit_rate.pct"), l2_hit=g("lts__t_sector_hit_rate.pct"),
                        lanes=g("smsp__thread_inst_executed_per_inst_executed.ratio")))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches"); ap.add_argument("--raw"); ap.add_argument("--tag", required=True)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    lines = ["# ncu summary %s" % a.tag, "", a.note, ""]
    traffic = {}
    if a.launches:
        agg = launches(a.launches)
        tot = sum(v[1] for v in agg.values())
        lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)", "",
                  "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append("| %s | %d | %.1f | %.1f %% |" % (k, n, us, 100 * us / tot))
        lines.append("")
    if a.raw:
        lines += ["## `ncu --set full` per launch", "",
                  "| kernel | us | regs | grid | warps active % | IPC/SM | DRAM rd MB | DRAM wr MB | DRAM % | SM % | L1 hit % | L2 hit % | active lanes/inst |",
                  "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
        f = lambda v, p=1: "-" if v is None else ("%." + str(p) + "f") % v
        for r in raw(a.raw):
            lines.append("| %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (
                r["name"], f(r["us"]), f(r["regs"], 0), f(r["grid"], 0), f(r["occ"]), f(r["ipc"], 2), f(r["dram_r"], 2), f(r["dram_w"], 2),
                f(r["dram_pct"]), f(r["sm_pct"]), f(r["l1_hit"]), f(r["l2_hit"]), f(r["lanes"])))
            if r["dram_r"] is not None and r["dram_r"] == r["dram_r"] and r["name"] not in traffic:
                traffic[r["name"]] = int((r["dram_r"] + (r["dram_w"] or 0)) * 1e6)
        lines.append("")
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", "%s_ncu_summary.md" % a.tag), "w").write("\n".join(lines))
    if traffic:
        json.dump(traffic, open(os.path.join(ROOT, "profiles", "%s_traffic_raw.json" % a.tag), "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
