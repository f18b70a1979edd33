#!/usr/bin/env python
"""bench.py -- throughput of the B200-native MapRead hot path on BASELINE.json's configs[1] workload
(100k synthetic ONT reads, N50 20 kb, 8 % error, vs a 3 Gb synthetic reference, `-ONT`).

What one "step" is: one pass of the hot path over one batch of --reads-per-step synthetic reads.  The stages of MapRead
that are on the GPU so far are listed in config.stages (SURVEY.md section 8(a) row ids); a step runs exactly those stages
over the work those reads generate in the reference:
  a18  AffineOneGapAlign      the job stream of the reads (job shapes drawn from tables captured from the reference on reads
                              of this profile, job content synthetic; tools/workload.py make_jobs)
  a19  IndelRefineAlignment   one segment per read (block list = gap-free runs of the read's true alignment; make_segments)
The metric is named "reads/sec (<stages>)" until every stage of MapRead is covered: it is NOT a whole-aligner reads/sec
yet and is not presented as one.  `--impl reference` runs the reference's own CPU code for the same stages on the same
kind of work, with all host threads.

  value     arenas / jobs / segments resident in HBM, results left in HBM (kernel pipeline only)
  e2e       the same through the host-buffer C-ABI calls: ASCII read arenas + descriptors H2D, results D2H, every step
  roofline  dominant kernel: algorithmic bytes / CUDA-event time vs the measured HBM peak (MEASURED_PEAKS.json)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

import workload  # noqa: E402

PROFILE = "ont"
STAGES = ["a18:AffineOneGapAlign", "a19:IndelRefineAlignment"]
WORKLOAD = ("BASELINE configs[1]: synthetic ONT reads (N50 20 kb, 8%% err) vs 3 Gb synthetic ref (24 x 125 Mb), -ONT; per step %d reads: "
            "their AffineOneGapAlign job stream (%.1f jobs/read, shapes captured from the reference) and one IndelRefineAlignment "
            "segment per read")


def metric_name(stages):
    return "reads/sec (MapRead stages on GPU so far: %s)" % ",".join(s for s in STAGES if s[:3] in stages)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads-per-step", type=int, default=16384)
    ap.add_argument("--genome-len", type=int, default=3_000_000_000)
    ap.add_argument("--cpu-sample-jobs", type=int, default=400_000)
    ap.add_argument("--cpu-sample-segments", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stages", default="a18,a19")
    return ap.parse_args()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mx = float(f[1])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[0]))
                    for nm, v in zip(names, f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
            except ValueError:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def gen_genome_device(n, device, seed=1234):
    """Uniform i.i.d. ACGT as ASCII on the GPU (contig structure is irrelevant to these stages)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty(n + 64, dtype=torch.uint8, device=device)
    out[n:] = ord("N")
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=device)
    CH = 1 << 28
    for s in range(0, n, CH):
        e = min(n, s + CH)
        c = torch.randint(0, 4, (e - s,), generator=g, device=device, dtype=torch.int64)
        out[s:e] = lut[c]
        del c
    return out


def pinned(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


# ------------------------------------------------------------------------------------------------------------ CPU arms
def cpu_stage_rates(args, jobs, segs, jobs_per_read, budget_s):
    """Time the reference's own functions (oracle/_ref/libref_lra.so = unmodified reference headers; else the C restatement) on
    bounded samples, all host threads.  Returns (reads/s over the covered stages, description dict)."""
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    kind = "reference" if po.ref() is not None else "port"
    per_read, parts = 0.0, {}
    if jobs is not None:
        n = min(args.cpu_sample_jobs, len(jobs["q_off"]))
        m, mm, indel = jobs["scoring"]
        sub = {k: np.ascontiguousarray(jobs[k][:n]) for k in ["q_off", "t_off_compact", "q_len", "t_len", "k"]}

        def one():
            t0 = time.perf_counter()
            if kind == "reference":
                po.aog_batch_ref(jobs["q_arena"], jobs["t_arena_compact"], sub["q_off"], sub["t_off_compact"], sub["q_len"], sub["t_len"],
                                 sub["k"], m, mm, indel, nthreads=cores)
            else:
                po.aog_batch_port(jobs["q_arena"], jobs["t_arena_compact"], sub["q_off"], sub["t_off_compact"], sub["q_len"], sub["t_len"],
                                  sub["k"], m, mm, indel)
            return time.perf_counter() - t0
        one()
        ts, tot = [one()], 0.0
        tot = ts[0]
        while tot < budget_s and len(ts) < 200:
            ts.append(one()); tot += ts[-1]
        rate = n * len(ts) / tot
        per_read += jobs_per_read / rate
        parts["a18"] = {"jobs_per_s": rate, "sample": "%d jobs x %d passes" % (n, len(ts))}
    if segs is not None:
        n = min(args.cpu_sample_segments, len(segs["blk_cnt"]))
        sub = dict(segs)
        for k in ["blk_off", "blk_cnt", "q_base", "read_len", "contig_len"]:
            sub[k] = np.ascontiguousarray(segs[k][:n])
        tb = np.ascontiguousarray(segs["t_base_compact"][:n])

        def one2():
            t0 = time.perf_counter()
            if kind == "reference":
                po.indel_refine_batch_ref(sub, segs["t_arena_compact"], tb, nthreads=cores, want_blocks=True)
            else:
                po.indel_refine_batch_port(sub, segs["t_arena_compact"], tb)
            return time.perf_counter() - t0
        one2()
        ts = [one2()]
        tot = ts[0]
        while tot < budget_s and len(ts) < 200:
            ts.append(one2()); tot += ts[-1]
        rate = n * len(ts) / tot
        per_read += 1.0 / rate
        parts["a19"] = {"segments_per_s": rate, "sample": "%d segments x %d passes" % (n, len(ts))}
    desc = {"kind": kind, "cores": cores if kind == "reference" else 1, "parts": parts}
    return 1.0 / per_read, desc


def run_reference(args, jobs_per_read, stages):
    import synth
    genome = synth.gen_ref(100_000_000, 1, 1234)[0][1]
    fetch = workload.host_genome_fetcher(genome)
    n_jobs = min(args.cpu_sample_jobs, int(round(args.reads_per_step * jobs_per_read)))
    jobs = workload.make_jobs(PROFILE, n_jobs, 1000, len(genome), fetch) if "a18" in stages else None
    segs = workload.make_segments(PROFILE, min(args.cpu_sample_segments, args.reads_per_step), 1001, len(genome), fetch) if "a19" in stages else None
    rates, desc = [], None
    for i in range(args.warmup + args.steps):       # each step = one pass over the bounded samples
        r, desc = cpu_stage_rates(args, jobs, segs, jobs_per_read, budget_s=0.0)
        if i >= args.warmup:
            rates.append(r)
    value = len(rates) / sum(1.0 / r for r in rates)
    sample = "; ".join("%s: %s" % (k, v["sample"].split(" x ")[0]) for k, v in desc["parts"].items()) + " per step"
    line = {"impl": "reference", "metric": metric_name(stages), "value": value, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * args.reads_per_step / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": WORKLOAD % (args.reads_per_step, jobs_per_read), "stages": [s for s in STAGES if s[:3] in stages], "profile": PROFILE},
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": desc["cores"], "kind": desc["kind"], "sample": sample, "parts": desc["parts"]},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    stages = args.stages.split(",")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    jobs_per_read = workload.meta()[PROFILE]["aog_jobs_per_read"]
    if args.impl == "reference":
        if rank == 0:
            run_reference(args, jobs_per_read, stages)
        return
    import torch
    import torch.distributed as dist
    import lra_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lra_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    R = args.reads_per_step
    n_jobs = int(round(R * jobs_per_read))

    # all device work (torch's and the library's) goes to ONE explicit stream, so the CUDA events below see it all
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    genome = gen_genome_device(args.genome_len, dev)
    ctx = lra_b200.Context(local)
    ctx.set_stream(stream.cuda_stream)
    tseq = ctx.seq_from_device(genome.data_ptr(), args.genome_len)
    fetch = workload.torch_genome_fetcher(genome)
    NB = 2
    batches = []
    for bi in range(NB):
        hb = {"jobs0": None, "segs0": None}
        seed = 7919 * (rank + 1) + bi
        if "a18" in stages:
            j = workload.make_jobs(PROFILE, n_jobs, seed, args.genome_len, fetch)
            A = {"scoring": j["scoring"]}
            for key in ["q_off", "t_off", "q_len", "t_len", "k"]:
                A[key + "_t"], A[key] = pinned(j[key])
            A["q_arena_t"], A["q_arena"] = pinned(j["q_arena"])
            A["cap"] = int(np.minimum(j["q_len"], j["t_len"]).sum()) + 1
            A["out"] = {}
            for key, shape, dt in [("score", n_jobs, np.int32), ("n_blocks", n_jobs, np.int32), ("block_off", n_jobs, np.uint64),
                                   ("blocks", (A["cap"], 3), np.uint32)]:
                A[key + "_ot"], A["out"][key] = pinned(np.zeros(shape, dt))
            d = {key: torch.from_numpy(j[key].view(np.int32)).to(dev) for key in ["q_off", "t_off", "q_len", "t_len", "k"]}
            d["qseq"] = ctx.seq_upload(j["q_arena"][:-16])
            d["score"] = torch.empty(n_jobs, dtype=torch.int32, device=dev)
            d["n_blocks"] = torch.empty(n_jobs, dtype=torch.int32, device=dev)
            d["block_off"] = torch.empty(n_jobs, dtype=torch.int64, device=dev)
            d["blocks"] = torch.empty((A["cap"], 3), dtype=torch.int32, device=dev)
            A["dev"] = d
            hb["aog"] = A
            if bi == 0 and rank == 0:
                hb["jobs0"] = j
        if "a19" in stages:
            sg = workload.make_segments(PROFILE, R, seed + 17, args.genome_len, fetch)
            I = {k: sg[k] for k in ["k", "match", "mismatch", "indel", "end_align"]}
            I["T"] = len(sg["blocks_in"])
            for key in ["blocks_in", "blk_off", "blk_cnt", "q_base", "t_base", "read_len", "contig_len"]:
                I[key + "_t"], I[key] = pinned(sg[key])
            I["q_arena_t"], I["q_arena"] = pinned(sg["q_arena"])
            I["cap"] = 2 * I["T"] + 64 * R + 1024
            I["out"] = {}
            for key, shape, dt in [("n_blocks", R, np.int32), ("block_off", R, np.uint64), ("blocks", (I["cap"], 3), np.uint32)]:
                I[key + "_ot"], I["out"][key] = pinned(np.zeros(shape, dt))
            d = {"blocks_in": torch.from_numpy(sg["blocks_in"].view(np.int32)).to(dev), "blk_off": torch.from_numpy(sg["blk_off"].view(np.int64)).to(dev)}
            for key in ["blk_cnt", "q_base", "t_base", "read_len", "contig_len"]:
                d[key] = torch.from_numpy(sg[key].view(np.int32)).to(dev)
            d["qseq"] = ctx.seq_upload(sg["q_arena"][:-16])
            d["n_blocks"] = torch.empty(R, dtype=torch.int32, device=dev)
            d["block_off"] = torch.empty(R, dtype=torch.int64, device=dev)
            d["blocks"] = torch.empty((I["cap"], 3), dtype=torch.int32, device=dev)
            I["dev"] = d
            hb["ir"] = I
            if bi == 0 and rank == 0:
                hb["segs0"] = sg
        batches.append(hb)
    del genome, fetch
    torch.cuda.empty_cache()
    eseq_a = ctx.seq_upload(batches[0]["aog"]["q_arena"][:-16]) if "a18" in stages else None
    eseq_i = ctx.seq_upload(batches[0]["ir"]["q_arena"][:-16]) if "a19" in stages else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_value(hb):
        out = {"cells": 0, "stats": []}
        if "a18" in stages:
            A = hb["aog"]; d = A["dev"]; m, mm, indel = A["scoring"]
            nbt, cells = ctx.aog_batch_device(d["qseq"], tseq, d["q_off"].data_ptr(), d["t_off"].data_ptr(), d["q_len"].data_ptr(),
                                              d["t_len"].data_ptr(), d["k"].data_ptr(), n_jobs, m, mm, indel, d["score"].data_ptr(),
                                              d["n_blocks"].data_ptr(), d["block_off"].data_ptr(), d["blocks"].data_ptr(), A["cap"])
            out["cells"] += cells; out["stats"] += ctx.kernel_stats()
        if "a19" in stages:
            I = hb["ir"]; d = I["dev"]
            r = ctx.indel_refine_batch_device(d["qseq"], tseq, [d[k].data_ptr() for k in ["blocks_in", "blk_off", "blk_cnt", "q_base", "t_base", "read_len", "contig_len"]],
                                              I["T"], R, I["k"], I["match"], I["mismatch"], I["indel"], I["end_align"], d["n_blocks"].data_ptr(),
                                              d["block_off"].data_ptr(), d["blocks"].data_ptr(), I["cap"])
            out["cells"] += r["cells"]; out["stats"] += ctx.kernel_stats()
        return out

    io = {"h2d": 0, "d2h": 0}

    def step_e2e(hb):
        h2d = d2h = 0
        if "a18" in stages:
            A = hb["aog"]; m, mm, indel = A["scoring"]
            eseq_a.reupload(A["q_arena"][:-16])
            r = ctx.aog_batch(eseq_a, tseq, A["q_off"], A["t_off"], A["q_len"], A["t_len"], A["k"], m, mm, indel, block_cap=A["cap"], out=A["out"])
            h2d += len(A["q_arena"]) - 16 + 5 * 4 * n_jobs
            d2h += n_jobs * 16 + 12 * r["n_blocks_total"]
        if "a19" in stages:
            I = hb["ir"]
            eseq_i.reupload(I["q_arena"][:-16])
            r = ctx.indel_refine_batch(eseq_i, tseq, I, block_cap=I["cap"], out=I["out"])
            h2d += len(I["q_arena"]) - 16 + 12 * I["T"] + R * (8 + 5 * 4)
            d2h += R * 12 + 12 * r["n_blocks_total"]
        io["h2d"], io["d2h"] = h2d, d2h
        return None

    def timed(fn, steps, collect=None):
        tot_ms = 0.0
        for s in range(steps):
            hb = batches[s % NB]
            flush.fill_(s & 255)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(hb)
            e1.record()
            e1.synchronize()
            tot_ms += e0.elapsed_time(e1)
            if collect is not None:
                collect(r)
        return tot_ms

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- kernel-only ("value")
    timed(step_value, args.warmup)
    kstats, cells_total = {}, [0]

    def collect(r):
        cells_total[0] += r["cells"]
        for s in r["stats"]:
            a = kstats.setdefault(s["name"], dict(ms=0.0, jobs=0, cells=0, algo_bytes=0, launches=0))
            a["ms"] += s["ms"]; a["jobs"] += s["jobs"]; a["cells"] += s["cells"]; a["algo_bytes"] += s["algo_bytes"]; a["launches"] += 1
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ctx.launch_count()
    t0 = time.time()
    ms_value = timed(step_value, args.steps, collect)
    sync_all()
    t1 = time.time()
    launches = ctx.launch_count() - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    # ---- end to end through the host-buffer C ABI
    timed(step_e2e, max(1, args.warmup))
    sync_all()
    ms_e2e = timed(step_e2e, args.steps)
    sync_all()

    if world > 1:
        t = torch.tensor([ms_value, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_value, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    reads_total = R * args.steps * world
    value = reads_total / (ms_value / 1000.0)
    e2e = reads_total / (ms_e2e / 1000.0)
    peak, peak_src = peaks()
    dp = {k: v for k, v in kstats.items() if v["algo_bytes"] > 0} or kstats
    dom_name = max(dp, key=lambda k: dp[k]["ms"])
    dom = kstats[dom_name]
    achieved = dom["algo_bytes"] / (dom["ms"] / 1000.0) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom_name)
    except Exception:
        pass
    line = {"metric": metric_name(stages), "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD % (R, jobs_per_read), "stages": [s for s in STAGES if s[:3] in stages], "profile": PROFILE,
                       "reads_per_step": R, "aog_jobs_per_step": n_jobs if "a18" in stages else 0, "genome_len": args.genome_len,
                       "l2": "flushed between timed steps (256 MiB fill)",
                       "parallelism": "reads sharded over %d GPU(s), no data-path collective" % world},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": int(io["h2d"]), "d2h_bytes_per_step": int(io["d2h"]),
                    "ms_per_step": ms_e2e / args.steps},
            "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algo_bytes_per_launch": dom["algo_bytes"] / dom["launches"], "ms_per_launch": dom["ms"] / dom["launches"],
                         "gcups": dom["cells"] / (dom["ms"] / 1000.0) / 1e9 if dom["cells"] else None,
                         "note": "integer DP: bound by ALU issue / shuffle latency, not HBM (SURVEY.md 8(d)); GCUPS is the telling figure"},
            "gcups": cells_total[0] * world / (ms_value / 1000.0) / 1e9,
            "kernels": {k: {"ms_per_step": v["ms"] / args.steps, "units_per_step": v["jobs"] / args.steps,
                            "gcups": (v["cells"] / (v["ms"] / 1000.0) / 1e9) if v["cells"] and v["ms"] > 0 else None,
                            "algo_GBps": v["algo_bytes"] / (v["ms"] / 1000.0) / 1e9 if v["ms"] > 0 else None} for k, v in kstats.items()}}
    if world == 1 and not args.no_cpu_baseline:
        v, desc = cpu_stage_rates(args, batches[0]["jobs0"], batches[0]["segs0"], jobs_per_read, budget_s=4.0)
        line["cpu_baseline"] = {"value": v, "unit": "reads/s", "cores": desc["cores"], "kind": desc["kind"],
                                "sample": "; ".join("%s: %s" % (k, p["sample"]) for k, p in desc["parts"].items()), "parts": desc["parts"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
