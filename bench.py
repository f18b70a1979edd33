#!/usr/bin/env python
"""bench.py -- throughput of the B200-native MapRead hot path on BASELINE.json's configs[1] workload
(100k synthetic ONT reads, N50 20 kb, 8 % error, vs a 3 Gb synthetic reference, `-ONT`).

What one "step" is: one pass of the hot path over one batch of READS_PER_STEP synthetic reads.  The stages of MapRead
that are on the GPU so far are listed in config.stages (SURVEY.md section 8(a) row ids); a step runs exactly those
stages over the job stream those reads generate in the reference (job shapes drawn from tables captured from the
reference on reads of this profile, job content synthetic -- tools/workload.py).  The metric is therefore named
"reads/sec (<stages>)" until every stage of MapRead is covered: it is NOT a whole-aligner reads/sec yet and is not
presented as one.  `--impl reference` runs the reference's own CPU code for the same stages on the same jobs.

  value     jobs/arenas resident in HBM, results left in HBM (kernel pipeline only)
  e2e       the same through the host-buffer C-ABI call: ASCII read arena + job arrays H2D, results D2H, every step
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
STAGES = ["a18:AffineOneGapAlign"]
METRIC = "reads/sec (MapRead stages on GPU so far: %s)" % ",".join(STAGES)
WORKLOAD = ("BASELINE configs[1]: synthetic ONT reads (N50 20 kb, 8%% err) vs 3 Gb synthetic ref (24 x 125 Mb), -ONT; "
            "per step the AffineOneGapAlign job stream of %d reads (%.1f jobs/read, shapes captured from the reference)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads-per-step", type=int, default=2048)
    ap.add_argument("--genome-len", type=int, default=3_000_000_000)
    ap.add_argument("--cpu-sample-jobs", type=int, default=400_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def gen_genome_device(n, device, seed=1234):
    """Uniform i.i.d. ACGT as ASCII on the GPU, 24 equal contigs (contig structure is irrelevant to this stage)."""
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


def run_reference(args, jobs_per_read):
    """The reference's own CPU implementation of the covered stages (oracle/_ref/libref_lra.so = unmodified reference
    headers; else the C restatement), all host threads, on a bounded sample of each step's jobs."""
    from oracle import pyoracle as po
    import synth
    cores = os.cpu_count() or 1
    n_jobs = min(args.cpu_sample_jobs, int(round(args.reads_per_step * jobs_per_read)))
    genome = synth.gen_ref(50_000_000, 1, 1234)[0][1]
    jobs = workload.make_jobs(PROFILE, n_jobs, 1000, len(genome), workload.host_genome_fetcher(genome))
    m, mm, indel = jobs["scoring"]
    kind = "reference" if po.ref() is not None else "port"

    def one():
        t0 = time.perf_counter()
        if kind == "reference":
            po.aog_batch_ref(jobs["q_arena"], jobs["t_arena_compact"], jobs["q_off"], jobs["t_off_compact"], jobs["q_len"],
                             jobs["t_len"], jobs["k"], m, mm, indel, nthreads=cores, want_blocks=True)
        else:
            po.aog_batch_port(jobs["q_arena"], jobs["t_arena_compact"], jobs["q_off"], jobs["t_off_compact"], jobs["q_len"],
                              jobs["t_len"], jobs["k"], m, mm, indel)
        return time.perf_counter() - t0
    for _ in range(args.warmup):
        one()
    ts = [one() for _ in range(args.steps)]
    total = sum(ts)
    reads = n_jobs / jobs_per_read
    value = reads * args.steps / total
    sample = "%d AffineOneGapAlign jobs (= %.0f reads) per step, %s" % (n_jobs, reads, "unmodified reference header via oracle/_ref/libref_lra.so" if kind == "reference" else "C restatement oracle/aog.c")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": WORKLOAD % (args.reads_per_step, jobs_per_read), "stages": STAGES, "profile": PROFILE},
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores if kind == "reference" else 1, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "aog_jobs_per_s": n_jobs * args.steps / total}
    print(json.dumps(line))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    jobs_per_read = workload.meta()[PROFILE]["aog_jobs_per_read"]
    if args.impl == "reference":
        if rank == 0:
            run_reference(args, jobs_per_read)
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
    n_jobs = int(round(args.reads_per_step * jobs_per_read))

    # ---- setup (untimed): genome resident + packed, NB distinct batches per rank
    # all device work (torch's and the library's) goes to ONE explicit stream, so the CUDA events below see it all
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    genome = gen_genome_device(args.genome_len, dev)
    ctx = lra_b200.Context(local)
    ctx.set_stream(stream.cuda_stream)
    tseq = ctx.seq_from_device(genome.data_ptr(), args.genome_len)
    fetch = workload.torch_genome_fetcher(genome)
    NB = 3
    batches = []
    for bi in range(NB):
        j = workload.make_jobs(PROFILE, n_jobs, 7919 * (rank + 1) + bi, args.genome_len, fetch)
        hb = {}
        for key in ["q_off", "t_off", "q_len", "t_len", "k"]:
            hb[key + "_t"], hb[key] = pinned(j[key])
        hb["q_arena_t"], hb["q_arena"] = pinned(j["q_arena"])
        cap = int(np.minimum(j["q_len"], j["t_len"]).sum()) + 1
        hb["cap"] = cap
        hb["out"] = {}
        for key, shape, dt in [("score", n_jobs, np.int32), ("n_blocks", n_jobs, np.int32), ("block_off", n_jobs, np.uint64),
                               ("blocks", (cap, 3), np.uint32)]:
            hb[key + "_ot"], hb["out"][key] = pinned(np.zeros(shape, dt))
        # device-resident copies for the kernel-only path
        db = {key: torch.from_numpy(j[key].view(np.int32)).to(dev) for key in ["q_off", "t_off", "q_len", "t_len", "k"]}
        db["qseq"] = ctx.seq_upload(j["q_arena"][:-16])
        db["score"] = torch.empty(n_jobs, dtype=torch.int32, device=dev)
        db["n_blocks"] = torch.empty(n_jobs, dtype=torch.int32, device=dev)
        db["block_off"] = torch.empty(n_jobs, dtype=torch.int64, device=dev)
        db["blocks"] = torch.empty((cap, 3), dtype=torch.int32, device=dev)
        hb["dev"] = db
        hb["scoring"] = j["scoring"]
        hb["jobs"] = j if (bi == 0 and rank == 0) else None
        batches.append(hb)
    del genome, fetch
    torch.cuda.empty_cache()
    eseq = ctx.seq_upload(batches[0]["q_arena"][:-16])  # arena re-used by the e2e path
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_value(hb):
        d = hb["dev"]; m, mm, indel = hb["scoring"]
        return ctx.aog_batch_device(d["qseq"], tseq, d["q_off"].data_ptr(), d["t_off"].data_ptr(), d["q_len"].data_ptr(),
                                    d["t_len"].data_ptr(), d["k"].data_ptr(), n_jobs, m, mm, indel, d["score"].data_ptr(),
                                    d["n_blocks"].data_ptr(), d["block_off"].data_ptr(), d["blocks"].data_ptr(), hb["cap"])

    def step_e2e(hb):
        m, mm, indel = hb["scoring"]
        eseq.reupload(hb["q_arena"][:-16])
        return ctx.aog_batch(eseq, tseq, hb["q_off"], hb["t_off"], hb["q_len"], hb["t_len"], hb["k"], m, mm, indel,
                             block_cap=hb["cap"], out=hb["out"])

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
    kstats, cells_total, blocks_total = {}, [0], [0]

    def collect(r):
        blocks_total[0] += r[0]; cells_total[0] += r[1]
        for s in ctx.kernel_stats():
            a = kstats.setdefault(s["name"], dict(ms=0.0, jobs=0, cells=0, algo_bytes=0, launches=0))
            a["ms"] += s["ms"]; a["jobs"] += s["jobs"]; a["cells"] += s["cells"]; a["algo_bytes"] += s["algo_bytes"]
            a["launches"] += 1
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
    nbt = batches[0]["out"]["n_blocks_total"] if "n_blocks_total" in batches[0]["out"] else 0
    h2d = int(len(batches[0]["q_arena"]) - 16 + 5 * 4 * n_jobs)
    d2h = int(n_jobs * (4 + 4 + 8) + 12 * nbt)

    if world > 1:
        t = torch.tensor([ms_value, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_value, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    reads_total = args.reads_per_step * args.steps * world
    value = reads_total / (ms_value / 1000.0)
    e2e = reads_total / (ms_e2e / 1000.0)
    peak, peak_src = peaks()
    dom_name = max(kstats, key=lambda k: kstats[k]["ms"])
    dom = kstats[dom_name]
    achieved = dom["algo_bytes"] / (dom["ms"] / 1000.0) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom_name)
    except Exception:
        pass
    line = {"metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD % (args.reads_per_step, jobs_per_read), "stages": STAGES, "profile": PROFILE,
                       "jobs_per_step": n_jobs, "genome_len": args.genome_len, "l2": "flushed between timed steps (256 MiB fill)",
                       "parallelism": "reads sharded over %d GPU(s), no data-path collective" % world},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algo_bytes_per_launch": dom["algo_bytes"] / dom["launches"], "ms_per_launch": dom["ms"] / dom["launches"],
                         "gcups": dom["cells"] / (dom["ms"] / 1000.0) / 1e9 if dom["cells"] else None,
                         "note": "integer DP: bound by ALU issue / shuffle latency, not HBM (SURVEY.md 8(d)); GCUPS is the telling figure"},
            "aog_jobs_per_s": n_jobs * args.steps * world / (ms_value / 1000.0),
            "gcups": cells_total[0] * world / (ms_value / 1000.0) / 1e9,
            "kernels": {k: {"ms_per_step": v["ms"] / args.steps, "jobs_per_step": v["jobs"] / args.steps,
                            "gcups": (v["cells"] / (v["ms"] / 1000.0) / 1e9) if v["cells"] and v["ms"] > 0 else None,
                            "algo_GBps": v["algo_bytes"] / (v["ms"] / 1000.0) / 1e9 if v["ms"] > 0 else None} for k, v in kstats.items()}}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, batches[0]["jobs"], jobs_per_read)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, jobs, jobs_per_read):
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    n = min(args.cpu_sample_jobs, len(jobs["q_off"]))
    m, mm, indel = jobs["scoring"]
    kind = "reference" if po.ref() is not None else "port"
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
    ts, tot = [], 0.0
    while tot < 4.0 and len(ts) < 200:
        ts.append(one()); tot += ts[-1]
    v = n * len(ts) / tot / jobs_per_read
    return {"value": v, "unit": "reads/s", "cores": cores if kind == "reference" else 1, "kind": kind,
            "sample": "%d AffineOneGapAlign jobs of step 0 x %d passes (%.1f s wall)" % (n, len(ts), tot),
            "aog_jobs_per_s": n * len(ts) / tot}


if __name__ == "__main__":
    main()
