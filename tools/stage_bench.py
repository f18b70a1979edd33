#!/usr/bin/env python
"""tools/stage_bench.py (round 1's bench.py, kept for the per-stage kernel profiles; the headline bench is ../bench.py) -- throughput of the B200-native MapRead hot path on BASELINE.json's configs[1] workload
(100k synthetic ONT reads, N50 20 kb, 8 % error, vs a 3 Gb synthetic reference, `-ONT`).

What one "step" is: one pass of the hot path over one batch of --reads-per-step synthetic reads.  The stages of MapRead
that are on the GPU so far are listed in config.stages (SURVEY.md section 8(a) row ids); a step runs exactly those stages
over the work those reads generate in the reference:
  a18  AffineOneGapAlign      the job stream of the reads (job shapes drawn from tables captured from the reference on reads
                              of this profile, job content synthetic; tools/workload.py make_jobs)
  a19  IndelRefineAlignment   one segment per read (block list = gap-free runs of the read's true alignment; make_segments)
  a12  LocalIndex::IndexSeq   both strands of every read (CreateRC included): the per-read two-strand local index
  a13  Refine_splitchain      (the low-accuracy pipeline's local refinement, ONT is low-accuracy) one split chain per read (anchors = the exact
                              >= 17-base stretches of its alignment with their lengths; make_clusters) against the LocalIndex of the 3 Gb genome
                              (built once, on the GPU, outside the timed region like the reference's .gli load)
  a21  CalculateStatistics    CIGAR + NV of every segment a19 refined
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

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

import workload  # noqa: E402

PROFILE = "ont"
STAGES = ["a12:LocalIndex::IndexSeq", "a13:Refine_splitchain", "a18:AffineOneGapAlign", "a19:IndelRefineAlignment", "a21:CalculateStatistics"]
WORKLOAD = ("BASELINE configs[1]: synthetic ONT reads (N50 20 kb, 8%% err) vs 3 Gb synthetic ref (24 x 125 Mb), -ONT; per step %d reads: "
            "their AffineOneGapAlign job stream (%.1f jobs/read, shapes captured from the reference) and one IndelRefineAlignment "
            "segment per read; local index of both strands of every read, one cluster per read refined against the genome's local index, "
            "CIGAR/NV statistics of every refined segment")


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
    ap.add_argument("--stages", default="a12,a13,a18,a19,a21")
    ap.add_argument("--e2e-pipeline", action="store_true", help="end-to-end leg: two batches in flight on two lanes of two contexts each "
                                                                "(measured on B200: 147.7 vs 148.8 ms/step -- the leg is not host-bound, so it is off by default)")
    ap.add_argument("--serial-stages", action="store_true", help="run the two stage groups back to back on one context instead of concurrently on two")
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
    if segs is not None and "a19" in args.stages.split(","):
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
    stages = args.stages.split(",")
    if segs is not None and ("a12" in stages or "a13" in stages):
        # the reference's LocalIndex::IndexSeq on both strands of every read + REFINEclusters on its cluster, all host threads
        # (one ctypes call per read and stage: the GIL is released inside); the genome's LocalIndex is built once, untimed
        from concurrent.futures import ThreadPoolExecutor
        n = min(args.cpu_sample_segments, len(segs["blk_cnt"]))
        which = "ref" if kind == "reference" else "port"
        cl = workload.make_clusters(segs, compact=True)
        comp = np.zeros(256, np.uint8); comp[[65, 67, 71, 84, 78]] = [84, 71, 67, 65, 78]
        ta = segs["t_arena_compact"]; hdr = cl["hdr_pos"]
        contigs = [ta[int(hdr[i]):int(hdr[i + 1])] for i in range(n)] + [ta[int(hdr[-2]):int(hdr[-1])]]
        hdr = np.concatenate([hdr[:n + 1], [hdr[n] + np.uint64(len(contigs[-1]))]]).astype(np.uint64)
        glh = po.RefLocalIndexHandle(contigs) if which == "ref" else None
        gl = po.local_index(contigs, which="port") if which == "port" else None
        hdr_n = np.ascontiguousarray(hdr)

        def one_read(i):
            qb = int(segs["q_base"][i]); L = int(segs["read_len"][i])
            r = segs["q_arena"][qb:qb + L]; rc = comp[r[::-1]]
            a, b = int(cl["m_off"][i]), int(cl["m_off"][i + 1])
            if which == "ref":
                f = po.RefLocalIndexHandle(r); v = po.RefLocalIndexHandle(rc)
                if "a13" in stages and cl["chain_ok"][i]:
                    po.refine_splitchain(cl["m_q"][a:b], cl["m_t"][a:b], cl["m_len"][a:b], cl["m_strand"][a:b], cl["chain_box"][i], int(cl["chrom"][i]), 0, L, hdr_n,
                                         None, None, None, 17, which="ref", ref_handles=(glh.h, f.h, v.h))
                f.close(); v.close()
            else:
                f = po.local_index(r); v = po.local_index(rc)
                if "a13" in stages and cl["chain_ok"][i]:
                    po.refine_splitchain(cl["m_q"][a:b], cl["m_t"][a:b], cl["m_len"][a:b], cl["m_strand"][a:b], cl["chain_box"][i], int(cl["chrom"][i]), 0, L, hdr_n,
                                         gl, f, v, 17, which="port")

        def one3():
            t0 = time.perf_counter()
            with ThreadPoolExecutor(cores if which == "ref" else 1) as ex:
                list(ex.map(one_read, range(n)))
            return time.perf_counter() - t0
        one_read(0)      # (binds the ctypes prototypes on this thread before the pool starts)
        one3()
        ts = [one3()]
        tot = ts[0]
        while tot < budget_s and len(ts) < 200:
            ts.append(one3()); tot += ts[-1]
        rate = n * len(ts) / tot
        per_read += 1.0 / rate
        parts["a12+a13"] = {"reads_per_s": rate, "sample": "%d reads x %d passes" % (n, len(ts))}
        if glh:
            glh.close()
    if segs is not None and "a21" in stages:
        from concurrent.futures import ThreadPoolExecutor
        n = min(args.cpu_sample_segments, len(segs["blk_cnt"]))
        sub = dict(segs)
        for k in ["blk_off", "blk_cnt", "q_base", "read_len", "contig_len"]:
            sub[k] = np.ascontiguousarray(segs[k][:n])
        tb = np.ascontiguousarray(segs["t_base_compact"][:n])
        if kind == "reference":
            nb, off, blk = po.indel_refine_batch_ref(sub, segs["t_arena_compact"], tb, nthreads=cores, want_blocks=True)
            refined = [blk[int(off[i]):int(off[i]) + int(nb[i])] for i in range(n)]
        else:
            refined = po.indel_refine_batch_port(sub, segs["t_arena_compact"], tb)
        items = []
        for i in range(n):
            qb, t0_ = int(segs["q_base"][i]), int(tb[i])
            items.append((segs["q_arena"][qb:qb + int(segs["read_len"][i])].tobytes(), segs["t_arena_compact"][t0_:t0_ + int(segs["contig_len"][i])].tobytes(), refined[i]))

        def one_seg(it):
            if kind == "reference":
                po.calc_stats_ref(it[0], it[1], it[2])
            else:
                po.calc_stats_port(it[0], it[1], 0, it[2])

        def one4():
            t0 = time.perf_counter()
            with ThreadPoolExecutor(cores if kind == "reference" else 1) as ex:
                list(ex.map(one_seg, items))
            return time.perf_counter() - t0
        one_seg(items[0])
        one4()
        ts = [one4()]
        tot = ts[0]
        while tot < budget_s and len(ts) < 200:
            ts.append(one4()); tot += ts[-1]
        rate = n * len(ts) / tot
        per_read += 1.0 / rate
        parts["a21"] = {"segments_per_s": rate, "sample": "%d segments x %d passes" % (n, len(ts))}
    desc = {"kind": kind, "cores": cores if kind == "reference" else 1, "parts": parts}
    return 1.0 / per_read, desc


def run_reference(args, jobs_per_read, stages):
    import synth
    genome = synth.gen_ref(100_000_000, 1, 1234)[0][1]
    fetch = workload.host_genome_fetcher(genome)
    n_jobs = min(args.cpu_sample_jobs, int(round(args.reads_per_step * jobs_per_read)))
    jobs = workload.make_jobs(PROFILE, n_jobs, 1000, len(genome), fetch) if "a18" in stages else None
    segs = workload.make_segments(PROFILE, min(args.cpu_sample_segments, args.reads_per_step), 1001, len(genome), fetch) if any(x in stages for x in ("a19", "a12", "a13", "a21")) else None
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
    need_reads = any(x in stages for x in ("a19", "a12", "a13", "a21"))
    if "a21" in stages and "a19" not in stages:
        raise SystemExit("stage a21 (statistics) runs on the segments a19 refines: add a19 to --stages")

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
        if need_reads:
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
            if "a12" in stages or "a13" in stages:
                I["read_off"] = np.ascontiguousarray(sg["q_base"].astype(np.uint64)); I["read_len_u"] = np.ascontiguousarray(sg["read_len"].astype(np.uint32))
            if "a13" in stages:
                cl = workload.make_clusters(sg, genome_len=args.genome_len)
                if not cl["chain_ok"].all():          # a chain across a contig boundary never reaches Refine_splitchain: make it empty
                    keepa = np.repeat(cl["chain_ok"], np.diff(cl["m_off"].astype(np.int64)))
                    for key in ["m_q", "m_t", "m_len", "m_strand"]:
                        cl[key] = cl[key][keepa]
                    cnt = np.where(cl["chain_ok"], np.diff(cl["m_off"].astype(np.int64)), 0)
                    cl["m_off"] = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint64)
                cl["box"] = cl["chain_box"]
                I["cl"] = cl
                I["M"] = int(cl["m_off"][-1])
                cl["chrom"] = np.ascontiguousarray(cl["chrom"], np.int32)
                I["acap"] = int(sg["read_len"].sum()) // 2 + 4096
                I["rf_out"] = {}
                for key, shape, dt in [("status", R, np.int32), ("chrom", R, np.int32), ("diag", 2 * R, np.int64), ("r_off", R + 1, np.uint64), ("r_q", I["acap"], np.uint32),
                                       ("r_t", I["acap"], np.uint32), ("r_tup", I["acap"], np.uint32), ("rbox", 4 * R, np.uint32), ("eff", R, np.float32)]:
                    I["rf_" + key + "_ot"], I["rf_out"][key] = pinned(np.zeros(shape, dt))
                dc = {key: torch.from_numpy(np.ascontiguousarray(cl[key]).view(np.int32 if cl[key].dtype in (np.uint32, np.int32) else (np.int64 if cl[key].dtype == np.uint64 else np.uint8))).to(dev)
                      for key in ["m_q", "m_t", "m_len", "m_strand", "m_off", "box", "strand", "chrom", "read_id", "hdr_pos"]}
                for key, shape, dt in [("status", R, torch.int32), ("chrom", R, torch.int32), ("diag", 2 * R, torch.int64), ("r_off", R + 1, torch.int64),
                                       ("r_q", I["acap"], torch.int32), ("r_t", I["acap"], torch.int32), ("r_tup", I["acap"], torch.int32),
                                       ("rbox", 4 * R, torch.int32), ("eff", R, torch.float32)]:
                    dc["o_" + key] = torch.empty(shape, dtype=dt, device=dev)
                I["dcl"] = dc
            if "a21" in stages:
                I["ccap"] = 4 * I["T"] + 64 * R + 1024
                I["st_out"] = {}
                for key, shape, dt in [("stats", (R, 16), np.int32), ("value", R, np.float32), ("cigar_off", R + 1, np.uint64), ("cigar", I["ccap"], np.uint32)]:
                    I["st_" + key + "_ot"], I["st_out"][key] = pinned(np.zeros(shape, dt))
                d["st_stats"] = torch.empty(16 * R, dtype=torch.int32, device=dev); d["st_value"] = torch.empty(R, dtype=torch.float32, device=dev)
                d["st_off"] = torch.empty(R + 1, dtype=torch.int64, device=dev); d["st_cigar"] = torch.empty(I["ccap"], dtype=torch.int32, device=dev)
            I["dev"] = d
            hb["ir"] = I
            if bi == 0 and rank == 0:
                hb["segs0"] = sg
        batches.append(hb)
    del genome, fetch
    torch.cuda.empty_cache()
    eseq_a = ctx.seq_upload(batches[0]["aog"]["q_arena"][:-16]) if "a18" in stages else None
    eseq_i = ctx.seq_upload(batches[0]["ir"]["q_arena"][:-16]) if need_reads else None
    # the genome's LocalIndex (<ref>.gli): built once on the GPU, like the reference loads it once
    gli = None
    if "a13" in stages:
        hdr = batches[0]["ir"]["cl"]["hdr_pos"]
        gli = ctx.lindex_build(tseq, hdr[:-1], np.diff(hdr).astype(np.uint32))
    log_lut = lra_b200.CreateLookUpTable() if "a21" in stages else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    keep = {"rc": None, "rf": None, "rr": None}      # the reverse-complement arena and the two read LocalIndex images are rebuilt in place every step
    io = {"h2d": 0, "d2h": 0}
    # second context for the stage group a12+a13 (one context per worker thread)
    two = (not args.serial_stages) and ("a12" in stages or "a13" in stages) and any(x in stages for x in ("a18", "a19", "a21"))
    if two:
        from concurrent.futures import ThreadPoolExecutor
        stream_b = torch.cuda.Stream(device=dev)
        ctxb = lra_b200.Context(local)
        ctxb.set_stream(stream_b.cuda_stream)
        pool = ThreadPoolExecutor(1)
    else:
        ctxb, pool = ctx, None

    def value_group_a(hb, out):          # a18, a19, a21 on the first context
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
        if "a21" in stages:
            I = hb["ir"]; d = I["dev"]
            ctx.calc_stats_batch_device(d["qseq"], tseq, [d[k].data_ptr() for k in ["blocks", "block_off", "n_blocks", "q_base", "t_base", "read_len"]],
                                        I["cap"], R, log_lut, d["st_stats"].data_ptr(), d["st_value"].data_ptr(), d["st_off"].data_ptr(),
                                        d["st_cigar"].data_ptr(), I["ccap"])
            out["stats"] += ctx.kernel_stats()

    def value_group_b(hb, out):          # a12, a13 on the second context (its own stream): independent of group a within a batch
        if "a12" in stages or "a13" in stages:
            I = hb["ir"]; d = I["dev"]
            rc = keep["rc"] = ctxb.seq_revcomp(d["qseq"], I["read_off"], I["read_len_u"], reuse=keep["rc"])
            rf = keep["rf"] = ctxb.lindex_build(d["qseq"], I["read_off"], I["read_len_u"], reuse=keep["rf"]); st1 = ctxb.kernel_stats()
            rr = keep["rr"] = ctxb.lindex_build(rc, I["read_off"], I["read_len_u"], reuse=keep["rr"]); st2 = ctxb.kernel_stats()
            for a, b2 in zip(st1, st2):
                a["ms"] += b2["ms"]; a["jobs"] += b2["jobs"]; a["algo_bytes"] += b2["algo_bytes"]
            out["stats"] += st1
            if "a13" in stages:
                dc = I["dcl"]; cl = I["cl"]
                ctxb.refine_splitchains_batch_device(gli, rf, rr, dict(m_q=dc["m_q"].data_ptr(), m_t=dc["m_t"].data_ptr(), m_len=dc["m_len"].data_ptr(),
                                                                       m_strand=dc["m_strand"].data_ptr(), m_off=dc["m_off"].data_ptr(), box=dc["box"].data_ptr(),
                                                                       strand=dc["strand"].data_ptr(), chrom=dc["chrom"].data_ptr(), read_id=dc["read_id"].data_ptr(),
                                                                       hdr_pos=dc["hdr_pos"].data_ptr(), n_hdr=len(cl["hdr_pos"])),
                                                     R, I["M"], (cl["global_k"], cl["small_k"], cl["window"], cl["local_max_freq"], cl["limitrefine"]),
                                                     {k: dc["o_" + k].data_ptr() for k in ["status", "chrom", "diag", "r_off", "r_q", "r_t", "r_tup", "rbox", "eff"]}, I["acap"])
                out["stats"] += ctxb.kernel_stats()

    def run_groups(fa, fb, hb, oa, ob):
        """The two stage groups of a batch are independent: like two of the reference's worker threads, each drives its own context
        (stream); --serial-stages runs them back to back on one."""
        if pool is not None:
            fut = pool.submit(fb, hb, ob)
            fa(hb, oa)
            fut.result()
        else:
            fa(hb, oa); fb(hb, ob)

    def step_value(hb):
        oa, ob = {"cells": 0, "stats": []}, {"cells": 0, "stats": []}
        run_groups(value_group_a, value_group_b, hb, oa, ob)
        return {"cells": oa["cells"] + ob["cells"], "stats": oa["stats"] + ob["stats"]}

    # ---- end to end: a lane = what one pair of the reference's worker threads owns (two contexts, its upload arenas, its re-used images)
    lanes = [{"ctx": ctx, "ctxb": ctxb, "pool": pool, "keep": keep, "eseq_a": eseq_a, "eseq_i": eseq_i}]
    if two and args.e2e_pipeline:
        s2, s3 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        c2 = lra_b200.Context(local); c2.set_stream(s2.cuda_stream)
        c3 = lra_b200.Context(local); c3.set_stream(s3.cuda_stream)
        lanes.append({"ctx": c2, "ctxb": c3, "pool": ThreadPoolExecutor(1), "keep": {"rc": None, "rf": None, "rr": None}, "streams": (s2, s3),
                      "eseq_a": c2.seq_upload(batches[0]["aog"]["q_arena"][:-16]) if "a18" in stages else None,
                      "eseq_i": c2.seq_upload(batches[0]["ir"]["q_arena"][:-16]) if need_reads else None})

    def e2e_group_a(hb, acc, L):
        lctx = L["ctx"]
        h2d = d2h = 0
        if "a18" in stages:
            A = hb["aog"]; m, mm, indel = A["scoring"]
            L["eseq_a"].reupload(A["q_arena"][:-16])
            r = lctx.aog_batch(L["eseq_a"], tseq, A["q_off"], A["t_off"], A["q_len"], A["t_len"], A["k"], m, mm, indel, block_cap=A["cap"], out=A["out"])
            h2d += len(A["q_arena"]) - 16 + 5 * 4 * n_jobs
            d2h += n_jobs * 16 + 12 * r["n_blocks_total"]
        if "a19" in stages:
            I = hb["ir"]
            r = lctx.indel_refine_batch(L["eseq_i"], tseq, I, block_cap=I["cap"], out=I["out"])
            h2d += 12 * I["T"] + R * (8 + 5 * 4)
            d2h += R * 12 + 12 * r["n_blocks_total"]
            if "a21" in stages:
                nb = I["out"]["n_blocks"]; tot = int(r["n_blocks_total"])
                o = lctx.calc_stats_batch(L["eseq_i"], tseq, dict(blocks_in=I["out"]["blocks"][:tot], blk_off=I["out"]["block_off"], blk_cnt=nb, q_base=I["q_base"],
                                                                   t_base=I["t_base"], read_len=I["read_len"]), log_lut, cigar_cap=I["ccap"], out=I["st_out"])
                h2d += 12 * tot + R * 24 + 2001 * 4
                d2h += R * (64 + 4 + 8) + 4 * o["n_cigar_total"]
        acc["h2d"] += h2d; acc["d2h"] += d2h

    def e2e_group_b(hb, acc, L):
        lctxb, lkeep = L["ctxb"], L["keep"]
        h2d = d2h = 0
        if "a12" in stages or "a13" in stages:
            I = hb["ir"]
            rc = lkeep["rc"] = lctxb.seq_revcomp(L["eseq_i"], I["read_off"], I["read_len_u"], reuse=lkeep["rc"])
            rf = lkeep["rf"] = lctxb.lindex_build(L["eseq_i"], I["read_off"], I["read_len_u"], reuse=lkeep["rf"])
            rr = lkeep["rr"] = lctxb.lindex_build(rc, I["read_off"], I["read_len_u"], reuse=lkeep["rr"])
            h2d += 2 * 12 * R
            if "a13" in stages:
                o = lctxb.refine_splitchains_batch(gli, rf, rr, I["cl"], anchor_cap=I["acap"], out=I["rf_out"])
                h2d += 13 * I["M"] + R * (8 + 16 + 1 + 4 + 4)
                d2h += R * (4 + 4 + 16 + 8 + 16 + 4) + 12 * o["n_anchors"]
        acc["h2d"] += h2d; acc["d2h"] += d2h

    def step_e2e(hb, L):
        a, b2 = {"h2d": 0, "d2h": 0}, {"h2d": 0, "d2h": 0}
        if need_reads:       # the read arena of the batch: uploaded (and packed) once, used by both groups
            I = hb["ir"]
            L["eseq_i"].reupload(I["q_arena"][:-16]); L["ctx"].synchronize()
            a["h2d"] += len(I["q_arena"]) - 16
        if L["pool"] is not None:
            fut = L["pool"].submit(e2e_group_b, hb, b2, L)
            e2e_group_a(hb, a, L)
            fut.result()
        else:
            e2e_group_a(hb, a, L); e2e_group_b(hb, b2, L)
        io["h2d"], io["d2h"] = a["h2d"] + b2["h2d"], a["d2h"] + b2["d2h"]
        return None

    def timed_e2e(steps):
        """K steps through the host-buffer API.  With two lanes two batches are in flight (lane w takes the steps s = w mod 2, i.e. always batch w: the
        pinned result buffers of a batch are never shared); the region is timed as a whole.  Every step's inputs come from host memory (~1 GB),
        so there is nothing resident to flush between steps."""
        nl = min(len(lanes), NB)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()

        def lane_loop(w):
            for s_ in range(w, steps, nl):
                step_e2e(batches[s_ % NB], lanes[w])
        if nl > 1:
            with ThreadPoolExecutor(nl) as ex:
                list(ex.map(lane_loop, range(nl)))
        else:
            lane_loop(0)
        torch.cuda.synchronize()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

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
    l0 = ctx.launch_count() + (ctxb.launch_count() if two else 0)
    t0 = time.time()
    ms_value = timed(step_value, args.steps, collect)
    sync_all()
    t1 = time.time()
    launches = ctx.launch_count() + (ctxb.launch_count() if two else 0) - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    # ---- end to end through the host-buffer C ABI
    timed_e2e(max(2, args.warmup))
    sync_all()
    ms_e2e = timed_e2e(args.steps)
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
    # kernel classes of one stage run concurrently on side streams, so their event times overlap: among the kernels within 10 % of the
    # longest, the dominant one is the one moving the most algorithmic bytes
    top = max(v["ms"] for v in dp.values())
    dom_name = max((k for k in dp if dp[k]["ms"] >= 0.9 * top), key=lambda k: dp[k]["algo_bytes"])
    dom = kstats[dom_name]
    achieved = dom["algo_bytes"] / (dom["ms"] / 1000.0) / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(dom_name)
        if traffic is not None:        # measured at a smaller batch: per-launch traffic scales with the reads of a launch
            traffic = int(traffic * R / float(tj.get("_captured_at_reads_per_step", R)))
    except Exception:
        pass
    line = {"metric": metric_name(stages), "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD % (R, jobs_per_read), "stages": [s for s in STAGES if s[:3] in stages], "profile": PROFILE,
                       "reads_per_step": R, "aog_jobs_per_step": n_jobs if "a18" in stages else 0, "genome_len": args.genome_len,
                       "l2": "flushed between timed steps (256 MiB fill)",
                       "stage_groups": "a18+a19+a21 and a12+a13 run concurrently on two contexts (streams)" if two else "one context, stages back to back",
                       "e2e_pipeline": "%d batch(es) in flight (one lane of two contexts each); a step's inputs (~1 GB) come from pinned host memory" % min(len(lanes), NB),
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
        try:
            v, desc = cpu_stage_rates(args, batches[0]["jobs0"], batches[0]["segs0"], jobs_per_read, budget_s=4.0)
            line["cpu_baseline"] = {"value": v, "unit": "reads/s", "cores": desc["cores"], "kind": desc["kind"],
                                    "sample": "; ".join("%s: %s" % (k, p["sample"]) for k, p in desc["parts"].items()), "parts": desc["parts"]}
        except Exception as e:          # the GPU numbers above stand on their own: report the failure instead of losing the line
            line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
