#!/usr/bin/env python
"""bench.py -- whole-aligner throughput of the B200-native MapRead path on BASELINE.json's configs[1] workload
(synthetic ONT reads, N50 20 kb, 8 % error, vs a synthetic reference, `lra align -ONT`): reads/s and aligned Gbp/s.

One "step" = one pass of the whole path (lra_b200_map_batch: CreateRC, minimizers, CompareLists against the global index, clustering,
SparseDP x3, LinearExtend, local refinement, AffineOneGapAlign, IndelRefineAlignment, CalculateStatistics, MAPQ; records out) over one batch of
--reads-per-step reads PER GPU.  K steps = K x N x reads-per-step reads (the default K = 8 is 131 k reads on one GPU: configs[1]'s 100 k).

  value     batches resident in HBM when the timed region starts (lra_b200_readset_upload done before), records left in HBM
            (lra_b200_map_resident): kernels only.
  e2e       FASTA-equivalent host buffers in, SAM text out, every step: rank 0 owns the read stream (pinned host memory), H2D, base-balanced
            shards scattered to the ranks over NCCL (N > 1), lra_b200_map_batch on every rank, records gathered to rank 0 over NCCL, D2H,
            lra_b200_format_sam on rank 0.  This is what `lra align reads.fa -p s` does between reading and writing.
  roofline  dominant kernel of the value leg (by summed CUDA-event time): algorithmic bytes (SURVEY 8(d) whole-path formula) / its
            time vs the measured HBM peak.  The path is not HBM-bound; the fraction is reported because the contract asks.
  cpu_baseline / --impl reference   the UNMODIFIED reference binary (oracle/_ref/lra_ref, built by oracle/Makefile from /root/reference with the
            reference's own release flags) run as `lra_ref align -ONT ref.fa sample.fa -t <all cores> -p s` on a bounded sample of the same
            reads against the same index files; its index load time (measured with a one-read run) is subtracted, as the GPU arm's is.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

import synth  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "lra_ref")
MODE = {"ont": "-ONT", "clr": "-CLR", "ccs": "-CCS"}
SEED = {"ont": 2, "clr": 4, "ccs": 11}
PROFILE = {"ont": "ont", "clr": "clr", "ccs": "hifi"}      # tools/synth.py read profiles; ccs = BASELINE configs[2] (HiFi, 15 kb, 0.5 % error)
METRIC = "reads/sec (whole MapRead path: reads in, alignment records out)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default="ont", choices=["ont", "clr", "ccs"])
    ap.add_argument("--reads-per-step", type=int, default=16384, help="reads per step and per GPU")
    ap.add_argument("--genome-len", type=int, default=3_000_000_000)
    ap.add_argument("--contigs", type=int, default=24)
    ap.add_argument("--index-builder", default="auto", choices=["auto", "gpu", "reference"],
                    help="who writes ref.fa.mms / ref.fa.gli: the library's GPU index builder or `lra_ref index` (CPU, ~2 min per Gb)")
    ap.add_argument("--cpu-sample-reads", type=int, default=2048, help="reads per lra_ref run (cpu_baseline and --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sam", action="store_true", help="leave SAM formatting out of the e2e leg")
    ap.add_argument("--e2e-batches", type=int, default=2, help="distinct global batches the e2e leg cycles through (host memory)")
    ap.add_argument("--workdir", default=None)
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


# ---------------------------------------------------------------------------------------------------------------- workload
def workload_name(args):
    reads = {"ont": "BASELINE configs[1]: synthetic ONT reads (log-normal lengths, N50 20 kb, 8% i.i.d. error sub:ins:del 1:1:1)",
             "clr": "BASELINE configs[3]: synthetic CLR reads (log-normal lengths, N50 12 kb, 12% i.i.d. error sub:ins:del 1:1:1)",
             "ccs": "BASELINE configs[2]: synthetic HiFi reads (15 kb +- 2 kb, 0.5% i.i.d. error sub:ins:del 1:1:1)"}[args.preset]
    return ("%s vs a %.3g Gb synthetic reference (%d contigs, uniform ACGT, seed 1234), lra align %s; %d reads per step per GPU"
            % (reads, args.genome_len / 1e9, args.contigs, MODE[args.preset], args.reads_per_step))


def build_workdir(args, builder, device=None):
    """ref.fa + ref.fa.mms + ref.fa.gli in a scratch directory (rank 0 only).  Returns the path."""
    d = args.workdir or tempfile.mkdtemp(prefix="lra_b200_bench_")
    os.makedirs(d, exist_ok=True)
    fa = os.path.join(d, "ref.fa")
    t0 = time.time()
    ref = synth.gen_ref(args.genome_len, args.contigs, 1234)
    synth.write_fasta(fa, ref)
    with open(os.path.join(d, "genome.bin"), "wb") as f:
        for _, seq in ref:
            f.write(seq.tobytes())
    t1 = time.time()
    if builder == "gpu":
        import lra_b200
        lra_b200.build_index_files(fa, ref, args.preset, device or 0)
    else:
        subprocess.run([REF_BIN, "index", MODE[args.preset], fa], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    sys.stderr.write("[bench] reference: %.1f s to generate + write, %.1f s to index (%s)\n" % (t1 - t0, time.time() - t1, builder))
    return d


def load_reference(d, preset):
    import lra_b200
    from lra_b200 import capi
    fa = os.path.join(d, "ref.fa")
    mms = capi.read_mms(fa + ".mms"); gli = capi.read_gli(fa + ".gli")
    names = mms["names"]; hdr = mms["hdr"].astype(np.uint64)
    genome = np.fromfile(os.path.join(d, "genome.bin"), np.uint8)      # the contigs back to back (what Genome::Read keeps in memory)
    assert len(genome) == int(hdr[-1]), (len(genome), int(hdr[-1]))
    opts = lra_b200.map_opts_preset(preset)
    opts.globalK = mms["k"]; opts.smallK = gli["k"]; opts.smallW = gli["w"]; opts.localIndexWindow = gli["window"]
    return dict(genome=genome, hdr=hdr, names=names, mms=mms, gli=gli, opts=opts)


def run_lra_ref(d, preset, reads_fa, cores):
    """wall seconds of `lra_ref align -MODE ref.fa reads.fa -t cores -p s -o ref_arm.sam`"""
    fa = os.path.join(d, "ref.fa")
    out = os.path.join(d, "ref_arm.sam")
    t0 = time.perf_counter()
    subprocess.run([REF_BIN, "align", MODE[preset], fa, reads_fa, "-t", str(cores), "-p", "s", "-o", out], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0


def index_load_seconds(d, preset, one_fa, cores):
    """start-up + index load of the reference binary: a one-read run, second of two (the first warms the page cache)"""
    run_lra_ref(d, preset, one_fa, cores)
    return run_lra_ref(d, preset, one_fa, cores)


def aligned_bases_of_sam(path):
    """sum of qEnd - qStart over primary records (from the CIGAR), for the reference arm's Gbp/s."""
    import re
    tot = 0
    with open(path) as f:
        for line in f:
            if line.startswith("@"):
                continue
            x = line.split("\t", 6)
            if int(x[1]) & 0x904 or x[5] == "*":
                continue
            tot += sum(int(n) for n, op in re.findall(r"(\d+)([=XIM])", x[5]))
    return tot


# ---------------------------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    cores = os.cpu_count() or 1
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    builder = args.index_builder
    if builder == "auto":
        builder = "gpu" if (dev == "cuda" and gpu_builder_available()) else "reference"
    d = build_workdir(args, builder)
    ref = load_reference(d, args.preset)
    gt = torch.from_numpy(ref["genome"]).to(dev)
    S = args.cpu_sample_reads
    one = os.path.join(d, "one.fa")
    vals, walls, loads, gbp = [], [], [], []
    t_load = None
    for i in range(args.warmup + args.steps):
        a, ro, rl, nm = synth.gen_reads_torch(gt, ref["hdr"], ref["names"], S, PROFILE[args.preset], SEED[args.preset] * 1000 + max(0, i - args.warmup), dev)
        fa = os.path.join(d, "sample.fa")
        synth.write_reads_fasta(fa, a, ro, rl, nm)
        if t_load is None:
            synth.write_reads_fasta(one, a, ro[:1], np.minimum(rl[:1], 1000), nm[:1])
            t_load = index_load_seconds(d, args.preset, one, cores)
        t_all = run_lra_ref(d, args.preset, fa, cores)
        if i >= args.warmup:
            dt = max(t_all - t_load, 1e-6)
            vals.append(S / dt); walls.append(t_all); loads.append(t_load)
            gbp.append(aligned_bases_of_sam(os.path.join(d, "ref_arm.sam")) / dt / 1e9)
    dt_mean = float(np.mean([w - l for w, l in zip(walls, loads)]))
    value = S / dt_mean
    sample = ("%d reads per step: `lra_ref align %s ref.fa sample.fa -t %d -p s`, wall %.2f s minus index load %.2f s (one-read run), mean of %d steps"
              % (S, MODE[args.preset], cores, float(np.mean(walls)), float(np.mean(loads)), len(walls)))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "gbp_per_s": float(np.mean(gbp)), "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * dt_mean, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "preset": args.preset, "genome_len": args.genome_len, "contigs": args.contigs, "index_builder": builder,
                       "reads_per_step_timed": S},
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def gpu_builder_available():
    try:
        import lra_b200
        return hasattr(lra_b200, "build_index_files")
    except Exception:
        return False


# ---------------------------------------------------------------------------------------------------------------- GPU arm
def pinned(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly one JSON line: everything else that writes to fd 1 (the NCCL version banner, library chatter) goes to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    args = parse()
    quiet_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return
    import torch
    import torch.distributed as dist
    import lra_b200
    from lra_b200 import shard, capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lra_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    builder = args.index_builder
    if builder == "auto":
        builder = "gpu" if gpu_builder_available() else "reference"

    # ---- reference + index (untimed; the reference arm's index load is excluded too)
    box = [None]
    if rank == 0:
        box[0] = build_workdir(args, builder, local)
    if world > 1:
        dist.broadcast_object_list(box, 0)
    d = box[0]
    ref = load_reference(d, args.preset)
    ctx = lra_b200.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    t0 = time.time()
    mapper = lra_b200.Mapper(ctx, ref["opts"], ref["genome"], ref["hdr"], ref["mms"], ref["gli"])
    sys.stderr.write("[bench] rank %d: index image resident in %.1f s\n" % (rank, time.time() - t0))
    gt = torch.from_numpy(ref["genome"]).to(dev)
    for k in ("genome", "mms", "gli"):      # the host copies are not needed any more (8 ranks x 10 GB otherwise)
        ref[k] = None
    R = args.reads_per_step
    K, W = args.steps, args.warmup

    def chunk(step, r):
        return synth.gen_reads_torch(gt, ref["hdr"], ref["names"], R, PROFILE[args.preset], SEED[args.preset] * 100000 + step * 64 + r, dev)

    # ---- value leg: every rank's shard of every step resident before the clock starts
    shards = []
    for s in range(K):
        a, ro, rl, nm = chunk(s, rank)
        shards.append((mapper.upload(a, ro, rl), len(rl), int(rl.sum())))
    torch.cuda.synchronize()
    torch.cuda.empty_cache()        # the read generator's temporaries go back to the driver before the library sizes its worker arenas

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        mapper.map_resident(shards[i % K][0])
    kstats = {}
    l0 = ctx.launch_count()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    tw0 = time.time()
    e0.record()
    for s in range(K):
        mapper.map_resident(shards[s][0])
        for st in ctx.kernel_stats():
            k = kstats.setdefault(st["name"], dict(ms=0.0, jobs=0, cells=0, algo_bytes=0, launches=0))
            k["ms"] += st["ms"]; k["jobs"] += st["jobs"]; k["cells"] += st["cells"]; k["algo_bytes"] += st["algo_bytes"]; k["launches"] += 1
    e1.record()
    barrier()
    tw1 = time.time()
    launches = ctx.launch_count() - l0
    clocks = sampler.stop(tw0, tw1) if sampler else None
    ms_value = shard.max_over_ranks(e0.elapsed_time(e1), dist if world > 1 else None, dev)
    # aligned bases of the value leg: from the last step's records (download is outside the timed region)
    res_last = mapper.map_download(shards[K - 1][1], shards[K - 1][2])
    ab_last = res_last["aligned_bases"]; bases_last = shards[K - 1][2]
    status_hist = np.bincount(res_last["status"][:shards[K - 1][1]], minlength=4).tolist()
    for h, _, _ in shards:
        mapper.free_readset(h)
    shards = None

    # ---- e2e leg: rank 0 owns the read stream in pinned host memory
    nb = max(1, min(args.e2e_batches, K))
    host = []
    if rank == 0:
        for b in range(nb):
            parts = [chunk(1000 + b, r) for r in range(world)]
            a = np.concatenate([p[0] for p in parts]); rl = np.concatenate([p[2] for p in parts]); nm = sum((p[3] for p in parts), [])
            ro = np.zeros(len(rl), np.uint64); ro[1:] = np.cumsum(rl[:-1].astype(np.uint64))
            host.append(dict(ascii=pinned(a), len=pinned(rl.astype(np.int64)), off=ro, len32=rl, names=nm))
    Rg = R * world
    res_host = None
    if rank == 0:
        nbases = max(int(h["len32"].sum()) for h in host)
        if world == 1:      # (N > 1: the records arrive as device tensors and are copied home shard by shard)
            res_host = [lra_b200.Mapper._result_buffers(Rg, nbases) for _ in range(2)]
            for rh in res_host:
                for k in ("status", "n_aln", "aln_nseg", "aln_seg0", "aln_rank", "records", "cigar"):
                    t = torch.from_numpy(rh[k].view(np.uint8)).pin_memory()
                    rh[k] = t.numpy().view(rh[k].dtype)
        sam_bufs = [np.empty(int(nbases * 1.6) + 1024 * Rg, np.uint8) for _ in range(2)]
    # SAM text of step s is formatted by a host thread (lra_b200_format_sam releases the GIL and uses all cores) while step s + 1 is on the GPU
    from concurrent.futures import ThreadPoolExecutor
    fmt_pool = ThreadPoolExecutor(1)
    fmt_futs = [None, None]
    step_no = [0]
    contig_names = ref["names"]
    h2d = d2h = 0
    sam_bytes = 0
    pin_pool = {}

    def e2e_step(b):
        nonlocal h2d, d2h, sam_bytes
        slot = step_no[0] & 1; step_no[0] += 1
        if fmt_futs[slot] is not None:          # the buffers of this slot are free once their text is out
            sam_bytes = fmt_futs[slot].result(); fmt_futs[slot] = None
        if world == 1:
            hb = host[b]
            res = mapper.map_batch(hb["ascii"].numpy(), hb["off"], hb["len32"], res=res_host[slot])
            h2d = int(hb["ascii"].numel()) + 12 * Rg
            d2h = 4 * Rg * 14 + res["n_records"] * capi.RECORD.itemsize + 4 * res["n_cigar"]
            if not args.no_sam:
                fmt_futs[slot] = fmt_pool.submit(capi.format_sam_into, ref["opts"], dict(res), hb["names_blob"], hb["ascii"].numpy(), hb["off"], hb["len32"], ref["contig_blob"],
                                                 len(contig_names), sam_bufs[slot])
            return res["aligned_bases"]
        # N > 1: H2D on rank 0, scatter shards over NCCL, map, gather records over NCCL, D2H + SAM on rank 0
        if rank == 0:
            hb = host[b]
            a_dev = hb["ascii"].to(dev, non_blocking=True); l_dev = hb["len"].to(dev, non_blocking=True)
            h2d = int(hb["ascii"].numel()) + 8 * Rg
        else:
            a_dev = l_dev = None
        mine, lens, (lo, hi), bounds = shard.scatter_batch(a_dev, l_dev, dist, dev)
        rl = lens.cpu().numpy().astype(np.uint32); n = len(rl)
        ro = np.zeros(n, np.uint64); ro[1:] = np.cumsum(rl[:-1].astype(np.uint64))
        rs = mapper.upload_device(mine, ro, rl)
        mapper.map_resident(rs)
        dres = mapper.download_device(n, int(rl.sum()), dev)
        mapper.free_readset(rs)
        got = shard.gather_parts(dres, dist, dev)
        ab = 0
        if rank == 0:
            hb = host[b]
            d2h = 0
            shards_res = []
            staged = []
            for r, parts in enumerate(got):      # D2H of every shard's records into pinned buffers that live as long as the slot (re-used across steps)
                hp = []
                for i, p in enumerate(parts):
                    key = (slot, r, i); n_el = int(p.numel())
                    buf = pin_pool.get(key)
                    if buf is None or buf.numel() < n_el or buf.dtype != p.dtype:
                        buf = torch.empty(n_el + n_el // 4 + 16, dtype=p.dtype).pin_memory(); pin_pool[key] = buf
                    v = buf[:n_el]; v.copy_(p, non_blocking=True); hp.append(v)
                staged.append(hp)
                d2h += sum(int(p.numel()) * p.element_size() for p in parts)
            torch.cuda.synchronize()
            for hp in staged:
                res = capi.result_from_parts(hp)
                ab += res["aligned_bases"]
                shards_res.append(res)
            if not args.no_sam:
                def fmt_all(shards_res=shards_res, hb=hb, bounds=bounds, buf=sam_bufs[slot]):
                    tot = 0
                    for r, res in enumerate(shards_res):
                        lo_r, hi_r = int(bounds[r]), int(bounds[r + 1])
                        key = ("nb", lo_r, hi_r)
                        if key not in hb:
                            hb[key] = capi.names_blob(hb["names"][lo_r:hi_r])
                        o0 = int(hb["off"][lo_r]) if lo_r < Rg else 0
                        tot += capi.format_sam_into(ref["opts"], res, hb[key], hb["ascii"].numpy()[o0:], hb["off"][lo_r:hi_r] - np.uint64(o0), hb["len32"][lo_r:hi_r],
                                                    ref["contig_blob"], len(contig_names), buf)
                    return tot
                fmt_futs[slot] = fmt_pool.submit(fmt_all)
        return ab

    def e2e_drain():
        nonlocal sam_bytes
        for i in range(2):
            if fmt_futs[i] is not None:
                sam_bytes = fmt_futs[i].result(); fmt_futs[i] = None

    ref["contig_blob"] = capi.names_blob(contig_names)
    if rank == 0:
        for hb in host:
            hb["names_blob"] = capi.names_blob(hb["names"])
    for i in range(W):
        e2e_step(i % nb)
    e2e_drain()
    barrier()
    t0 = time.perf_counter()
    ab_e2e = 0
    for s in range(K):
        ab_e2e += e2e_step(s % nb)
    e2e_drain()
    barrier()
    ms_e2e = shard.max_over_ranks(1000.0 * (time.perf_counter() - t0), dist if world > 1 else None, dev)

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return
    # ---- report
    reads_total = K * Rg
    value = reads_total / (ms_value / 1000.0)
    e2e = reads_total / (ms_e2e / 1000.0)
    peak, peak_src = peaks()
    top = max(kstats.items(), key=lambda kv: kv[1]["ms"])
    tname, tk = top
    ach = tk["algo_bytes"] / 1e9 / (tk["ms"] / 1000.0) if tk["ms"] > 0 else 0.0
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tname in tr.get("kernels", {}):
            kt = tr["kernels"][tname]
            traffic = kt.get("by_preset", {}).get(args.preset, kt)["dram_bytes_per_read"] * R
    except Exception:
        pass
    line = {"metric": METRIC, "value": value, "unit": "reads/s", "gbp_per_s": (ab_last / max(1, R)) * value / 1e9,
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_value / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32+f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "preset": args.preset, "genome_len": args.genome_len, "contigs": args.contigs, "index_builder": builder,
                       "reads_per_step_per_gpu": R, "mean_read_len": bases_last / R, "status_hist_last_step[mapped,unaligned,arena,cap]": status_hist,
                       "l2": "every step maps a different batch; a batch's working set (packed reads + local indexes + worker scratch, several GB) is far larger than the 126 MB L2"},
            "e2e": {"value": e2e, "unit": "reads/s", "gbp_per_s": ab_e2e / (ms_e2e / 1000.0) / 1e9, "ms_per_step": ms_e2e / K, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "sam_bytes_per_step": int(sam_bytes), "includes": "H2D of reads, NCCL scatter/gather (N>1), all kernels, D2H of records" + ("" if args.no_sam else ", SAM formatting on rank 0 (host threads, overlapped with the next step's GPU work)")},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": tname, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms_per_step": tk["ms"] / K, "share_of_step": tk["ms"] / ms_value if world == 1 else None,
                         "note": "algorithmic bytes = SURVEY 8(d) whole-path formula per read; the path is bound by dependent integer/float chain arithmetic, not HBM"},
            "kernels_ms_per_step": {k: round(v["ms"] / K, 3) for k, v in sorted(kstats.items(), key=lambda kv: -kv[1]["ms"])}}
    if world == 1 and not args.no_cpu_baseline and os.path.exists(REF_BIN):
        cores = os.cpu_count() or 1
        S = min(args.cpu_sample_reads, Rg)
        hb = host[0]
        fa = os.path.join(d, "sample.fa"); one = os.path.join(d, "one.fa")
        synth.write_reads_fasta(fa, hb["ascii"].numpy(), hb["off"][:S], hb["len32"][:S], hb["names"][:S])
        synth.write_reads_fasta(one, hb["ascii"].numpy(), hb["off"][:1], np.minimum(hb["len32"][:1], 1000), hb["names"][:1])
        t_load = index_load_seconds(d, args.preset, one, cores)
        t_all = run_lra_ref(d, args.preset, fa, cores)
        dt = max(t_all - t_load, 1e-6)
        line["cpu_baseline"] = {"value": S / dt, "unit": "reads/s", "gbp_per_s": aligned_bases_of_sam(os.path.join(d, "ref_arm.sam")) / dt / 1e9, "cores": cores, "kind": "reference",
                                "sample": "first %d reads of an e2e batch: `lra_ref align %s ref.fa sample.fa -t %d -p s`, wall %.2f s minus index load %.2f s (one-read run)"
                                          % (S, MODE[args.preset], cores, t_all, t_load)}
        # parity at bench scale: the same sample through lra_b200_map_batch + the SAM emitter, record for record against the reference's SAM
        try:
            import re
            o1 = int(hb["off"][S - 1]) + int(hb["len32"][S - 1])
            arr = hb["ascii"].numpy()[:o1]
            off_s = np.ascontiguousarray(hb["off"][:S]).astype(np.uint64); len_s = np.ascontiguousarray(hb["len32"][:S]).astype(np.uint32)
            res_s = mapper.map_batch(arr, off_s, len_s)
            text = lra_b200.format_sam(ref["opts"], res_s, hb["names"][:S], arr, off_s, len_s, contig_names)
            canon = lambda l: re.sub(r"\tRT:i:\d+", "\tRT:i:0", l.rstrip("\n"))
            ours = sorted(canon(l) for l in text.split("\n") if l)
            with open(os.path.join(d, "ref_arm.sam")) as f:
                theirs = sorted(canon(l) for l in f if l.strip() and not l.startswith("@"))
            same = sum(1 for a, b in zip(ours, theirs) if a == b) if len(ours) == len(theirs) else len(set(ours) & set(theirs))
            line["parity_sample"] = {"reads": int(S), "reference_records": len(theirs), "our_records": len(ours), "identical_records": int(same),
                                     "identical": bool(ours == theirs), "what": "SAM records of the cpu_baseline sample, `lra_ref align` vs lra_b200_map_batch + lra_b200_format_sam (RT:i masked, sorted)"}
        except Exception as e:      # the check must not cost the bench line
            line["parity_sample"] = {"error": repr(e)[:200]}
    emit(line)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
