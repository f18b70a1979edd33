#!/bin/bash
# GPU box: warp-state / scheduler counters of the worker kernel at full load (8192 ONT reads) -> gpurun_out/<tag>_mapfull_raw.csv, and the
# PC-sampling source page reduced to (address, samples, inst executed) -> gpurun_out/<tag>_mapfull_src.csv
TAG=${1:-r02g}
cd "$(dirname "$0")/.."
export LRA_B200_MAP_ARENA_MB=${ARENA_MB:-20}
ncu --section SpeedOfLight --section SchedulerStats --section WarpStateStats --section Occupancy --section SourceCounters --clock-control none --import-source on -k regex:map_reads_kernel -c 1 -f -o /tmp/mapfull_$TAG python tools/map_timing.py --preset ont --reads ${2:-8192} --reps 1 --no-ref > gpurun_out/${TAG}_mapfull.log 2>&1
ncu -i /tmp/mapfull_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_mapfull_raw.csv 2>/dev/null
ncu -i /tmp/mapfull_$TAG.ncu-rep --page source --csv > /tmp/mapfull_${TAG}_src.csv 2>/dev/null
python - <<PY
import csv
rows = list(csv.reader(open("/tmp/mapfull_${TAG}_src.csv")))
keep = ("Address", "Source", "Warp Stall Sampling (All Samples)", "# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_no_inst", "stall_long_sb", "stall_wait", "stall_barrier")
out = []; cols = None
for r in rows:
    if r and r[0] == "Kernel Name":
        out.append(r[:2]); cols = None; continue
    if r and r[0] == "Address":
        cols = [i for i, c in enumerate(r) if c in keep]
    if cols is not None:
        out.append([r[i] if i < len(r) else "" for i in cols])
csv.writer(open("gpurun_out/${TAG}_mapfull_src.csv", "w", newline="")).writerows(out)
PY
tail -2 gpurun_out/${TAG}_mapfull.log
