#!/bin/bash
# GPU box: ncu captures of the mapper worker kernel on a small ONT batch.  Pass 1: counters + PC sampling (no SASS patching); pass 2: SourceCounters
# (per-instruction executed counts).  The source pages are reduced to (address, SASS, samples, inst executed, thread inst executed) per function.
TAG=${1:-r02}
READS=${2:-1024}
cd "$(dirname "$0")/.."
export LRA_B200_MAP_ARENA_MB=${ARENA_MB:-20}
SEC1="--section SpeedOfLight --section SchedulerStats --section WarpStateStats --section Occupancy --section LaunchStats --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section InstructionStats"
ncu $SEC1 --clock-control none --import-source on -k regex:map_reads_kernel -c 1 -f -o /tmp/mapprof_$TAG python tools/map_timing.py --preset ont --reads $READS --reps 1 --no-ref > gpurun_out/mapprof_$TAG.log 2>&1
ncu -i /tmp/mapprof_$TAG.ncu-rep --page raw --csv > gpurun_out/mapprof_${TAG}_raw.csv 2> /dev/null
ncu -i /tmp/mapprof_$TAG.ncu-rep --page source --csv > /tmp/mapprof_${TAG}_src.csv 2> /dev/null
ncu --section SourceCounters --clock-control none --import-source on -k regex:map_reads_kernel -c 1 -f -o /tmp/mapprof2_$TAG python tools/map_timing.py --preset ont --reads $READS --reps 1 --no-ref > gpurun_out/mapprof2_$TAG.log 2>&1
ncu -i /tmp/mapprof2_$TAG.ncu-rep --page source --csv > /tmp/mapprof2_${TAG}_src.csv 2> /dev/null
python - <<PY
import csv
for src, dst in (("/tmp/mapprof_${TAG}_src.csv", "gpurun_out/mapprof_${TAG}_src.csv"), ("/tmp/mapprof2_${TAG}_src.csv", "gpurun_out/mapprof2_${TAG}_src.csv")):
    try:
        rows = list(csv.reader(open(src)))
    except Exception as e:
        print(src, e); continue
    keep = ("Address", "Source", "Warp Stall Sampling (All Samples)", "# Samples", "Instructions Executed", "Thread Instructions Executed")
    out = []; cols = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            out.append(r[:2]); cols = None; continue
        if r and r[0] == "Address":
            cols = [i for i, c in enumerate(r) if c in keep]
        if cols is not None:
            out.append([r[i] if i < len(r) else "" for i in cols])
    with open(dst, "w", newline="") as f:
        csv.writer(f).writerows(out)
PY
ls -la /tmp/mapprof*_$TAG.ncu-rep gpurun_out/mapprof*_${TAG}_*.csv
tail -4 gpurun_out/mapprof_$TAG.log gpurun_out/mapprof2_$TAG.log
