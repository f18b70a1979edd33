#!/bin/bash
# GPU box: (1) ncu launch list of a short bench.py run (200 Mb reference so that the run under ncu stays short; same code path as the default
# 3 Gb run), (2) counters + DRAM bytes of the worker kernel at full load (8192 ONT reads, 5 Mb reference).
TAG=${1:-r02q}
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --genome-len 200000000 --contigs 4 --reads-per-step 8192 > gpurun_out/${TAG}_launches.log 2>&1
tail -2 gpurun_out/${TAG}_launches.log | cut -c1-400
export LRA_B200_MAP_ARENA_MB=20
ncu --section SpeedOfLight --section SchedulerStats --section WarpStateStats --section Occupancy --section LaunchStats --section MemoryWorkloadAnalysis --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:map_reads_kernel -c 1 -f -o /tmp/map_$TAG python tools/map_timing.py --preset ont --reads 8192 --reps 1 --no-ref > gpurun_out/${TAG}_map_ncu.log 2>&1
ncu -i /tmp/map_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_map_ncu_raw.csv 2>/dev/null
tail -2 gpurun_out/${TAG}_map_ncu.log
