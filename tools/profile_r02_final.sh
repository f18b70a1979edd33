#!/bin/bash
# (the .ncu-rep files are ~40 MB each and stay on the box: gpurun_out/ is capped at 64 MiB; the raw page comes back as CSV)
# GPU box, final code of round 2: (1) ncu launch list of a short bench.py run (200 Mb reference so that the run under ncu stays short; same code path as
# the default 3 Gb run), (2) `ncu --set full` of the worker kernel at full load, ONT and CCS reads (8192 reads, 5 Mb reference), raw pages as CSV.
TAG=${1:-r02ah}
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --genome-len 200000000 --contigs 4 --reads-per-step 8192 > gpurun_out/${TAG}_launches.log 2>&1
tail -2 gpurun_out/${TAG}_launches.log | cut -c1-300
export LRA_B200_MAP_ARENA_MB=20
for preset in ont ccs; do
  ncu --set full --import-source on --clock-control none -k regex:map_reads_kernel -c 1 -f -o /tmp/${TAG}_map_${preset} python tools/map_timing.py --preset $preset --reads 8192 --reps 1 --no-ref > gpurun_out/${TAG}_map_${preset}_ncu.log 2>&1
  ncu -i /tmp/${TAG}_map_${preset}.ncu-rep --page raw --csv > gpurun_out/${TAG}_map_${preset}_ncu_raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_map_${preset}.ncu-rep --page details --csv 2>/dev/null | head -400 > gpurun_out/${TAG}_map_${preset}_ncu_details.csv
  tail -2 gpurun_out/${TAG}_map_${preset}_ncu.log | cut -c1-200
done
ls -la gpurun_out/${TAG}_*
