#!/bin/bash
# GPU box: ncu launch list and one `--set full` capture per hot kernel of a short bench run (classes serialised so that kernels do
# not overlap under the profiler).  Usage: tools/profile.sh <tag>   -> gpurun_out/launches_<tag>.csv, prof_<tag>.ncu-rep, prof_<tag>_raw.csv
TAG=${1:-r01}
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --reads-per-step 4096"
export LRA_B200_SERIAL=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > gpurun_out/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ir_dp_pipe|ir_dp_warp|ir_band|lidx_window|lref_task_literal|stats_warp|aog_thread_kernel|aog_warp_literal|lref_prep|ir_group" \
    -c 36 -f -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --reads-per-step 4096 > gpurun_out/prof_$TAG.log 2>&1
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2> /dev/null
ls -la gpurun_out/prof_$TAG.ncu-rep gpurun_out/launches_$TAG.csv gpurun_out/prof_${TAG}_raw.csv
