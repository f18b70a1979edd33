#!/bin/bash
# GPU box: ncu launch list and one `--set full` capture per hot kernel of a short bench run (classes serialised so that kernels do
# not overlap under the profiler).  Usage: tools/profile.sh <tag>   -> gpurun_out/launches_<tag>.csv, prof_<tag>.ncu-rep, prof_<tag>_raw.csv
TAG=${1:-r01}
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --reads-per-step 4096 --serial-stages"
export LRA_B200_SERIAL=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv python tools/stage_bench.py $ARGS > gpurun_out/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ir_dp_pipe|ir_band|lidx_window|lref_task_literal|lref_chain_unit|stats_warp|aog_thread_kernel|aog_warp_literal|lref_prep|ir_group" \
    -c 24 -f -o /tmp/prof_$TAG python tools/stage_bench.py --steps 1 --warmup 1 --no-cpu-baseline --reads-per-step 2048 --serial-stages > gpurun_out/prof_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2> /dev/null
# source-level hot spots of the two heaviest kernels (the report itself is too large to bring back: gpurun_out is capped at 64 MiB)
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv -k regex:"lidx_window" > gpurun_out/prof_${TAG}_src_lidx.csv 2> /dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv -k regex:"ir_dp_pipe" > gpurun_out/prof_${TAG}_src_irpipe.csv 2> /dev/null
ls -la /tmp/prof_$TAG.ncu-rep gpurun_out/launches_$TAG.csv gpurun_out/prof_${TAG}_raw.csv gpurun_out/prof_${TAG}_src_*.csv
