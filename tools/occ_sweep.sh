#!/bin/bash
# occupancy sweep of the mapper worker kernel: min blocks/SM (register cap) x resident warps/SM, fixed 24 MB worker arenas
cd "$(dirname "$0")/.."
out=gpurun_out/r02b_occ_sweep.log; : > $out
for cfg in "4 16" "6 24" "8 32" "8 24" "6 16"; do
  set -- $cfg
  lib=lra_b200/liblra_b200_mb$1.so; [ $1 = 4 ] && lib=lra_b200/liblra_b200.so
  echo "== min_blocks=$1 warps_per_sm=$2" >> $out
  LRA_B200_LIB=$PWD/$lib LRA_B200_MAP_WARPS_PER_SM=$2 LRA_B200_MAP_ARENA_MB=24 python tools/map_timing.py --preset ont --reads 8192 --reps 2 --no-ref 2>&1 | grep -E "rep|map_reads|status" >> $out
done
cat $out
