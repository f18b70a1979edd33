#!/bin/bash
# phase-aligned worker: CTAs per SM x warps per CTA at constant resident warps (24 per SM): smaller CTAs wait for fewer warps at every phase barrier
cd "$(dirname "$0")/.."
out=gpurun_out/${1:-r02ae}_cta_split.log; : > $out
for preset in ont ccs; do
  for cfg in "24 1" "12 2" "8 3" "6 4"; do
    set -- $cfg
    echo "== $preset block_warps=$1 blocks_per_sm=$2" >> $out
    LRA_B200_MAP_BLOCK_WARPS=$1 LRA_B200_MAP_BLOCKS_PER_SM=$2 python tools/map_timing.py --preset $preset --reads 16384 --reps 2 --no-ref 2>&1 | grep -E "rep 1|  map_reads" >> $out
  done
done
cat $out
