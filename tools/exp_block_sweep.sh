#!/bin/bash
# phase-aligned worker: warps per CTA sweep (one CTA per SM)
cd "$(dirname "$0")/.."
out=gpurun_out/${1:-r02f}_block_sweep.log; : > $out
for bw in ${2:-4 8 16}; do
  echo "== block_warps=$bw" >> $out
  LRA_B200_MAP_BLOCK_WARPS=$bw LRA_B200_MAP_ARENA_MB=24 python tools/map_timing.py --preset ont --reads 8192 --reps 2 --no-ref 2>&1 | grep -E "rep 1|map_reads" >> $out
done
cat $out
