#!/bin/bash
# phase-aligned worker: which phase points need to be CTA barriers (LRA_B200_MAP_PHASE_MASK, bit i = i-th phase point of a read:
# bits 0-3 stage 1, bits 4-9 chain 0, bits 10-15 chain 1)
cd "$(dirname "$0")/.."
out=gpurun_out/${1:-r02af}_phase_mask.log; : > $out
run() { echo "== $1 mask=$2" >> $out; LRA_B200_MAP_PHASE_MASK=$2 python tools/map_timing.py --preset $1 --reads 16384 --reps 2 --no-ref 2>&1 | grep -E "rep 1|  map_reads" >> $out; }
for m in ffff 0 3cff 30c4 aaaa f fff0; do run ont $m; done
for m in ffff 0 8e3f 861f 3f; do run ccs $m; done
cat $out
