#!/bin/bash
# Development aid (GPU box): sweep the long-group threshold of the IndelRefine DP dispatch and the batch size.
for n in 4096 16384; do
  for lr in 1 6000 12000 24576 1000000000; do
    echo "== n=$n LRA_B200_IR_LONG_ROWS=$lr"
    LRA_B200_IR_LONG_ROWS=$lr python tools/ir_timing.py ont $n 2>&1 | grep -E "iter 2|ir_dp"
  done
done
