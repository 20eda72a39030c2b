#!/bin/bash
# A/B of the split-length decode attention (warps per (sequence, head)) on the small-batch single-channel workloads.
# usage (GPU box): bash tools/ab_decode_attn_warps.sh > gpurun_out/ab_decode_attn_warps.txt
for wl_b in "yptf 64" "yptf 256" "t5_small 256" "t5_small 512"; do
  set -- $wl_b
  for w in 1 0; do
    echo "# workload=$1 batch=$2 YMT3_DECODE_ATTN_WARPS=$w (0 = automatic)"
    YMT3_DECODE_ATTN_WARPS=$w timeout 300 python bench.py --workload $1 --batch $2 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1
  done
done
