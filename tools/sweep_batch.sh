#!/bin/bash
# Batch-size sweep of the default workload: GEMM tile counts of the decode step are multiples of the 148 SMs at
# 37 / 74 M-tiles (364 / 728 segments x 13 channels), see DESIGN.md.
for b in "$@"; do
  echo "# batch=$b"
  timeout 400 python bench.py --batch $b --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1
done
