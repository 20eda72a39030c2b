#!/bin/bash
# A/B of the 128 x 256 GEMM tiles + cost-model tile choice (YMT3_GEMM_MAX_BN=128 = previous behaviour up to the choice rule).
for cap in 128 256; do
  export YMT3_GEMM_MAX_BN=$cap
  echo "# max_bn=$cap : micro-benchmarks"
  timeout 300 python tools/bench_ops.py gemm 2>&1 | tail -14
  echo "# max_bn=$cap : in-graph cost of small kernels (M = 6656)"
  timeout 300 python tools/bench_graph_gap.py 6656 2>&1 | tail -6
  for b in 512 728; do
    echo "# max_bn=$cap : bench.py batch $b"
    timeout 300 python bench.py --batch $b --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1
  done
done
