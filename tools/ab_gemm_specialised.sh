#!/bin/bash
# A/B of the compile-time specialised GEMM epilogues (YMT3_GEMM_GENERIC=1 = run-time epilogue kernel for everything).
# usage (GPU box): bash tools/ab_gemm_specialised.sh > gpurun_out/ab_gemm_specialised.txt
for g in 1 0; do
  if [ $g = 1 ]; then export YMT3_GEMM_GENERIC=1; else unset YMT3_GEMM_GENERIC; fi
  echo "# generic=$g : micro-benchmarks"
  timeout 300 python tools/bench_ops.py gemm 2>&1 | tail -14
  echo "# generic=$g : in-graph cost of small kernels (M = 6656)"
  timeout 300 python tools/bench_graph_gap.py 6656 2>&1 | tail -6
  echo "# generic=$g : bench.py (default workload)"
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1
done
