#!/bin/bash
# A/B of the TMA-store GEMM epilogue (YMT3_GEMM_DIRECT_STORE=1 = the previous 16-byte st.global epilogue).
# usage (GPU box): bash tools/ab_gemm_tma_store.sh > gpurun_out/ab_gemm_tma_store.txt
for d in 1 0; do
  if [ $d = 1 ]; then export YMT3_GEMM_DIRECT_STORE=1; else unset YMT3_GEMM_DIRECT_STORE; fi
  echo "# direct_store=$d : micro-benchmarks"
  timeout 300 python tools/bench_ops.py gemm 2>&1 | tail -14
  echo "# direct_store=$d : bench.py (default workload)"
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1
done
