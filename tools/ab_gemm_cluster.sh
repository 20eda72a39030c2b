#!/bin/bash
# A/B of the 2-CTA cluster + TMA multicast of the W tile (YMT3_GEMM_CLUSTER=1: previous single-CTA kernel).
O=gpurun_out
for cl in 1 2; do
  export YMT3_GEMM_CLUSTER=$cl
  echo "# cluster=$cl : micro-benchmarks"
  timeout 300 python tools/bench_ops.py gemm 2>&1 | tail -16
  echo "# cluster=$cl : bench.py default"
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'ms_per_step', d['ms_per_step'], d['clocks'])"
done
