#!/bin/bash
# A/B of the CTA-pair (cta_group::2) GEMM against the single-CTA kernel (YMT3_GEMM_CLUSTER=1).
for cl in 1 2; do
  export YMT3_GEMM_CLUSTER=$cl
  echo "# cluster=$cl : decode-step GEMMs at M = 9464"
  timeout 200 python tools/bench_ops.py decode728 2>&1 | tail -7
  echo "# cluster=$cl : bench.py default"
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'ms_per_step', d['ms_per_step'], d['clocks'])"
done
