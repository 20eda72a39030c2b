#!/bin/bash
# Same-box alternating A/B of the chained decode-step GEMM launches (YMT3_GEMM_CHAIN=1 = chained, default = separate launches).
# usage (on the GPU box): tools/ab_gemm_chain.sh > gpurun_out/r02_ab_gemm_chain.txt
for rep in 1 2; do
  for mode in chain separate; do
    echo "# mode=$mode rep=$rep"
    if [ $mode = chain ]; then export YMT3_GEMM_CHAIN=1; else unset YMT3_GEMM_CHAIN; fi
    timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('value', j['value'], 'ms_per_step', j['ms_per_step'], j['clocks']['sm_mhz'], 'launches', j.get('gpu_launches'))
"
  done
done
