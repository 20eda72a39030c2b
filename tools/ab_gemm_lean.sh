#!/bin/bash
# Same-box A/B of the lean bf16 GEMM epilogue (YMT3_GEMM_NO_LEAN=1 = general chunk path) on micro-benchmarks and the bench.
for rep in 1 2; do
  for mode in lean general; do
    echo "# epilogue=$mode rep=$rep"
    if [ $mode = general ]; then export YMT3_GEMM_NO_LEAN=1; else unset YMT3_GEMM_NO_LEAN; fi
    if [ $rep = 1 ]; then
      timeout 120 python tools/bench_ops.py decode728 2>&1 | grep "^gemm"
      timeout 200 python tools/bench_ops.py gemm 2>&1 | grep "^gemm"
    fi
    timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('value', j['value'], 'ms_per_step', j['ms_per_step'], j['clocks']['sm_mhz'])
"
  done
done
