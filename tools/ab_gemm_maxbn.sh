#!/bin/bash
# Same-box A/B of the GEMM tile-width cap (YMT3_GEMM_MAX_BN) on the decode-step GEMMs and the default bench.
for rep in 1 2; do
  for bn in 256 128; do
    echo "# max_bn=$bn rep=$rep"
    export YMT3_GEMM_MAX_BN=$bn
    [ $rep = 1 ] && timeout 120 python tools/bench_ops.py decode728 2>&1 | grep "^gemm"
    timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('value', j['value'], 'ms_per_step', j['ms_per_step'], j['clocks']['sm_mhz'])
"
  done
done
