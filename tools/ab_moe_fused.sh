#!/bin/bash
# Same-box A/B of the fused MoE expert kernel (YMT3_NO_MOE_FUSED=1 = two grouped GEMMs): encoder phase + default bench.
for rep in 1 2; do
  for mode in fused grouped; do
    echo "# experts=$mode rep=$rep"
    if [ $mode = grouped ]; then export YMT3_NO_MOE_FUSED=1; else unset YMT3_NO_MOE_FUSED; fi
    [ $rep = 1 ] && timeout 200 python tools/time_phases.py yptf_moe_multi 728 2>&1 | sed -n 2p
    timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('value', j['value'], 'ms_per_step', j['ms_per_step'], j['clocks']['sm_mhz'])
"
  done
done
