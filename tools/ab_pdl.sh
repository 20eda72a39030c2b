#!/bin/bash
for rep in 1 2; do
  for mode in off on; do
    echo "# pdl=$mode rep=$rep"
    if [ $mode = on ]; then export YMT3_PDL=1; else unset YMT3_PDL; fi
    timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('value', j['value'], 'ms_per_step', j['ms_per_step'], j['clocks']['sm_mhz'])
"
  done
done
