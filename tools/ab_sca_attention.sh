#!/bin/bash
# A/B of the wide-head tensor-core spectral cross-attention (YMT3_NO_TC_ATTN=1 also disables the dk-16 kernel, so the
# comparison below is the full bench with and without both, plus the previous record of the dk-16 kernel alone).
for n in 1 0; do
  if [ $n = 1 ]; then export YMT3_NO_TC_ATTN=1; else unset YMT3_NO_TC_ATTN; fi
  echo "# no_tc_attn=$n : encoder phase timing"
  timeout 300 python tools/time_phases.py yptf_moe_multi 512 2>&1 | head -2
done
echo "# bench.py default"
timeout 400 python bench.py --steps 3 --no-cpu-baseline 2>&1 | tail -1
