#!/bin/bash
# Pipeline shape of the absorbed cross-attention kernel: ring stages per CTA x CTAs per SM.
for cfg in "3 1" "1 3" "1 2" "2 1"; do
  set -- $cfg
  export YMT3_XATTN_STAGES=$1 YMT3_XATTN_CTAS=$2
  echo "# stages=$1 ctas_per_sm=$2"
  timeout 200 python tools/bench_ops.py xattn 2>&1 | tail -2
done
for cfg in "3 1" "1 3"; do
  set -- $cfg
  export YMT3_XATTN_STAGES=$1 YMT3_XATTN_CTAS=$2
  echo "# stages=$1 ctas_per_sm=$2 : bench.py"
  timeout 400 python bench.py --steps 3 --no-cpu-baseline 2>&1 | tail -1
done
