#!/bin/bash
# A/B of the speculative first-block fetch in decode_attn_kernel (YMT3_DECODE_ATTN_NO_SPEC=1 = off).
for ns in 1 0; do
  if [ $ns = 1 ]; then export YMT3_DECODE_ATTN_NO_SPEC=1; else unset YMT3_DECODE_ATTN_NO_SPEC; fi
  echo "# no_spec=$ns : kernel alone, cache length sweep"
  timeout 300 python tools/bench_ops.py decode_attn 2>&1 | tail -8
  echo "# no_spec=$ns : bench.py (default workload)"
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1
done
