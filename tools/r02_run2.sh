#!/bin/bash
# round-2 GPU call: log-mel v3 (packed f32x2) parity + speed, bf16 diagnostics after the fp32 residual stream change
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_logmel_gpu.py -m gpu -q 2>&1 | tail -15 > $O/r02_run2_logmel_tests.txt
timeout 300 python bench.py --workload frontend --steps 10 --no-cpu-baseline 2>/dev/null | tail -1 > $O/r02_run2_bench_frontend.json
timeout 300 python tools/diag_bf16_logits.py > $O/r02_run2_diag_bf16.txt 2>&1
timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | tail -1 > $O/r02_run2_bench_default.json
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r02_run2_pytest_gpu.txt
tail -4 $O/r02_run2_pytest_gpu.txt; cat $O/r02_run2_logmel_tests.txt | tail -3
