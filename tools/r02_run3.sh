#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_logmel_gpu.py -m gpu -q -x 2>&1 | tail -15 > $O/r02_run4_logmel_tests.txt
timeout 300 python bench.py --workload frontend --steps 10 --no-cpu-baseline 2>/dev/null | tail -1 > $O/r02_run4_bench_frontend.json
timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | tail -1 > $O/r02_run4_bench_default.json
timeout 900 python -m pytest tests/test_fulldepth_gpu.py tests/test_moe_gpu.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -80 > $O/r02_run4_newtests.txt
tail -3 $O/r02_run4_logmel_tests.txt; tail -3 $O/r02_run4_newtests.txt
