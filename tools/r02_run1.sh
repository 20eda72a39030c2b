#!/bin/bash
# round-2 GPU call 1: new parity tests (verbose), whole GPU tier, bench lines, ncu at the bench shape
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_fulldepth_gpu.py tests/test_moe_gpu.py -m gpu -q -s 2>&1 | tail -60 > $O/r02_run1_newtests.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/r02_run1_pytest_gpu.txt
timeout 900 python bench.py --steps 5 2>$O/r02_run1_bench_default.err | tail -1 > $O/r02_run1_bench_default.json
timeout 600 python bench.py --workload hour --steps 3 --no-cpu-baseline --no-gpu-eager-baseline 2>$O/r02_run1_bench_hour.err | tail -1 > $O/r02_run1_bench_hour_1gpu.json
timeout 600 python bench.py --workload frontend --steps 10 2>/dev/null | tail -1 > $O/r02_run1_bench_frontend.json
python tools/run_decode_attn.py 9464 6 128 > $O/r02_run1_decode_attn_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_attn_kernel -s 1 -c 1 -o $O/r02_decode_attn_9464x6_len128 python tools/run_decode_attn.py 9464 6 128 > $O/r02_run1_ncu_decode.log 2>&1
python tools/profile_step.py yptf_moe_multi 256 2 bf16 1 > $O/r02_run1_profile_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_wide_tc_kernel -c 1 -o $O/r02_attn_wide_tc_b256 python tools/profile_step.py yptf_moe_multi 256 2 bf16 1 > $O/r02_run1_ncu_wide.log 2>&1
tail -5 $O/r02_run1_pytest_gpu.txt
