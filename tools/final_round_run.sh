#!/bin/bash
# End-of-session evidence run on one B200 (no profiler): tests, smoke, cross-attention stress, every bench workload, reference arm.
# usage (GPU box): bash tools/final_round_run.sh   (writes gpurun_out/final_*.{txt,json})
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/final_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 > $O/final_smoke.txt
timeout 300 python tools/stress_cross_absorbed.py 9464 3328 5000 4000 > $O/final_stress_xattn.txt 2>&1
timeout 600 python bench.py 2>$O/final_bench_default.err | tail -1 > $O/final_bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $O/final_bench_reference_arm.json
timeout 300 python bench.py --workload frontend --steps 10 2>/dev/null | tail -1 > $O/final_bench_frontend.json
timeout 300 python bench.py --workload t5_small --steps 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/final_bench_t5_small.json
timeout 300 python bench.py --workload yptf --steps 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/final_bench_yptf_b64.json
timeout 300 python bench.py --workload yptf --batch 256 --steps 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/final_bench_yptf_b256.json
timeout 300 python bench.py --batch 512 --steps 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/final_bench_default_b512.json
tail -2 $O/final_pytest_gpu.txt; tail -2 $O/final_smoke.txt; cat $O/final_stress_xattn.txt
