#!/bin/bash
# End-of-round evidence run on one B200 (no profiler): whole GPU test tier, smoke, every bench workload, reference arm.
# usage (GPU box): bash tools/final_round_run.sh [TAG]   (writes gpurun_out/<TAG>_*.{txt,json}; default TAG = final)
O=gpurun_out
T=${1:-final}
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/${T}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 > $O/${T}_smoke.txt
timeout 900 python bench.py 2>$O/${T}_bench_default.err | tail -1 > $O/${T}_bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > $O/${T}_bench_reference_arm.json
timeout 300 python bench.py --workload frontend --steps 10 2>/dev/null | tail -1 > $O/${T}_bench_frontend.json
timeout 400 python bench.py --workload t5_small --steps 3 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | tail -1 > $O/${T}_bench_t5_small.json
timeout 400 python bench.py --workload yptf --steps 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/${T}_bench_yptf_b64.json
timeout 400 python bench.py --workload yptf --batch 256 --steps 3 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | tail -1 > $O/${T}_bench_yptf_b256.json
timeout 400 python bench.py --batch 64 --steps 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/${T}_bench_default_b64.json
timeout 400 python bench.py --workload hour --steps 3 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | tail -1 > $O/${T}_bench_hour_1gpu.json
tail -2 $O/${T}_pytest_gpu.txt; tail -2 $O/${T}_smoke.txt
