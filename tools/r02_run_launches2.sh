#!/bin/bash
# Launch list of the final build at the bench shape (B = 728, 6 decode steps, two runs): per-kernel durations under ncu.
O=gpurun_out
python tools/profile_step.py yptf_moe_multi 728 6 bf16 2 > $O/r02_profile_plain2.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launches2_b728_6steps.csv python tools/profile_step.py yptf_moe_multi 728 6 bf16 2 > $O/r02_ncu_launches2.log 2>&1
tail -2 $O/r02_ncu_launches2.log
