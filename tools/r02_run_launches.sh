#!/bin/bash
O=gpurun_out
python tools/profile_step.py yptf_moe_multi 728 6 bf16 2 > $O/r02_profile_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launches_b728_6steps.csv python tools/profile_step.py yptf_moe_multi 728 6 bf16 2 > $O/r02_ncu_launches.log 2>&1
tail -2 $O/r02_ncu_launches.log
