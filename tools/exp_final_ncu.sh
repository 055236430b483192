#!/bin/bash
# launch list + full capture of the default bench pass (nx = 4, T = 1e6, K = 54) of the final build
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-secondary --steps 2 --warmup 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_final4.csv $B > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_filter_reduce|k_mid_scan|k_filter_apply|k_smooth_apply' -c 4 -o /tmp/final4_full -f $B > gpurun_out/final4_ncu.log 2>&1
ncu -i /tmp/final4_full.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/r02_ncu_full_final4_pass.txt
grep -E "^void|duration|dram__bytes|fp64|issue_active|warps_active" gpurun_out/r02_ncu_full_final4_pass.txt | head -40
