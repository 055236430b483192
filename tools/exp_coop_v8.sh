#!/bin/bash
mkdir -p gpurun_out
tools/exp_coop_variants.sh v8
export PSQRT_LIB=$PWD/devlibs/libpsqrt_v8.so
for m in 4 6; do PSQRT_COOP=$m timeout 300 python tools/check_coop.py 2>&1 | grep -E "WORST|rror" | tail -2; done
B="python bench.py --no-cpu-baseline --no-secondary --steps 1 --warmup 1 --nx 8 --ny 4"
timeout 900 ncu --set full --clock-control none -k regex:'k_coopr_' -c 3 -o /tmp/coop_g4_full -f $B > gpurun_out/coop_g4_ncu.log 2>&1
ncu -i /tmp/coop_g4_full.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/coop_g4_full_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/coop_g4_launches_n8.csv $B > /dev/null 2>&1
grep -E "^void|duration|fp64|issue_active|warps_active|inst_executed.sum|stall" gpurun_out/coop_g4_full_summary.txt | head -60
