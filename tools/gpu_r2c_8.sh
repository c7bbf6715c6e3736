#!/bin/bash
mkdir -p gpurun_out
for cfg in 0 1; do for dbg in 0 1 2 3; do
echo "cfg $cfg dbg $dbg"
PTTSPP_AAF_CFG=$cfg PTTSPP_AAF_DBG=$dbg timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 135 -c 135 --csv --log-file gpurun_out/v8_l.csv python bench.py --leg bigvgan --steps 1 --warmup 1 > gpurun_out/v8_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/v8_l.csv > gpurun_out/v8_sum_${cfg}_${dbg}.txt; grep aa_conv gpurun_out/v8_sum_${cfg}_${dbg}.txt
done; done
