#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_umma.py -q -m gpu -x -k "aa_conv" > gpurun_out/v7_tests.log 2>&1; echo "exit $?" >> gpurun_out/v7_tests.log
tail -3 gpurun_out/v7_tests.log
for cfg in 0 1; do
PTTSPP_AAF_CFG=$cfg python bench.py --leg bigvgan --steps 5 --warmup 2 2>> gpurun_out/v7_bigvgan.err | tail -c 200
PTTSPP_AAF_CFG=$cfg timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 135 -c 135 --csv --log-file gpurun_out/v7_launches_$cfg.csv python bench.py --leg bigvgan --steps 1 --warmup 1 > gpurun_out/v7_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/v7_launches_$cfg.csv > gpurun_out/v7_sum_$cfg.txt; head -5 gpurun_out/v7_sum_$cfg.txt
done
