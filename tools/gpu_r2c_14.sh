#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_serving.py tests/test_abi.py -q -m gpu -x -s > gpurun_out/v14_tests.log 2>&1; echo "exit $?" >> gpurun_out/v14_tests.log
grep -i "rms\|wav\|passed\|failed\|exit" gpurun_out/v14_tests.log | tail -14
python bench.py --leg bigvgan --steps 5 --warmup 2 2>/dev/null | tail -c 100
L=165
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s $L -c $L --csv --log-file gpurun_out/bigvgan_traffic.csv python bench.py --leg bigvgan --steps 1 --warmup 1 > gpurun_out/v14_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/bigvgan_traffic.csv > gpurun_out/v14_sum.txt; head -16 gpurun_out/v14_sum.txt
