#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py -q -m gpu -x > gpurun_out/umma_all.log 2>&1; echo "exit $?" >> gpurun_out/umma_all.log
tail -4 gpurun_out/umma_all.log
BENCH_QUICK=1 timeout 300 python tools/bench_conv.py > gpurun_out/bench_conv_quick.txt 2>&1
cat gpurun_out/bench_conv_quick.txt
