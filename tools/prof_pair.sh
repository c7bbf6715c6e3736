#!/bin/bash
mkdir -p gpurun_out
BENCH_QUICK=1 timeout 300 python tools/bench_conv.py > gpurun_out/bench_conv_quick.txt 2>&1
cat gpurun_out/bench_conv_quick.txt
PTTSPP_UMMA_PAIR=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pair -s 2 -c 2 -f -o gpurun_out/full_pair python tools/prof_umma.py > gpurun_out/ncu_full_pair.log 2>&1
tail -3 gpurun_out/ncu_full_pair.log
