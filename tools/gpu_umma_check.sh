#!/bin/bash
# parity of every tcgen05 kernel variant, then the per-kernel timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_umma.py -q -m gpu -k "not pair" > gpurun_out/umma_nopair.log 2>&1; echo "exit $?" >> gpurun_out/umma_nopair.log
tail -4 gpurun_out/umma_nopair.log
timeout 600 python -m pytest tests/test_gpu_umma.py -q -m gpu -k "pair" > gpurun_out/umma_pair.log 2>&1; echo "exit $?" >> gpurun_out/umma_pair.log
tail -15 gpurun_out/umma_pair.log
timeout 600 python tools/bench_conv.py > gpurun_out/bench_conv.txt 2>&1
head -40 gpurun_out/bench_conv.txt
