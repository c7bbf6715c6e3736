#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_umma.py -q -m gpu -x -k "32-32" > gpurun_out/umma_c32.log 2>&1; echo "exit $?" >> gpurun_out/umma_c32.log
tail -25 gpurun_out/umma_c32.log
