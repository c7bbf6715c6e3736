#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py -q -m gpu -x -k "plain and 64-64-11" > gpurun_out/v19_tests.log 2>&1; echo "exit $?" >> gpurun_out/v19_tests.log
tail -4 gpurun_out/v19_tests.log
python bench.py --leg bigvgan --steps 5 --warmup 2 2>/dev/null | tail -c 100
PTTSPP_UMMA_AS64_NSUB1=1 python bench.py --leg bigvgan --steps 5 --warmup 2 2>/dev/null | tail -c 100
timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu -x -k "bigvgan or vocoder" > gpurun_out/v19_models.log 2>&1; echo "exit $?" >> gpurun_out/v19_models.log
tail -3 gpurun_out/v19_models.log
