#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py -q -m gpu -x -k "plain or aa_conv" > gpurun_out/v15_tests.log 2>&1; echo "exit $?" >> gpurun_out/v15_tests.log
tail -4 gpurun_out/v15_tests.log
python bench.py --leg bigvgan --steps 5 --warmup 2 2>/dev/null | tail -c 100
PTTSPP_UMMA_NO_AS64=1 python bench.py --leg bigvgan --steps 5 --warmup 2 2>/dev/null | tail -c 100
timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu -x -k "bigvgan or vocoder" > gpurun_out/v15_models.log 2>&1; echo "exit $?" >> gpurun_out/v15_models.log
tail -3 gpurun_out/v15_models.log
