#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py -q -m gpu -x -k "impl4" -s > gpurun_out/v18_tests.log 2>&1; echo "exit $?" >> gpurun_out/v18_tests.log
grep -i "impl4\|passed\|failed\|exit\|Error" gpurun_out/v18_tests.log | tail -8
python bench.py --leg bigvgan --steps 5 --warmup 2 2>/dev/null | tail -c 100
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_serving.py -q -m gpu -x > gpurun_out/v18_models.log 2>&1; echo "exit $?" >> gpurun_out/v18_models.log
tail -3 gpurun_out/v18_models.log
