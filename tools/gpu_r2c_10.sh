#!/bin/bash
mkdir -p gpurun_out
python bench.py --leg bigvgan --steps 5 --warmup 2 2>/dev/null | tail -c 100
PTTSPP_UMMA_NARROW_NACC=4 python bench.py --leg bigvgan --steps 5 --warmup 2 2>/dev/null | tail -c 100
timeout 1500 python -m pytest tests -q -m gpu -x -s > gpurun_out/v10_pytest.log 2>&1; echo "exit $?" >> gpurun_out/v10_pytest.log
tail -15 gpurun_out/v10_pytest.log
python bench.py --leg acoustic --steps 3 --warmup 1 2>/dev/null | tail -c 200
