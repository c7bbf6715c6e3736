#!/bin/bash
mkdir -p gpurun_out
for tb in 8 4 2; do echo "AA_TB $tb"; PTTSPP_AA_TB=$tb python tools/bench_aa.py 2>&1 | grep pair; PTTSPP_AA_TB=$tb python bench.py --leg bigvgan --steps 5 --warmup 2 2>/dev/null | tail -c 100; done
PTTSPP_AA_TB=4 timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "aa_snake" 2>&1 | tail -2
PTTSPP_AA_TB=2 timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "aa_snake" 2>&1 | tail -2
