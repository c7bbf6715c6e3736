#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "aa_snake" > gpurun_out/v3_tests.log 2>&1; echo "exit $?" >> gpurun_out/v3_tests.log
tail -3 gpurun_out/v3_tests.log
python tools/bench_aa.py > gpurun_out/aa_v3.txt 2>&1; grep pair gpurun_out/aa_v3.txt
python bench.py --leg bigvgan --steps 5 --warmup 2 > gpurun_out/v3_bigvgan.jsonl 2> gpurun_out/v3_bigvgan.err; tail -c 300 gpurun_out/v3_bigvgan.jsonl
