#!/bin/bash
# visit 2: AA kernel (direct fast sine + input prefetch) and N-stacked weight-resident conv: parity + timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_umma.py -q -m gpu -x -k "aa_snake or 32-32 or 64-64 or wres or c32" > gpurun_out/v2_tests.log 2>&1; echo "exit $?" >> gpurun_out/v2_tests.log
tail -5 gpurun_out/v2_tests.log
python tools/bench_aa.py > gpurun_out/aa_v2.txt 2>&1; grep pair gpurun_out/aa_v2.txt
python bench.py --leg bigvgan --steps 5 --warmup 2 > gpurun_out/v2_bigvgan.jsonl 2> gpurun_out/v2_bigvgan.err; tail -c 1500 gpurun_out/v2_bigvgan.jsonl
timeout 600 python -m pytest tests/test_gpu_models.py -q -m gpu -x -k "vocoder or bigvgan or wav" -s > gpurun_out/v2_models.log 2>&1; echo "exit $?" >> gpurun_out/v2_models.log
tail -8 gpurun_out/v2_models.log
