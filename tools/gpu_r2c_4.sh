#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_umma.py -q -m gpu -x -k "aa_conv" -s > gpurun_out/v4_tests.log 2>&1; echo "exit $?" >> gpurun_out/v4_tests.log
tail -20 gpurun_out/v4_tests.log
if grep -q "exit 0" gpurun_out/v4_tests.log; then
python bench.py --leg bigvgan --steps 5 --warmup 2 > gpurun_out/v4_bigvgan.jsonl 2> gpurun_out/v4_bigvgan.err; tail -c 300 gpurun_out/v4_bigvgan.jsonl
PTTSPP_AA_FUSE=0 python bench.py --leg bigvgan --steps 5 --warmup 2 > gpurun_out/v4_bigvgan_nofuse.jsonl 2>> gpurun_out/v4_bigvgan.err; tail -c 300 gpurun_out/v4_bigvgan_nofuse.jsonl
timeout 600 python -m pytest tests/test_gpu_models.py -q -m gpu -x -k "vocoder or bigvgan or wav" -s > gpurun_out/v4_models.log 2>&1; echo "exit $?" >> gpurun_out/v4_models.log
tail -8 gpurun_out/v4_models.log
fi
