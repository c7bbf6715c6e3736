#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_umma.py -q -m gpu -x -k "aa_conv" > gpurun_out/v6_tests.log 2>&1; echo "exit $?" >> gpurun_out/v6_tests.log
tail -3 gpurun_out/v6_tests.log
python bench.py --leg bigvgan --steps 5 --warmup 2 > gpurun_out/v6_bigvgan.jsonl 2> gpurun_out/v6_bigvgan.err; tail -c 300 gpurun_out/v6_bigvgan.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 135 -c 135 --csv --log-file gpurun_out/v6_launches.csv python bench.py --leg bigvgan --steps 1 --warmup 1 > gpurun_out/v6_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/v6_launches.csv | head -8
