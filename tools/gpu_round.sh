#!/bin/bash
# One GPU visit: parity tests, bench (both arms), ncu launch lists and full captures. Outputs -> gpurun_out/
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.jsonl 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.jsonl
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.jsonl 2>> gpurun_out/bench.err
if [ "$1" != "nonc" ]; then
L=$(python bench.py --leg acoustic --steps 1 --warmup 1 | python -c "import sys,json; print(json.loads(sys.stdin.readlines()[-1])['gpu_launches_per_step'])")
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $L -c $L --csv --log-file gpurun_out/launches_acoustic.csv python bench.py --leg acoustic --steps 1 --warmup 1 > gpurun_out/ncu_ac.log 2>&1
L=$(python bench.py --leg bigvgan --steps 1 --warmup 1 | python -c "import sys,json; print(json.loads(sys.stdin.readlines()[-1])['gpu_launches_per_step'])")
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s $L -c $L --csv --log-file gpurun_out/launches_bigvgan.csv python bench.py --leg bigvgan --steps 1 --warmup 1 > gpurun_out/ncu_voc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma -s 2 -c 2 -f -o gpurun_out/full_umma_pair python tools/prof_umma.py > gpurun_out/ncu_full_umma.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:aa_snake -s 80 -c 2 -f -o gpurun_out/full_aa python bench.py --leg bigvgan --steps 1 --warmup 1 > gpurun_out/ncu_full_aa.log 2>&1
fi
ls -la gpurun_out
