#!/bin/bash
mkdir -p gpurun_out
export PTTSPP_UMMA_PAIR_LONGK=1
python bench.py --leg bigvgan --steps 5 --warmup 2 2>/dev/null | tail -c 100
timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu -x -k "bigvgan or vocoder" -s > gpurun_out/v17_models.log 2>&1; echo "exit $?" >> gpurun_out/v17_models.log
grep -i "rms\|passed\|failed\|exit\|Error" gpurun_out/v17_models.log | tail -12
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 166 -c 166 --csv --log-file gpurun_out/v17_l.csv python bench.py --leg bigvgan --steps 1 --warmup 1 > gpurun_out/v17_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/v17_l.csv > gpurun_out/v17_sum.txt; head -16 gpurun_out/v17_sum.txt
