#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 135 -c 135 --csv --log-file gpurun_out/v5_launches.csv python bench.py --leg bigvgan --steps 1 --warmup 1 > gpurun_out/v5_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/v5_launches.csv | head -20
timeout 300 ncu --set full --clock-control none --import-source on -k regex:aa_conv -s 45 -c 1 -f -o gpurun_out/aaconv_c64 python bench.py --leg bigvgan --steps 1 --warmup 1 > gpurun_out/v5_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:aa_conv -s 57 -c 1 -f -o gpurun_out/aaconv_c32 python bench.py --leg bigvgan --steps 1 --warmup 1 > gpurun_out/v5_ncu3.log 2>&1
ls -la gpurun_out | tail -5
