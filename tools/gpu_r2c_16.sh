#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_umma_as_kernel -s 7 -c 1 -f -o gpurun_out/as64_full python bench.py --leg bigvgan --steps 1 --warmup 1 > gpurun_out/v16_ncu.log 2>&1
ncu -i gpurun_out/as64_full.ncu-rep --page raw --csv > gpurun_out/as64_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/as64_raw.csv
