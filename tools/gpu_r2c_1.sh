#!/bin/bash
# round-2 session C, visit 1: AA-Snake baseline timings + full ncu captures (warp states) of the AA pair kernel and the
# weight-resident 32-channel conv
mkdir -p gpurun_out
python tools/bench_aa.py > gpurun_out/aa_base.txt 2>&1
tail -8 gpurun_out/aa_base.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:aa_snake_pair -s 25 -c 1 -f -o gpurun_out/aa_full python tools/bench_aa.py > gpurun_out/ncu_aa.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:c32 -s 24 -c 2 -f -o gpurun_out/c32_full python bench.py --leg bigvgan --steps 1 --warmup 1 > gpurun_out/ncu_c32.log 2>&1
ls -la gpurun_out
