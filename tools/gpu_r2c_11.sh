#!/bin/bash
mkdir -p gpurun_out
python tools/bench_aa_l2.py 2>&1 | tee gpurun_out/v11_aa_l2.txt
