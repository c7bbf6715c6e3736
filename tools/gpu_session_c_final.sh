#!/bin/bash
# round 2, session C: final evidence (tests, smoke, both bench arms, launch lists, ncu captures) -> gpurun_out/
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02c_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r02c_pytest_gpu_final.log
python __graft_entry__.py smoke > gpurun_out/r02c_smoke_final.log 2>&1; tail -1 gpurun_out/r02c_smoke_final.log
# BigVGAN: per-launch duration + DRAM bytes of one forward -> per-family traffic for bench.py's bigvgan.roofline.traffic
L=$(python bench.py --leg bigvgan --steps 1 --warmup 1 2>/dev/null | python -c "import sys,json; print(json.loads(sys.stdin.readlines()[-1])['gpu_launches_per_step'])")
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s $L -c $L --csv --log-file gpurun_out/r02c_launches_bigvgan.csv python bench.py --leg bigvgan --steps 1 --warmup 1 > gpurun_out/r02c_ncu_voc.log 2>&1
python tools/bigvgan_traffic.py gpurun_out/r02c_launches_bigvgan.csv > gpurun_out/r02c_bigvgan_traffic.txt; cat gpurun_out/r02c_bigvgan_traffic.txt
cp profiles/r02_ncu_traffic.json gpurun_out/r02_ncu_traffic.json
python tools/summarize_launches.py gpurun_out/r02c_launches_bigvgan.csv > gpurun_out/r02c_launches_bigvgan_summary.txt; head -20 gpurun_out/r02c_launches_bigvgan_summary.txt
python bench.py --impl reference > gpurun_out/r02c_bench_final_reference.jsonl 2> gpurun_out/r02c_bench_final_reference.err
python bench.py > gpurun_out/r02c_bench_final.jsonl 2> gpurun_out/r02c_bench_final.err; tail -c 300 gpurun_out/r02c_bench_final.err
python tools/gap_profile.py > gpurun_out/r02c_gap_profile_final.txt 2>&1; tail -6 gpurun_out/r02c_gap_profile_final.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02c_launches_acoustic.csv python bench.py --leg acoustic --steps 1 --warmup 1 > gpurun_out/r02c_launch_ac.log 2>&1
python tools/summarize_launches.py gpurun_out/r02c_launches_acoustic.csv > gpurun_out/r02c_launches_acoustic_summary.txt; head -12 gpurun_out/r02c_launches_acoustic_summary.txt
# full captures: weight-resident kernel after the stacked issue, the pair kernel on a long contraction, the activation kernel
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wres_pair_kernel -s 48 -c 2 -f -o gpurun_out/r02c_c32_full python bench.py --leg bigvgan --steps 1 --warmup 1 > gpurun_out/r02c_ncu_c32.log 2>&1
ncu -i gpurun_out/r02c_c32_full.ncu-rep --page raw --csv > gpurun_out/r02c_c32_raw.csv 2>/dev/null; python tools/ncu_summary.py gpurun_out/r02c_c32_raw.csv > gpurun_out/r02c_ncu_c32.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma_pair_kernel -s 20 -c 2 -f -o gpurun_out/r02c_pair_full python bench.py --leg bigvgan --steps 1 --warmup 1 > gpurun_out/r02c_ncu_pair.log 2>&1
ncu -i gpurun_out/r02c_pair_full.ncu-rep --page raw --csv > gpurun_out/r02c_pair_raw.csv 2>/dev/null; python tools/ncu_summary.py gpurun_out/r02c_pair_raw.csv > gpurun_out/r02c_ncu_pair.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:aa_snake_pair -s 100 -c 1 -f -o gpurun_out/r02c_aa_full python bench.py --leg bigvgan --steps 1 --warmup 1 > gpurun_out/r02c_ncu_aa.log 2>&1
ncu -i gpurun_out/r02c_aa_full.ncu-rep --page raw --csv > gpurun_out/r02c_aa_raw.csv 2>/dev/null; python tools/ncu_summary.py gpurun_out/r02c_aa_raw.csv > gpurun_out/r02c_ncu_aa.txt
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -30
