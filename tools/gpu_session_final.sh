set -x
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_final.log
python __graft_entry__.py smoke > gpurun_out/r02_smoke_final.log 2>&1; tail -1 gpurun_out/r02_smoke_final.log
python bench.py --impl reference > gpurun_out/r02_bench_final_reference.jsonl 2> gpurun_out/r02_bench_final_reference.err
python bench.py > gpurun_out/r02_bench_final.jsonl 2> gpurun_out/r02_bench_final.err; tail -c 300 gpurun_out/r02_bench_final.err
python tools/gap_profile.py > gpurun_out/r02_gap_profile_final.txt 2>&1; tail -12 gpurun_out/r02_gap_profile_final.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_acoustic_final.csv python bench.py --leg acoustic --steps 1 --warmup 1 > gpurun_out/r02_launch_ac_final.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_acoustic_final.csv > gpurun_out/r02_launches_acoustic_final_summary.txt; head -14 gpurun_out/r02_launches_acoustic_final_summary.txt
BD_REPS=2 ncu --set full --clock-control none --import-source on -k regex:diffnet_layers -s 3 -c 1 -o gpurun_out/r02_diffnet_final_full -f python tools/bench_diffnet.py > gpurun_out/r02_ncu_diffnet_final.log 2>&1
ncu -i gpurun_out/r02_diffnet_final_full.ncu-rep --page raw --csv > gpurun_out/r02_diffnet_final_full_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02_diffnet_final_full_raw.csv | head -24
