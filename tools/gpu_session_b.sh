set -x
python tools/bench_attention.py > gpurun_out/r02_attn_bench.txt 2>&1
PTTSPP_ATTN_TC=0 python tools/bench_attention.py >> gpurun_out/r02_attn_bench.txt 2>&1
BA_LEGACY=0 python tools/bench_attention.py >> gpurun_out/r02_attn_bench.txt 2>&1
BA_LEGACY=0 PTTSPP_ATTN_TC=0 python tools/bench_attention.py >> gpurun_out/r02_attn_bench.txt 2>&1
BA_B=1 BA_T=52 python tools/bench_attention.py >> gpurun_out/r02_attn_bench.txt 2>&1
BA_B=1 BA_T=52 PTTSPP_ATTN_TC=0 python tools/bench_attention.py >> gpurun_out/r02_attn_bench.txt 2>&1
cat gpurun_out/r02_attn_bench.txt
BA_REPS=2 ncu --set full --clock-control none --import-source on -k regex:relpos_attention_umma -c 1 -s 3 -o gpurun_out/r02_attn_full -f python tools/bench_attention.py > gpurun_out/r02_ncu_attn.log 2>&1
ncu -i gpurun_out/r02_attn_full.ncu-rep --page raw --csv > gpurun_out/r02_attn_full_raw.csv 2>/dev/null
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_b.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_b.log
python bench.py --no-cpu > gpurun_out/r02_bench_b.jsonl 2> gpurun_out/r02_bench_b.err; tail -c 600 gpurun_out/r02_bench_b.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_acoustic_b.csv python bench.py --leg acoustic --steps 1 --warmup 1 > gpurun_out/r02_launch_ac_b.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_acoustic_b.csv > gpurun_out/r02_launches_acoustic_b_summary.txt; head -30 gpurun_out/r02_launches_acoustic_b_summary.txt
