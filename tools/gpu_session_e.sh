python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_e.jsonl 2> gpurun_out/r02_bench_e.err
PTTSPP_DIFFNET_TAIL=0 python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_e_notail.jsonl 2>/dev/null
python - <<'PY'
import json
for f in ('r02_bench_e','r02_bench_e_notail'):
    d=json.loads(open(f'gpurun_out/{f}.jsonl').read().strip().splitlines()[-1])
    print(f, round(d['ms_per_step'],2), d['gpu_launches'], [ (k['kernel'][:12],round(k['ms'],2),k['calls']) for k in d['kernel_families'] if k['ms']>0])
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_acoustic_e.csv python bench.py --leg acoustic --steps 1 --warmup 1 > gpurun_out/r02_launch_ac_e.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_acoustic_e.csv > gpurun_out/r02_launches_acoustic_e_summary.txt; head -12 gpurun_out/r02_launches_acoustic_e_summary.txt
