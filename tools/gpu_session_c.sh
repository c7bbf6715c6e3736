python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_c.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c.log
python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_c.jsonl 2> gpurun_out/r02_bench_c.err; tail -c 300 gpurun_out/r02_bench_c.err
PTTSPP_UMMA_LONGK=narrow python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_c_narrow.jsonl 2>/dev/null
python - <<'PY'
import json
for f in ('r02_bench_c','r02_bench_c_narrow'):
    d=json.loads(open(f'gpurun_out/{f}.jsonl').read().strip().splitlines()[-1])
    print(f, d['ms_per_step'], d['bigvgan']['ms_per_step'], [ (k['kernel'],round(k['ms'],2)) for k in d['bigvgan']['kernel_families'] if k['ms']>0], [ (k['kernel'],round(k['ms'],2)) for k in d['kernel_families'] if k['ms']>0])
PY
