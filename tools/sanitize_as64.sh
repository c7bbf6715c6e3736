mkdir -p gpurun_out
PTTSPP_AA_FUSE=1 timeout 600 python -m pytest tests/test_gpu_models.py -q -m gpu -x -k "bigvgan or vocoder" 2>&1 | tail -2
for tool in memcheck racecheck synccheck; do
  echo "== $tool: as64 (256-row units)" > gpurun_out/r02_sanitizer_conv_as64_${tool}.log
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_umma.py -k "plain and stream_tma and (64-64-11-5-4-12100 or 64-64-11-5-1-12000)" -q -x -p no:cacheprovider >> gpurun_out/r02_sanitizer_conv_as64_${tool}.log 2>&1
  echo "exit code $?" >> gpurun_out/r02_sanitizer_conv_as64_${tool}.log
  tail -3 gpurun_out/r02_sanitizer_conv_as64_${tool}.log | cut -c1-160
done
