#!/bin/bash
# compute-sanitizer over the CTA-pair weight-resident kernel (cluster barriers, multicast commits, remote arrives)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool: wres pair" > gpurun_out/r02_sanitizer_conv_wres_pair_${tool}.log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_umma.py -k "wres_pair and (32-3-1-3-13000 or 64-11-1-3-13001)" -q -x -p no:cacheprovider >> gpurun_out/r02_sanitizer_conv_wres_pair_${tool}.log 2>&1
  echo "exit code $?" >> gpurun_out/r02_sanitizer_conv_wres_pair_${tool}.log
  tail -3 gpurun_out/r02_sanitizer_conv_wres_pair_${tool}.log | cut -c1-160
done
