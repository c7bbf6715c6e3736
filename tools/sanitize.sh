#!/bin/bash
# compute-sanitizer evidence for the hand-rolled mbarrier / TMA / tcgen05 pipelines: one small case
# per kernel family under memcheck, racecheck and synccheck.  Logs -> gpurun_out/r02_sanitizer_*.log (copied to profiles/).
# Each run is bounded by `timeout`: the tools slow these kernels down by one to two orders of magnitude.
OUT=gpurun_out
run() {  # name tool pytest-args...
  local name=$1 tool=$2; shift 2
  echo "== $tool: $*" > $OUT/r02_sanitizer_${name}_${tool}.log
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest "$@" -q -x -p no:cacheprovider >> $OUT/r02_sanitizer_${name}_${tool}.log 2>&1
  echo "exit code $?" >> $OUT/r02_sanitizer_${name}_${tool}.log
  tail -4 $OUT/r02_sanitizer_${name}_${tool}.log | cut -c1-200
}
for tool in memcheck racecheck synccheck; do
  run diffnet $tool tests/test_gpu_diffnet.py -k "single_layer and (1-77-8 or 2-256-1)"
  run attention $tool tests/test_gpu_ops.py -k "relpos and (129 or 37)"
  run conv_stream $tool tests/test_gpu_umma.py -k "plain and stream_tma and (256-512-3-1-2-300 or 64-128-1-1-1-128)"
  run conv_pair $tool tests/test_gpu_umma.py -k "diffnet_chain and pair and not pair_co"
  run conv_chunked $tool tests/test_gpu_umma.py -k "chunked and 256-1024"
  run conv_wres $tool tests/test_gpu_umma.py -k "plain and stream_tma and 32-32-3-1-2-1000"
  run aa_mel $tool tests/test_gpu_ops.py tests/test_frontend.py -k "aa_snake_pair or mel_transform"
done
grep -l "ERROR SUMMARY: [1-9]" $OUT/r02_sanitizer_*.log
echo done
