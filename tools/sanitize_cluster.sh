#!/bin/bash
# re-run of the racecheck cases of the two cluster kernels (tools/sanitize.sh has the full matrix)
OUT=gpurun_out
run() {
  local name=$1; shift
  echo "== racecheck: $*" > $OUT/r02_sanitizer_${name}_racecheck.log
  timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest "$@" -q -x -p no:cacheprovider >> $OUT/r02_sanitizer_${name}_racecheck.log 2>&1
  echo "exit code $?" >> $OUT/r02_sanitizer_${name}_racecheck.log
  tail -3 $OUT/r02_sanitizer_${name}_racecheck.log | cut -c1-200
}
run diffnet tests/test_gpu_diffnet.py -k "single_layer and (1-77-8 or 2-256-1)"
run conv_pair tests/test_gpu_umma.py -k "diffnet_chain and pair and not pair_co"
