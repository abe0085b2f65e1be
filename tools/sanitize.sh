#!/bin/bash
# run on the GPU box: compute-sanitizer memcheck + racecheck over smoke() and two cases of the differential fuzz.
# Summaries land in gpurun_out/r02_sanitizer_*.log
cd "$(dirname "$0")/.."
run() {  # tool, tag, command...
  tool=$1; tag=$2; shift 2
  timeout 240 compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 9 "$@" > gpurun_out/r02_sanitizer_${tool}_${tag}.log 2>&1
  echo "$tool $tag rc=$?  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/r02_sanitizer_${tool}_${tag}.log | tr '\n' ' ')"
}
run memcheck smoke python -c "import __graft_entry__ as g; g.smoke()"
SMC_FUZZ_SEEDS=106:108 run memcheck fuzz python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "randomised or many_dynamic or long_insertions or compact"
run racecheck smoke python -c "import __graft_entry__ as g; g.smoke()"
SMC_FUZZ_SEEDS=106:108 run racecheck fuzz python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "randomised or compact_upload_encodings_give_identical_bits and 3-2"
for f in gpurun_out/r02_sanitizer_*.log; do echo "== $f"; tail -6 $f; done
