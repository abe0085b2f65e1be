#!/bin/bash
# (tuning aid) builds A/B variants of libsmc_b200.so into gpurun_variants/ (git-ignored, travels with gpurun): tools/build_variants.sh name "-DKNOB=.." ...
# usage: build_variants.sh name "flags" [name "flags" ...]   -> gpurun_variants/<name>.so
cd "$(dirname "$0")/.."
while [ $# -gt 1 ]; do
  n=$1; f=$2; shift 2
  SMC_NVCC_EXTRA="$f" SMC_B200_OUT=$PWD/gpurun_variants/$n.so python -c "
from smcounter_b200 import build; build.build(force=True)" >/dev/null 2>&1 && echo built $n || echo FAILED $n
done
