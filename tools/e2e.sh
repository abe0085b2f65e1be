#!/bin/bash
export SMC_BENCH_CACHE=/tmp/smc_batch
for g in "$@"; do
  SMC_PIPE_CHUNKS=$g python bench.py --no-cpu-baseline --pipeline-intervals 0 --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('chunks %-3s e2e %.3f ms  h2d %.3f  -> %.0f loci/s   resident %.3f ms' % ('$g', e['ms_per_step'], e['ms_h2d'], e['value'], d['ms_per_step']))"
done
