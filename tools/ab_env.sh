#!/bin/bash
# run on the GPU box: A/B of runtime knobs (env assignments) on the cached bench batch: tools/ab_env.sh "SMC_CHUNK=192" "SMC_CHUNK=256 SMC_B200_LIB=..."
export SMC_BENCH_CACHE=/tmp/smc_batch
for kv in "$@"; do
  env $kv python bench.py --no-cpu-baseline --no-strong --pipeline-intervals 0 --batches 2 --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms_per_batch_rank0']; print('%-60s total %.3f prep %.3f sort %.3f gather %.3f merge %.3f stats %.3f' % ('$kv', d['ms_per_step']/2, s['ms_prep'], s['ms_sort'], s['ms_k_gather'], s['ms_k_merge'], s['ms_stats']))"
done
