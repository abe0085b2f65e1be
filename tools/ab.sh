#!/bin/bash
# run on the GPU box: A/B of the variant libraries on the cached bench batch (resident timing only)
export SMC_BENCH_CACHE=/tmp/smc_batch
for so in "$@"; do
  SMC_B200_LIB=$so python bench.py --no-cpu-baseline --no-strong --pipeline-intervals 0 --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms_per_batch_rank0']; print('%-44s total %.3f prep %.3f sort %.3f gather %.3f merge %.3f stats %.3f' % ('$so'.split('/')[-1], d['ms_per_step'], s['ms_prep'], s['ms_sort'], s['ms_k_gather'], s['ms_k_merge'], s['ms_stats']))"
done
