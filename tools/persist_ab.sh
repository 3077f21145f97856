#!/bin/bash
# A/B: one CTA per tile vs persistent work-queue grid (QB_TILE_PERSIST = CTAs per SM)
NT=${1:-2048}
export QB_REPS=${QB_REPS:-3}
for p in 0 3 0 3 2; do echo -n "persist=$p :: "; QB_TILE_PERSIST=$p python tools/prof_run.py c3 $NT 2>&1 | tail -1; done
