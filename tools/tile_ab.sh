#!/bin/bash
# A/B of the tile-staged pass kernel (env knobs of qb_engine_create x build variants) on C3.
# usage: tools/tile_ab.sh [ntraj]  -- one line per variant: min ms over QB_REPS batches
NT=${1:-1024}
export QB_REPS=${QB_REPS:-3}
run() { echo -n "$1 :: "; env $2 python tools/prof_run.py c3 $NT 2>&1 | tail -1; }
run "r1 SELL warp-autonomous        " "QB_NO_RSELL=1"
for lib in $(cd qutip_b200; ls lib_*.so 2>/dev/null || echo libqutip_b200.so); do
  for R in ${QB_AB_ROWS:-1024 2048}; do
    for THR in ${QB_AB_THR:-256}; do
      for NSB in ${QB_AB_NSB:-4 6 8}; do
        run "$lib rows=$R thr=$THR nsb=$NSB" "QUTIP_B200_LIB=$PWD/qutip_b200/$lib QB_TILE_ROWS=$R QB_TILE_THREADS=$THR QB_TILE_NSB=$NSB"
      done
    done
  done
done
