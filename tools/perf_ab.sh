#!/bin/bash
# A/B timing of kernel tuning variants (qutip_b200/lib_*.so built with different -D flags);
# with no variants present, times the default library.  Variants are visited round-robin
# twice and every visit repeats the batch (QB_REPS) -- box noise is larger than most effects.
libs=$(ls qutip_b200/lib_*.so 2>/dev/null)
[ -z "$libs" ] && libs=qutip_b200/libqutip_b200.so
for pass in 1 2; do
for lib in $libs; do
  echo "== $lib"
  [ -z "$QB_AB_NOC3" ] && QB_REPS=${QB_REPS:-4} QUTIP_B200_LIB=$PWD/$lib python tools/prof_run.py c3 ${QB_AB_NTRAJ:-1024} 2>&1 | tail -1
  [ -n "$QB_AB_C2" ] && QB_REPS=${QB_REPS:-4} QUTIP_B200_LIB=$PWD/$lib python tools/prof_run.py ${QB_AB_C2} 2>&1 | tail -2
done
done
