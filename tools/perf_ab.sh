#!/bin/bash
# A/B timing of kernel tuning variants (qutip_b200/lib_*.so built with different -D flags)
for lib in qutip_b200/lib_*.so; do
  echo "== $lib"
  QUTIP_B200_LIB=$PWD/$lib python tools/prof_run.py c3 1024 2>&1 | tail -1
  QUTIP_B200_LIB=$PWD/$lib python tools/prof_run.py c2 2>&1 | tail -2
done
