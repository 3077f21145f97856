#!/bin/bash
# A/B timing of kernel tuning variants (qutip_b200/lib_*.so built with different -D flags);
# with no variants present, times the default library.
libs=$(ls qutip_b200/lib_*.so 2>/dev/null)
[ -z "$libs" ] && libs=qutip_b200/libqutip_b200.so
for lib in $libs; do
  echo "== $lib"
  QUTIP_B200_LIB=$PWD/$lib python tools/prof_run.py c3 1024 2>&1 | tail -1
  QUTIP_B200_LIB=$PWD/$lib python tools/prof_run.py c2 2>&1 | tail -2
done
