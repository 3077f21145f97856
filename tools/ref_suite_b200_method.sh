#!/bin/bash
# the reference's own solver tests with a device integrator as the default method
cd "$(dirname "$0")/.."
REF=$PWD/oracle/_ref
export PYTHONPATH=$REF:$PWD QUTIP_B200_DEFAULT_METHOD=${1:-b200_vern7} OMP_NUM_THREADS=1
shift
T=$REF/qutip/tests/solver
python -m pytest -p qutip_b200.plugin -q -p no:cacheprovider --timeout 600 "$@" \
    $T/test_mesolve.py $T/test_sesolve.py $T/test_propagator.py $T/test_correlation.py $T/test_floquet.py $T/test_mcsolve.py
