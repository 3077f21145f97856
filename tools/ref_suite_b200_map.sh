#!/bin/bash
# the reference's own mcsolve / nm_mcsolve tests with the device map as the default map
cd "$(dirname "$0")/.."
REF=$PWD/oracle/_ref
export PYTHONPATH=$REF:$PWD QUTIP_B200_DEFAULT_MAP=1 OMP_NUM_THREADS=1
python -m pytest -p qutip_b200.plugin -q -p no:cacheprovider --timeout 300 "$@" \
    $REF/qutip/tests/solver/test_mcsolve.py $REF/qutip/tests/solver/test_nm_mcsolve.py
