#!/bin/bash
# build kernel-tuning variants: tools/build_variants.sh name1 "flags1" name2 "flags2" ...
# -> qutip_b200/lib_<name>.so (A/B scripts pick them up; never committed)
cd "$(dirname "$0")/../qutip_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  (
    d=$(mktemp -d)
    F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2,-pthread $flags"
    for f in qb_ops qb_engine qb_dense qb_comm qb_build; do $NVCC $F -c $f.cu -o $d/$f.o & done
    wait
    $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../lib_$name.so $d/*.o -lcudart -ldl
    rm -rf $d; echo "built lib_$name.so"
  ) &
done
wait
