#!/bin/bash
# Build libqutip_b200.so for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2,-pthread ${QB_NVCC_EXTRA}"
$NVCC $FLAGS -c qb_ops.cu -o qb_ops.o &
$NVCC $FLAGS -c qb_engine.cu -o qb_engine.o &
$NVCC $FLAGS -c qb_dense.cu -o qb_dense.o &
$NVCC $FLAGS -c qb_comm.cu -o qb_comm.o &
$NVCC $FLAGS -c qb_build.cu -o qb_build.o &
wait
OUT=${QB_OUT:-../libqutip_b200.so}
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT qb_ops.o qb_engine.o qb_dense.o qb_comm.o qb_build.o -lcudart -ldl
echo "built $OUT"
