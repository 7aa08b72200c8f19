#!/bin/bash
# Builds libdcb200.so (the C-ABI library of include/dcb200.h) for sm_100a, in-tree.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="${DCB_EXTRA_FLAGS:-} -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -diag-suppress 177 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
OUT=${DCB_OUT:-libdcb200.so}      # experiment builds: DCB_OUT=libdcb200_exp.so DCB_EXTRA_FLAGS=-D...  (loaded via DCB200_LIB)
TAG=${OUT%.so}
$NVCC $FLAGS -c kernels.cu -o $TAG.kernels.o
$NVCC $FLAGS -Xcompiler -fvisibility=default -c engine.cu -o $TAG.engine.o
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $TAG.kernels.o $TAG.engine.o -Xlinker -soname=$OUT
echo "built $(pwd)/$OUT"
