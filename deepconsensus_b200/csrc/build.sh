#!/bin/bash
# Builds libdcb200.so (the C-ABI library of include/dcb200.h) for sm_100a, in-tree.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="${DCB_EXTRA_FLAGS:-} -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
$NVCC $FLAGS -c kernels.cu -o kernels.o
$NVCC $FLAGS -Xcompiler -fvisibility=default -c engine.cu -o engine.o
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libdcb200.so kernels.o engine.o -Xlinker -soname=libdcb200.so
echo "built $(pwd)/libdcb200.so"
