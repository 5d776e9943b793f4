#!/bin/sh
# Builds libaz_b200.so (the C-ABI library of include/az_b200.h) in-tree for sm_100a.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
      -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared ${AZ_NVCC_EXTRA} \
      -o libaz_b200.so csrc/az_kernels.cu csrc/az_net.cu csrc/az_chess.cu csrc/az_gemm.cu
