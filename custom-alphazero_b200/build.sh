#!/bin/sh
# Builds libaz_b200.so (the C-ABI library of include/az_b200.h) in-tree for sm_100a.
# The five translation units are compiled in parallel and linked into one shared object.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -fvisibility=hidden ${AZ_NVCC_EXTRA}"
mkdir -p build
pids=""
for unit in az_kernels az_net az_chess az_gemm az_tower; do
    $NVCC $FLAGS -c -o build/$unit.o csrc/$unit.cu &
    pids="$pids $!"
done
for pid in $pids; do
    wait $pid
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o libaz_b200.so build/az_kernels.o build/az_net.o build/az_chess.o build/az_gemm.o build/az_tower.o
