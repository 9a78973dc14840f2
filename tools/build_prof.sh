#!/bin/bash
# Measurement build: same sources with -DSDF_CLOCKS (per-phase cycle counters in the chain epilogue, synchronous
# launches).  Output: build_prof/libsdfnet_prof.so; select it with SDFNET_LIB=$PWD/build_prof/libsdfnet_prof.so
set -e
cd "$(dirname "$0")/../sdf-representation_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xcompiler -fPIC -std=c++17 -DSDF_CLOCKS"
mkdir -p ../../build_prof
for f in chain dw aux api; do $NVCC $FLAGS -c $f.cu -o ../../build_prof/$f.o & done
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../../build_prof/libsdfnet_prof.so ../../build_prof/*.o -lcudart
echo built build_prof/libsdfnet_prof.so
