#!/bin/bash
# Build a variant of the library with extra -D flags for same-session A/B runs:
#   tools/build_variant.sh <name> <flags...>   ->  build_var/libsdfnet_<name>.so   (select with SDFNET_LIB)
set -e
NAME=$1; shift
cd "$(dirname "$0")/../sdf-representation_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xcompiler -fPIC -std=c++17 $*"
mkdir -p ../../build_var/$NAME
for f in chain dw aux api; do $NVCC $FLAGS -c $f.cu -o ../../build_var/$NAME/$f.o & done
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../../build_var/libsdfnet_$NAME.so ../../build_var/$NAME/*.o -lcudart
echo built build_var/libsdfnet_$NAME.so
