#!/bin/bash
# Build libsdfnet.so (sm_100a only) in-tree.  Used by __graft_entry__.build().
set -e
cd "$(dirname "$0")/sdf-representation_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xcompiler -fPIC -std=c++17"
mkdir -p ../../build
for f in chain dw aux api; do
  if [ ! -f ../../build/$f.o ] || [ $f.cu -nt ../../build/$f.o ] || [ mc_tables.h -nt ../../build/$f.o ] || [ ptx.cuh -nt ../../build/$f.o ] || [ sdf_internal.h -nt ../../build/$f.o ] || [ ../../include/sdfnet.h -nt ../../build/$f.o ]; then
    $NVCC $FLAGS -c $f.cu -o ../../build/$f.o &
  fi
done
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../libsdfnet.so ../../build/chain.o ../../build/dw.o ../../build/aux.o ../../build/api.o -lcudart
# native query tool (equivalent of the reference's ops/DeepTrace and ops/conversion_test programs)
mkdir -p ../bin
if [ ! -f ../bin/sdf_query ] || [ tools/sdf_query.cpp -nt ../bin/sdf_query ] || [ ../libsdfnet.so -nt ../bin/sdf_query ]; then
  $NVCC -O2 -std=c++17 -o ../bin/sdf_query tools/sdf_query.cpp -L.. -lsdfnet -lcudart -Xlinker -rpath -Xlinker '$ORIGIN/..'
fi
echo "built $(cd .. && pwd)/libsdfnet.so"
