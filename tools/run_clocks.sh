#!/bin/bash
# per-phase cycle tables of the chain epilogue (SDF_CLOCKS measurement build) with the measurement switches
export SDFNET_LIB=$PWD/build_prof/libsdfnet_prof.so
for f in ${FLAGSETS:-0 0x80000 0x20000 0x300000 0x320000 0x10000}; do
  echo "=== SDFNET_DEBUG_FLAGS=$f"
  SDFNET_DEBUG_FLAGS=$f timeout 120 python tools/train_probe.py 2>&1 | sed -n '/--- step 1/,$p' | grep "^clk"
done
