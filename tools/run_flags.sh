#!/bin/bash
# kernel-class times of the 1M-point step under the measurement switches (SDF_CLOCKS build; results are wrong when a
# switch is set, only the timing is meaningful):  FLAGSETS="0 0x100000 ..." bash tools/run_flags.sh
export SDFNET_LIB=$PWD/build_prof/libsdfnet_prof.so
for f in ${FLAGSETS:-0 0x100000 0x200000 0x300000 0x320000 0x10000 0x40000}; do
  out=$(SDFNET_DEBUG_FLAGS=$f timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>/tmp/clk_$f.txt)
  k=$(echo "$out" | grep -o '"kernel_ms_per_step": {[^}]*}')
  echo "FLAGS $f: $k"
  grep "^clk" /tmp/clk_$f.txt | tail -8
done
