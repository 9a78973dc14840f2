#!/bin/bash
# same-session A/B: short bench (+ optional grid probe, GRID=1) for the default library and each build_var/ variant,
# REPS times interleaved:   VARIANTS="a b" [REPS=2] [GRID=1] bash tools/ab.sh
for rep in $(seq 1 ${REPS:-2}); do
for v in default $VARIANTS; do
  if [ $v = default ]; then unset SDFNET_LIB; else export SDFNET_LIB=$PWD/build_var/libsdfnet_$v.so; fi
  g=""
  if [ -n "$GRID" ]; then g="grid $(python tools/grid_probe.py 384 2>&1 | grep -o '[0-9.]* Mq/s') |"; fi
  out=$(python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null)
  b=$(echo "$out" | grep -o '"ms_per_step": [0-9.]*' | head -1)
  k=$(echo "$out" | grep -o '"kernel_ms_per_step": {[^}]*}')
  echo "AB $v rep$rep: $g $b | $k"
done; done
