#!/bin/bash
# same-session A/B of the CTA-pair chain kernel: SDFNET_PAIR=0 (never), train (training programs only), unset (default:
# everywhere)
for rep in 1 2; do
for pr in 0 train default; do
  if [ $pr = default ]; then unset SDFNET_PAIR; else export SDFNET_PAIR=$pr; fi
  g=$(python tools/grid_probe.py 384 2>&1 | grep -o "[0-9.]* Mq/s")
  out=$(python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-grid 2>/dev/null)
  b=$(echo "$out" | grep -o '"ms_per_step": [0-9.]*' | head -1)
  k=$(echo "$out" | grep -o '"kernel_ms_per_step": {[^}]*}')
  echo "PAIR=$pr rep$rep: grid $g | $b | $k"
done; done
