#!/bin/bash
# Full-size (1M points) ncu capture of the three hot kernels of one training step + launch list.
mkdir -p gpurun_out
TAG=${TAG:-r01c}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-grid"
$CMD > gpurun_out/plain_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"chain_kernel|dw_kernel" -s 9 -c 3 -o gpurun_out/prof_step_$TAG $CMD > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
$CMD > gpurun_out/plain_full2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log; ls -la gpurun_out | tail -4
