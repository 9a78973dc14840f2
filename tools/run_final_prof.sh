#!/bin/bash
# Round-end evidence: full-size ncu capture (one launch of each hot kernel), launch list, inference figures.
mkdir -p gpurun_out
TAG=${TAG:-r01f}
python tools/bench_inference.py 1024 > gpurun_out/inference_$TAG.json 2> gpurun_out/inference_$TAG.err; echo "inference rc=$?"; cut -c1-600 gpurun_out/inference_$TAG.json
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-grid"
$CMD > gpurun_out/plain_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"chain_kernel|dw2_kernel" -s 6 -c 3 -o gpurun_out/prof_step_$TAG $CMD > gpurun_out/ncu_step.log 2>&1
tail -1 gpurun_out/ncu_step.log
$CMD > gpurun_out/plain_full2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log
