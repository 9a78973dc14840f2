#!/bin/bash
mkdir -p gpurun_out
for cfg in "256 0 2" "128 0 2" "128 1 2" "64 0 2"; do set -- $cfg; timeout 60 ./tools/mma_rate $1 $2 2000 148 $3 >> gpurun_out/probe3.log 2>&1; done
cat gpurun_out/probe3.log
CMD="python bench.py --points 131072 --steps 2 --warmup 3 --no-cpu-baseline --no-grid"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 6 -c 2 -o gpurun_out/prof_chain_r1 $CMD > gpurun_out/ncu2.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dw_kernel -s 3 -c 1 -o gpurun_out/prof_dw_r1 $CMD > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu2.log gpurun_out/ncu3.log; ls -la gpurun_out
