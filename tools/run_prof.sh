#!/bin/bash
# ncu capture of the chain kernels (one fwd+rev launch, one tan+adj launch) and the launch list of a short bench.
mkdir -p gpurun_out
TAG=${TAG:-r01b}
CMD="python bench.py --points 131072 --steps 2 --warmup 3 --no-cpu-baseline --no-grid"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 6 -c 2 -o gpurun_out/prof_chain_$TAG $CMD > gpurun_out/ncu_chain.log 2>&1
tail -2 gpurun_out/ncu_chain.log
if [ -n "$DW" ]; then
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dw_kernel -s 3 -c 1 -o gpurun_out/prof_dw_$TAG $CMD > gpurun_out/ncu_dw.log 2>&1
tail -2 gpurun_out/ncu_dw.log
fi
if [ -n "$LIST" ]; then
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list.log 2>&1
fi
ls -la gpurun_out | tail -5
