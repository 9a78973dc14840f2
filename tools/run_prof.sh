#!/bin/bash
# ncu capture of the chain kernels (one fwd+rev launch, one tan+adj launch) of a short bench.
mkdir -p gpurun_out
TAG=${TAG:-r01b}
CMD="python bench.py --points ${POINTS:-131072} --steps 2 --warmup 3 --no-cpu-baseline --no-grid"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 6 -c ${COUNT:-2} -o gpurun_out/prof_chain_$TAG $CMD > gpurun_out/ncu_chain.log 2>&1
tail -2 gpurun_out/ncu_chain.log
ls -la gpurun_out | tail -3
