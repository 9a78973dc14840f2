#!/bin/bash
mkdir -p gpurun_out
for cfg in "256 0" "256 1" "128 0" "128 1" "64 0"; do set -- $cfg; timeout 60 ./tools/mma_rate $1 $2 2000 148 1 >> gpurun_out/probe2.log 2>&1; done
cat gpurun_out/probe2.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench1.err; cat gpurun_out/bench1.json
