#!/bin/bash
# Runs the UMMA descriptor probe over the encodings the library relies on (each in its own process, bounded).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/probe.log 2>&1
for mode in 0 1; do for n in 256 128 16; do for k in 64 128; do for bf in 1 0; do for sw in 0; do
  timeout 30 ./tools/umma_probe $mode $n $k $bf $sw >> gpurun_out/probe.log 2>&1 || echo "  (exit $?)" >> gpurun_out/probe.log
done; done; done; done; done
cat gpurun_out/probe.log
for n in 256 128 64; do for ts in 0 1; do
  timeout 60 ./tools/mma_rate $n $ts 2000 148 >> gpurun_out/probe.log 2>&1 || echo "  (exit $?)" >> gpurun_out/probe.log
done; done
timeout 60 ./tools/mma_rate 256 0 2000 1 >> gpurun_out/probe.log 2>&1
tail -8 gpurun_out/probe.log
