#!/bin/bash
# Round-end evidence in one GPU call: parity suite, smoke, default bench (with every BASELINE config under "extra"),
# reference arm, full-size ncu capture (one launch of each hot kernel) and launch list.   TAG=r02a bash tools/run_round_end.sh
mkdir -p gpurun_out
TAG=${TAG:-r02a}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_$TAG.json
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"chain_kernel|dw2_kernel" -s 6 -c 3 -o gpurun_out/prof_step_$TAG $CMD > gpurun_out/ncu_step.log 2>&1
tail -1 gpurun_out/ncu_step.log
$CMD > gpurun_out/plain_full2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log
if [ -n "$GRIDPROF" ]; then
  python tools/grid_probe.py 512 > gpurun_out/grid_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"chain_kernel" -s 1 -c 1 -o gpurun_out/prof_grid_$TAG python tools/grid_probe.py 512 > gpurun_out/ncu_grid.log 2>&1
  tail -1 gpurun_out/ncu_grid.log
fi
# summaries on the box (ncu -i needs no GPU); the reports themselves come back only while gpurun_out stays under the 64 MiB
# merge limit
[ -f gpurun_out/prof_step_$TAG.ncu-rep ] && python tools/ncu_summary.py gpurun_out/prof_step_$TAG.ncu-rep 14 --json gpurun_out/dram_bytes_per_point_$TAG.json --points 1048576 > gpurun_out/ncu_step_summary_$TAG.txt 2>/dev/null
[ -f gpurun_out/prof_grid_$TAG.ncu-rep ] && python tools/ncu_summary.py gpurun_out/prof_grid_$TAG.ncu-rep 16 > gpurun_out/ncu_grid_summary_$TAG.txt 2>/dev/null
for f in gpurun_out/prof_grid_$TAG.ncu-rep gpurun_out/prof_step_$TAG.ncu-rep; do
  if [ "$(du -sm gpurun_out | cut -f1)" -ge 60 ] && [ -f $f ]; then echo "dropping $f (merge limit)"; rm -f $f; fi
done

