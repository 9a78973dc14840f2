#!/usr/bin/env python
"""Turn the outputs of tools/run_round_end.sh (gpurun_out/*_TAG.*) into the tracked summaries under profiles/.

    python tools/make_profile_summary.py r01p
"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def last_json(path):
    return json.loads([l for l in open(path) if l.startswith("{")][-1])


b = last_json(f"{G}/bench_{tag}.json")
r = last_json(f"{G}/bench_ref_{tag}.json")
pcd = last_json(f"{G}/bench_pcd_{tag}.txt")
inf = last_json(f"{G}/inference_{tag}.json")
json.dump(inf, open(f"{P}/{tag}_inference_1xB200.json", "w"))
k = b["roofline"]["kernel_ms_per_step"]
out = f"""# python bench.py --steps 10 --warmup 3          (1 B200; gpurun box, clocks {b['clocks']})
train_points_per_s {b['value']/1e6:.2f} M ({b['ms_per_step']:.2f} ms per 1M-point IGRLOSS step incl. Adam); e2e from pinned host memory {b['e2e']['ms_per_step']:.2f} ms
whole step: {b['roofline']['whole_step']['algorithmic_tflops']:.0f} TFLOP/s algorithmic (22.03 MFLOP/point) = {100*b['roofline']['whole_step']['frac']:.1f} % of MEASURED_PEAKS bf16_tflops_sustained (1384.8)
kernel classes, CUDA events on the launch stream, ms per step: chain fwd+rev {k['chain_fwd_rev']:.2f} | chain tan+adj {k['chain_tan_adj']:.2f} | dW {k['dw_gemm']:.2f} | other {k['other']:.2f}
dominant kernel (tan+adj): {b['roofline']['achieved']:.0f} TFLOP/s algorithmic (2F) = {100*b['roofline']['frac']:.1f} % of sustained peak
grid 512^3 side figure: {b['extra']['grid_queries_per_s']/1e6:.1f} M queries/s = {100*b['extra']['grid_frac_of_tensor_peak']:.1f} % of sustained peak (1F per query)
# python bench.py --impl reference --steps 3 --warmup 1   (oracle port of the reference path, torch CPU fp32, {r['cpu_baseline']['cores']} host threads)
{r['value']/1e3:.2f} k points/s (8192-point IGRLOSS steps, {r['ms_per_step']:.0f} ms per step)
# python tools/bench_pcd.py                      (BASELINE config 3 per-GPU share: 1M surface + 349525 perturbed points)
{pcd['ms_per_step']:.2f} ms per step = {pcd['surface_points_per_s']/1e6:.1f} M surface points/s = {pcd['algorithmic_tflops']:.0f} TFLOP/s algorithmic (5F per surface point) = {100*pcd['algorithmic_tflops']/1384.8:.1f} % of sustained peak
# python tools/bench_inference.py 1024           (profiles/{tag}_inference_1xB200.json)
1024^3 grid: {inf['grid_ms']/1e3:.2f} s on one GPU = {inf['grid_queries_per_s']/1e6:.1f} M queries/s ({100*inf['grid_frac_of_sustained_bf16_peak']:.1f} % of sustained peak); octree to level 10 (sphere, {inf['octree_queries']/1e6:.1f} M cell centres): {inf['octree_level10_ms']:.0f} ms; marching-cubes classification of 256 planes of 1024^2: {inf['mc_classify_256_planes_ms']:.1f} ms
"""
open(f"{P}/{tag}_bench_1xB200.txt", "w").write(out)
print(out)

rep = f"{G}/prof_step_{tag}.ncu-rep"
if os.path.exists(rep):
    txt = subprocess.run([sys.executable, f"{ROOT}/tools/ncu_summary.py", rep, "12"], capture_output=True, text=True).stdout
    open(f"{P}/{tag}_train_step_1M_ncu.txt", "w").write(txt)

lst = f"{G}/launches_{tag}.csv"
if os.path.exists(lst):
    rows = list(csv.reader(open(lst)))
    h = [i for i, row in enumerate(rows) if "Kernel Name" in row][0]
    ix = {n: i for i, n in enumerate(rows[h])}
    agg = defaultdict(lambda: [0, 0.0])
    for row in rows[h + 2:]:
        if len(row) <= ix["Metric Value"]:
            continue
        v = float(row[ix["Metric Value"]].replace(",", ""))
        u = row[ix["Metric Unit"]]
        v = v / 1e6 if u.startswith("ns") else v / 1e3 if u.startswith("us") else v * 1e3 if u in ("s", "second") else v
        a = agg[row[ix["Kernel Name"]][:60]]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-grid (1M points, 4 steps + e2e leg)",
             "# kernel | launches | total ms | share"]
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
        lines.append(f"{n:60s} {c:5d} {t:10.3f} {100*t/tot:6.2f} %")
    open(f"{P}/{tag}_launch_list_1M.txt", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:8]))
