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
k = b["roofline"]["kernel_ms_per_step"]
x = b.get("extra", {})
par, pcd, grid, octr = x.get("train_step_parity"), x.get("pcd_train_step"), x.get("grid"), x.get("octree")
json.dump(b, open(f"{P}/{tag}_bench_1xB200.json", "w"), indent=1)
out = f"""# python bench.py --steps {b['steps']} --warmup {b['warmup']}          (1 B200; gpurun box, clocks {b['clocks']})
train_points_per_s {b['value']/1e6:.2f} M ({b['ms_per_step']:.2f} ms per 1M-point IGRLOSS step incl. fused Adam); e2e from pinned host memory {b['e2e']['ms_per_step']:.2f} ms = {b['e2e']['value']/1e6:.2f} M points/s
whole step: {b['roofline']['whole_step']['algorithmic_tflops']:.0f} TFLOP/s algorithmic (22.03 MFLOP/point) = {100*b['roofline']['whole_step']['frac']:.1f} % of MEASURED_PEAKS bf16_tflops_sustained ({b['roofline']['peak']})
kernel classes, CUDA events on the launch stream, ms per step: chain fwd+rev {k['chain_fwd_rev']:.2f} | chain tan+adj {k['chain_tan_adj']:.2f} | dW {k['dw_gemm']:.2f} | other {k['other']:.2f} | outside kernels {b['roofline']['outside_kernels_ms_per_step']:.2f}
dominant kernel ({b['roofline']['kernel']}): {b['roofline']['achieved']:.0f} TFLOP/s algorithmic (2F) = {100*b['roofline']['frac']:.1f} % of sustained peak; DRAM traffic per launch {('%.1f GB' % (b['roofline']['traffic']/1e9)) if b['roofline'].get('traffic') else 'n/a'} ({b['roofline']['traffic_source'][:60]})
gpu_launches in the timed region: {b['gpu_launches']} ({b['gpu_launches']/b['steps']:.0f} per step)
cpu_baseline: {b.get('cpu_baseline')}
"""
if par:
    out += f"parity precision (hi/lo split operands, inside the 3.46e-4 tolerance): {par['ms_per_step']:.1f} ms per step = {par['points_per_s']/1e6:.2f} M points/s = {100*par['frac_of_tensor_peak_algorithmic_6F']:.1f} % of peak counting 6F algorithmic\n"
if pcd:
    out += f"point-cloud step (BASELINE cfg 3 share: {pcd['surface_points_per_gpu']} surface + {pcd['perturbed_points_per_gpu']} perturbed points): {pcd['ms_per_step']:.2f} ms = {pcd['surface_points_per_s']/1e6:.1f} M surface points/s = {100*pcd['frac_of_tensor_peak_algorithmic_5F']:.1f} % of peak (5F)\n"
if grid:
    out += f"grid {grid['n']}^3 (BASELINE cfg 4, {grid['planes_per_rank']} planes on this rank, halo {grid['halo']}): {grid['ms_grid_eval']/1e3:.3f} s = {grid['queries_per_s']/1e6:.1f} M queries/s = {100*grid['frac_of_tensor_peak']:.1f} % of peak (1F); MC classify + compact {grid['ms_mc_classify_compact']:.1f} ms ({grid['active_cells_rank0']} active cells)\n"
if octr:
    out += f"octree to level {octr['max_level']} (BASELINE cfg 5): {octr['ms_build']:.1f} ms for {octr['centres_total']} centres = {octr['queries_per_s']/1e6:.0f} M queries/s; cells per level {octr['cells_per_level']}\n"
out += f"""# python bench.py --impl reference --steps 3 --warmup 1   ({r['cpu_baseline']['kind']}: {r['cpu_baseline']['sample'][:110]}, {r['cpu_baseline']['cores']} host threads)
{r['value']/1e3:.2f} k points/s ({r['ms_per_step']:.0f} ms per 8192-point step)
"""
open(f"{P}/{tag}_bench_1xB200.txt", "w").write(out)
print(out)

rep = f"{G}/prof_step_{tag}.ncu-rep"
pre = f"{G}/ncu_step_summary_{tag}.txt"
if os.path.exists(pre) and os.path.getsize(pre) > 0:
    # summarised on the GPU box (tools/run_round_end.sh): the report itself may not have come back (64 MiB merge limit)
    import shutil
    shutil.copy(pre, f"{P}/{tag}_train_step_1M_ncu.txt")
    if os.path.exists(f"{G}/dram_bytes_per_point_{tag}.json"):
        shutil.copy(f"{G}/dram_bytes_per_point_{tag}.json", f"{P}/dram_bytes_per_point.json")
    if os.path.exists(f"{G}/ncu_grid_summary_{tag}.txt"):
        shutil.copy(f"{G}/ncu_grid_summary_{tag}.txt", f"{P}/{tag}_grid_512_fwd_kernel_ncu.txt")
elif os.path.exists(rep):
    txt = subprocess.run([sys.executable, f"{ROOT}/tools/ncu_summary.py", rep, "14", "--json", f"{P}/dram_bytes_per_point.json", "--points", str(b["config"]["points_per_gpu"])], capture_output=True, text=True).stdout
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
    lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras (1M points, 4 steps + e2e leg)",
             "# kernel | launches | total ms | share"]
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
        lines.append(f"{n:60s} {c:5d} {t:10.3f} {100*t/tot:6.2f} %")
    open(f"{P}/{tag}_launch_list_1M.txt", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:8]))
