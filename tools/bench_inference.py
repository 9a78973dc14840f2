"""Inference-side figures of BASELINE configs 4 and 5 on one GPU (or one rank's share under torchrun):
dense 1024^3 grid by ix-slabs and an adaptive octree to level 10 (trained-ish weights: geometric init = sphere).
Device time by CUDA events.  Prints one JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdf_representation_b200 as pkg
from sdf_representation_b200 import inference

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = pkg.ImplicitNet(3, [512] * 8, (4,), 100.0, True, precision="fast").to(dev)
F = 3_671_040


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), r


n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
inference.grid_sdf(net, n, 0, 2)
out = torch.empty(n * n * n, dtype=torch.float32, device=dev)
ms_grid, vol = timed(lambda: inference.grid_sdf(net, n, 0, n, out=out))
inference.mc_classify(vol[: min(n, 256)])  # warm-up at the timed size (the caching allocator then holds the 267 MB blocks)
ms_mc, (cases, active) = timed(lambda: inference.mc_classify(vol[: min(n, 256)]))
inference.octree_build(net, 10)  # warm-up at the timed depth (caching allocator, on-demand weight layout)
ms_oct, levels = timed(lambda: inference.octree_build(net, 10))
per_level = []
for lv in levels:
    per_level.append({"level": lv["level"], "cells": int(lv["keys"].numel()), "refined": int(lv["flags"].sum())})
tot_q = sum(p["cells"] for p in per_level)
print(json.dumps({
    "grid_n": n, "grid_ms": ms_grid, "grid_queries_per_s": n ** 3 / ms_grid * 1e3,
    "grid_frac_of_sustained_bf16_peak": n ** 3 / ms_grid * 1e3 * F / 1384.8e12,
    "mc_classify_256_planes_ms": ms_mc, "mc_active_cells": int(active.numel()),
    "octree_level10_ms": ms_oct, "octree_queries": tot_q, "octree_queries_per_s": tot_q / ms_oct * 1e3, "octree_levels": per_level,
}))
