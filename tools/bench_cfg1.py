#!/usr/bin/env python
"""BASELINE configs[0] timed three ways on one box (BASELINE.md section 5, row "cfg 1"): the compute of the shipped
tests/test.ini pipeline -- 8 x 512 ReLU / tanh ImplicitNet, WeightedSmoothL2Loss(0.5, 0.1), Adam, batch 4096 over the
37,800 training rows a 42,000-row torus sample leaves after the 10 % split, 3 epochs, then the 128^3 grid of
executor/executor.py:351-387 -- with

  (a) the UNMODIFIED reference modules (baseline/_ref) on the host cores,
  (b) the same modules on cuda (stock PyTorch fp32, TF32 off),
  (c) the drop-in modules (libsdfnet.so).

    python tools/bench_cfg1.py > gpurun_out/cfg1.json

Synthetic analytic-torus labels stand in for igl.signed_distance (sampling is out of scope).  Wall clock with a device
synchronize on both sides.  Not part of the product.
"""
import contextlib
import io
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
for p in (ROOT, os.path.join(ROOT, "oracle"), REF):
    if p not in sys.path:
        sys.path.insert(0, p)

import sdf_oracle as orc  # noqa: E402  (label generator only)

N_TRAIN, BATCH, EPOCHS, GRID = 37800, 4096, 3, 128


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def pipeline(net, loss_fn, opt, x, y, dev, grid_fn):
    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize()

    sync()
    t0 = time.perf_counter()
    for epoch in range(EPOCHS):
        perm = torch.randperm(N_TRAIN, generator=torch.Generator().manual_seed(epoch))
        for b0 in range(0, N_TRAIN, BATCH):  # executor/executor.py:167-178
            idx = perm[b0:b0 + BATCH].to(dev)
            opt.zero_grad()
            loss = quiet(loss_fn, x[idx], y[idx], net, epoch)
            loss.backward()
            opt.step()
            last = loss.item()
    sync()
    t1 = time.perf_counter()
    vol = grid_fn()
    sync()
    t2 = time.perf_counter()
    return {"train_s": t1 - t0, "grid_s": t2 - t1, "total_s": t2 - t0, "last_loss": last, "grid_mean": float(vol.mean())}


def reference_grid(net, dev, chunk=262144):
    # executor/executor.py:351-383
    ax = torch.linspace(-1, 1, GRID, device=dev)
    xx, yy, zz = torch.meshgrid(ax, ax, ax, indexing="ij")
    coords = torch.stack((xx, yy, zz), dim=-1).reshape(-1, 3)
    out = []
    with torch.no_grad():
        for i in range(0, coords.shape[0], chunk):
            out.append(net(coords[i:i + chunk]))
    return torch.cat(out).reshape(GRID, GRID, GRID)


def main():
    g = torch.Generator().manual_seed(100)
    x = torch.rand(N_TRAIN, 3, generator=g) * 2 - 1
    y = orc.torus_labels(x.double()).float()
    res = {"config": {"net": "8x512 ReLU/tanh (tests/test.ini)", "loss": "WeightedSmoothL2Loss(0.5, 0.1)", "train_rows": N_TRAIN, "batch": BATCH,
                      "epochs": EPOCHS, "grid": GRID}}
    have_ref = os.path.exists(os.path.join(REF, "model", "networks.py"))
    if have_ref:
        with contextlib.redirect_stdout(io.StringIO()):
            from model import losses as rl
            from model import networks as rn
        torch.set_num_threads(os.cpu_count() or 1)
        for name, dev in (("reference_cpu", torch.device("cpu")), ("reference_cuda_fp32", torch.device("cuda", 0))):
            if dev.type == "cuda" and not torch.cuda.is_available():
                continue
            torch.manual_seed(0)
            net = quiet(rn.ImplicitNet, 3, [512] * 8, (), 0, False).to(dev)
            opt = torch.optim.Adam(net.parameters(), lr=5e-4)
            xd, yd = x.to(dev), y.to(dev)
            if dev.type == "cuda":  # warm-up (cuBLAS handles, autotune)
                pipeline(net, rl.WeightedSmoothL2Loss(0.5, 0.1), opt, xd, yd, dev, lambda: reference_grid(net, dev))
            res[name] = pipeline(net, rl.WeightedSmoothL2Loss(0.5, 0.1), opt, xd, yd, dev, lambda: reference_grid(net, dev))
            res[name]["threads"] = os.cpu_count() if dev.type == "cpu" else None
    if torch.cuda.is_available():
        import sdf_representation_b200 as pkg
        from sdf_representation_b200 import inference

        dev = torch.device("cuda", 0)
        for mode, optname in (("fast", "FusedAdam"), ("fast", "torch.optim.Adam"), ("parity", "FusedAdam")):
            torch.manual_seed(0)
            net = pkg.ImplicitNet(3, [512] * 8, (), 0, False, precision=mode).to(dev)
            opt = pkg.FusedAdam(net, lr=5e-4) if optname == "FusedAdam" else torch.optim.Adam(net.parameters(), lr=5e-4)
            xd, yd = x.to(dev), y.to(dev)
            pipeline(net, pkg.WeightedSmoothL2Loss(0.5, 0.1), opt, xd, yd, dev, lambda: inference.reconstruct_grid(net, GRID))
            key = f"dropin_{mode}" + ("" if optname == "FusedAdam" else "_torch_adam")
            res[key] = pipeline(net, pkg.WeightedSmoothL2Loss(0.5, 0.1), opt, xd, yd, dev, lambda: inference.reconstruct_grid(net, GRID))
            res[key]["optimizer"] = optname
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
