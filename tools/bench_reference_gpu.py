#!/usr/bin/env python
"""Like-for-like GPU bar: the UNMODIFIED reference modules
(baseline/_ref/model/{networks,losses}.py, installed by baseline/install_ref.sh) on one B200 through stock PyTorch
(cuBLAS SGEMM + ATen element-wise kernels + autograd), fp32 with TF32 off (torch default) and on.

    python tools/bench_reference_gpu.py > gpurun_out/reference_gpu.json

Measured: (1) grid inference loop of executor/executor.py:374-383 on a 256^3 grid, postprocessbatchsize 262,144;
(2) the train loop body executor/executor.py:167-178 with the reference's IGRLOSS as written ([N,N] data term) at the
batch sizes the [N,N] matrix allows; (3) the same step with the O(N) closed-form data term (the reference's network,
compute_normal and autograd, only the broadcast replaced) at 262,144 points; (4) the point-cloud step of
executor/executordistributed.py:63-83.  CUDA events, 3 warm-ups, median of the timed steps.  Not part of the product.
"""
import contextlib
import io
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

F_FLOP = 3_671_040


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def timed(fn, steps=5, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    ms = []
    for i in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return statistics.median(ms)


def closed_form_igr(rl, model, x, y, delta=0.1, tau=1.0, lambda_g=0.1, thr=1.0):
    """model/losses.py:109-137 with the [N,1]-[N] broadcast's mean written as moments; everything else is
    the reference's own code path (forward, compute_normal -> autograd.grad(create_graph=True))."""
    pred = model(x)
    p = torch.clamp(pred, -delta, delta)[:, 0]
    t = torch.clamp(y[:, 0], -delta, delta)
    data = (p * p).mean() + (t * t).mean() - 2 * p.mean() * t.mean()
    g = rl.compute_normal(model, x)
    cos = torch.nn.functional.cosine_similarity(y[:, 1:], g / torch.norm(g), dim=-1)
    reg = torch.where(torch.abs(t) < thr, (1 - cos) ** 2, torch.tensor(1e-8, device=x.device))
    eik = torch.where(torch.abs(t) < thr, (g.norm(2, dim=-1) - 1) ** 2, torch.tensor(1e-8, device=x.device))
    return data + tau * reg.mean() + lambda_g * eik.mean()


def main():
    if not os.path.exists(os.path.join(REF, "model", "networks.py")):
        print(json.dumps({"unavailable": "baseline/_ref is missing (run baseline/install_ref.sh in the build container)"}))
        return
    import sdf_oracle as orc
    from model import losses as rl
    from model import networks as rn

    dev = torch.device("cuda:0")
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "results": []}
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        tag = "tf32" if tf32 else "fp32"
        torch.manual_seed(0)
        model = quiet(rn.ImplicitNet, 3, [512] * 8, (4,), 100.0, True).to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=4e-4)
        crit = rl.IGRLOSS(0.1, 1.0, 0.1, 1.0)

        # (1) grid loop, executor/executor.py:351-383
        n = 256
        ax = torch.linspace(-1, 1, n, device=dev)
        xx, yy, zz = torch.meshgrid(ax, ax, ax, indexing="ij")
        coords = torch.stack((xx, yy, zz), dim=-1).reshape(-1, 3)
        bs = 262144

        def grid(i):
            outs = []
            with torch.no_grad():
                for o in range(0, coords.shape[0], bs):
                    outs.append(model(coords[o:o + bs]))
            return torch.cat(outs)

        ms = timed(grid, 3, 1)
        out["results"].append({"what": "grid_256^3_forward", "mode": tag, "ms": ms, "queries_per_s": n ** 3 / (ms * 1e-3),
                               "tflops_algorithmic": n ** 3 * F_FLOP / (ms * 1e-3) / 1e12})
        del coords, xx, yy, zz

        # (2) train loop body with the reference's own IGRLOSS ([N,N] data term)
        for N in (8192, 16384, 32768):
            g = torch.Generator().manual_seed(1)
            x = (torch.rand(N, 3, generator=g) * 2 - 1).to(dev)
            y = orc.torus_labels(x.cpu()).float().to(dev)

            def step(i):
                opt.zero_grad()
                loss = quiet(crit, x.clone(), y, model, i)
                loss.backward()
                opt.step()
                return loss.item()

            try:
                ms = timed(step, 5, 3)
                out["results"].append({"what": f"igr_step_reference_NxN_N={N}", "mode": tag, "ms": ms, "points_per_s": N / (ms * 1e-3),
                                       "tflops_algorithmic_6F": N * 6 * F_FLOP / (ms * 1e-3) / 1e12,
                                       "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9})
            except RuntimeError as e:  # out of memory on the [N,N] temporaries
                out["results"].append({"what": f"igr_step_reference_NxN_N={N}", "mode": tag, "error": str(e)[:120]})
                torch.cuda.empty_cache()

        # (3) closed-form data term, large batch
        for N in (262144,):
            g = torch.Generator().manual_seed(1)
            x = (torch.rand(N, 3, generator=g) * 2 - 1).to(dev)
            y = orc.torus_labels(x.cpu()).float().to(dev)

            def step_cf(i):
                opt.zero_grad()
                loss = closed_form_igr(rl, model, x.clone(), y)
                loss.backward()
                opt.step()
                return loss.item()

            try:
                torch.cuda.reset_peak_memory_stats()
                ms = timed(step_cf, 5, 3)
                out["results"].append({"what": f"igr_step_closed_form_N={N}", "mode": tag, "ms": ms, "points_per_s": N / (ms * 1e-3),
                                       "tflops_algorithmic_6F": N * 6 * F_FLOP / (ms * 1e-3) / 1e12,
                                       "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9})
            except RuntimeError as e:
                out["results"].append({"what": f"igr_step_closed_form_N={N}", "mode": tag, "error": str(e)[:120]})
                torch.cuda.empty_cache()

            # (4) point-cloud step, executor/executordistributed.py:63-83 (gradient() of :10-19)
            xp0 = (x[torch.randperm(N, device=dev)[: N // 3]] + torch.randn(N // 3, 3, device=dev) * 1e-4)

            def step_pcd(i):
                xp = xp0.clone().requires_grad_(True)
                pred = model(x)
                pred_s = model(xp)
                surf = pred.abs().mean()
                ones = torch.ones_like(pred_s)
                gr = torch.autograd.grad(pred_s, xp, ones, create_graph=True, retain_graph=True, only_inputs=True)[0][:, -3:]
                loss = surf + 0.1 * ((gr.norm(2, dim=-1) - 1) ** 2).mean()
                opt.zero_grad()
                loss.backward()
                opt.step()
                return loss.item()

            try:
                ms = timed(step_pcd, 5, 3)
                out["results"].append({"what": f"pcd_step_N={N}", "mode": tag, "ms": ms, "surface_points_per_s": N / (ms * 1e-3),
                                       "tflops_algorithmic_5F": N * 5 * F_FLOP / (ms * 1e-3) / 1e12})
            except RuntimeError as e:
                out["results"].append({"what": f"pcd_step_N={N}", "mode": tag, "error": str(e)[:120]})
                torch.cuda.empty_cache()
        del model, opt
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
