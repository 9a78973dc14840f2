#!/usr/bin/env python
"""Headline benchmark: the fused IGRLOSS training step (forward + input gradient + weight backward + Adam) of the
8x512 skip-in geometric-init SDF network on 1M synthetic points per GPU (BASELINE.json configs[1]; weak scaling
across ranks = configs[2]'s data-parallel shape), plus dense-grid query throughput as a side figure.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision fast|parity]

One JSON line on rank 0 (contract in the task statement).  ``value`` = whole-job points/s with inputs resident in
HBM; ``e2e`` = the same step fed from pinned host memory through the public drop-in API (H2D copy of the batch and D2H
of the loss inside the timed region).  ``--impl reference`` times the CPU port of the reference path (oracle/) on the
host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

F_FLOP = 3_671_040            # FLOP per point per forward of the 8x512 skip-(4,) net (verified)
TRAIN_FLOP = 6 * F_FLOP       # algorithmic FLOP per IGRLOSS training point (fwd 1 + input-grad 1 + backward 2 + dW 2)
POINTS_PER_GPU = 1 << 20
N_BATCHES = 8                 # distinct resident batches cycled through: 8 x 28 MB = 224 MB > 126 MB L2
# measured DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this command at
# 1M points, profiles/r01q_train_step_1M_ncu.txt) per point and kernel class; roofline.traffic scales it to the launch
NCU_DRAM_BYTES_PER_POINT = {"chain_fwd_rev": 22665.0, "chain_tan_adj": 62263.0, "dw_gemm": 28829.0}
METRIC = "train_points_per_s"
UNIT = "points/s"


def workload_name():
    return "igr_train_step_8x512_skip4_beta100_1Mpts_per_gpu"


def make_labels(x):
    import sdf_oracle as orc  # synthetic analytic torus labels (test-infrastructure generator, not on the timed path)

    return orc.torus_labels(x)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_step(n_pts, reps, threads):
    """IGRLOSS step of the CPU port (oracle restatement of model/networks.py + model/losses.py + the backward the
    reference gets from autograd), fp32, all host threads.  Returns points/s."""
    import sdf_oracle as orc

    torch.set_num_threads(threads)
    torch.manual_seed(0)
    shapes = orc.layer_shapes(3, [512] * 8, (4,))
    W = [torch.randn(o, i) * (2.0 ** 0.5 / o ** 0.5) for (o, i) in shapes]
    b = [torch.zeros(o) for (o, _) in shapes]
    W[-1] = torch.full((1, 512), (3.14159265 ** 0.5) / 512 ** 0.5)
    b[-1] = torch.full((1,), -1.0)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(n_pts, 3, generator=g) * 2 - 1
    y = orc.torus_labels(x)
    orc.igr_loss_and_grads(W, b, x[:256], y[:256], (4,), 100.0)  # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        orc.igr_loss_and_grads(W, b, x, y, (4,), 100.0)
    dt = time.perf_counter() - t0
    return n_pts * reps / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU path (here: its port, oracle/sdf_oracle.py -- the reference is Python and
    /root/reference does not exist on the GPU box) on the host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 8192
    t_all = []
    import sdf_oracle as orc  # noqa: F401

    for i in range(args.warmup + args.steps):
        v, dt = cpu_port_step(sample, 1, threads)
        if i >= args.warmup:
            t_all.append(dt)
    ms = 1e3 * sum(t_all) / len(t_all)
    value = sample / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "sample_points_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample}-point IGRLOSS step (closed-form O(N) data term) per timed step, torch CPU fp32"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft")
    ap.add_argument("--precision", default=os.environ.get("SDFNET_BENCH_PRECISION", "fast"))
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--points", type=int, default=POINTS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-grid", action="store_true", help="skip the dense-grid side figure (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist

    import sdf_representation_b200 as pkg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = args.points

    # ---- network: BASELINE configs[1] (SURVEY.md §8d cfg 2): seed-0 geometric init, identical on all ranks
    torch.manual_seed(0)
    net = pkg.ImplicitNet(3, [512] * 8, (4,), 100.0, True, precision=args.precision, dtype=args.dtype).to(dev)
    crit = pkg.IGRLOSS(0.1, 1.0, 0.1, 1.0)
    opt = torch.optim.Adam(net.parameters(), lr=4e-4)
    eng = net.engine()

    # ---- synthetic data: N_BATCHES distinct batches per rank, labels = analytic torus
    host_x, host_y = [], []
    for bidx in range(N_BATCHES):
        g = torch.Generator().manual_seed(1 + 1000 * rank + bidx)
        x = torch.rand(n, 3, generator=g) * 2 - 1
        host_x.append(x.pin_memory())
        host_y.append(make_labels(x).float().pin_memory())
    dev_x = [t.to(dev) for t in host_x]
    dev_y = [t.to(dev) for t in host_y]

    def step(x, y, i):
        opt.zero_grad(set_to_none=True)
        loss = crit(x, y, net, i)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up (the clock sampler starts here so that nvidia-smi is already streaming in the timed region; its
    # median is taken over samples under load only)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(dev_x[i % N_BATCHES], dev_y[i % N_BATCHES], i)
    barrier()

    # ---- timed region 1: inputs resident in HBM
    launches0 = eng.launch_count()
    eng.profile(True)
    ms_dev = timed(lambda i: step(dev_x[i % N_BATCHES], dev_y[i % N_BATCHES], i), args.steps)
    prof = eng.profile_read()
    eng.profile(False)
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2: end to end through the public API, host buffers in, loss out
    def e2e_step(i):
        x = host_x[i % N_BATCHES].to(dev, non_blocking=True)
        y = host_y[i % N_BATCHES].to(dev, non_blocking=True)
        return step(x, y, i).item()

    ms_e2e = timed(e2e_step, args.steps)

    # ---- side figure: dense-grid SDF queries/s on this rank's slab of a 512^3 grid (cfg 4 shape, reduced n)
    from sdf_representation_b200 import inference

    gn = 512
    grid_qps = 0.0
    if not args.no_grid:
        ix0, ix1 = pkg.distributed.slab_bounds(gn, rank, world)
        inference.grid_sdf(net, gn, ix0, min(ix0 + 8, ix1))
        ms_grid = timed(lambda i: inference.grid_sdf(net, gn, ix0, ix1), 1)
        grid_qps = gn ** 3 / (ms_grid * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"

    pts_total = n * world * args.steps
    value = pts_total / (ms_dev * 1e-3)
    e2e = pts_total / (ms_e2e * 1e-3)
    # dominant kernel = the class with the largest accumulated device time; each chain launch runs 2 sweeps (2F per
    # point), the dW launch is 2F per point as well
    dom = max(("chain_fwd_rev", "chain_tan_adj", "dw_gemm"), key=lambda k: prof[k][0])
    dom_ms, dom_n = prof[dom]
    pts_per_launch = n * args.steps / max(dom_n, 1)
    achieved = (2 * F_FLOP * pts_per_launch) / (dom_ms / max(dom_n, 1) * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    step_tf = TRAIN_FLOP * n / (ms_dev / args.steps * 1e-3) / 1e12

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16" if args.dtype == "fp16" else "bf16", "data": "synthetic",
        "config": {
            "workload": workload_name(), "points_per_gpu": n, "global_points_per_step": n * world,
            "precision_mode": args.precision, "operands": f"{args.dtype} operands, fp32 TMEM accumulate",
            "loss": "IGRLOSS(0.1,1,0.1,1) pairwise data term", "optimizer": "torch.optim.Adam(lr=4e-4) inside the timed step",
            "l2": f"{N_BATCHES} distinct resident batches cycled ({N_BATCHES * n * 28 / 1e6:.0f} MB > 126 MB L2); the step also "
                  "streams its multi-GB spill workspace through L2",
            "parallelism": f"dp{world}: contiguous row shards, one NCCL sum-allreduce of the flat gradient per step" if world > 1 else "single GPU",
        },
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": n * 28 * world, "d2h_bytes_per_step": 4 * world,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved / peak_tf, "traffic": NCU_DRAM_BYTES_PER_POINT[dom] * pts_per_launch,
            "traffic_source": "ncu --set full, profiles/r01q_train_step_1M_ncu.txt (bytes/point of this kernel x points per launch)",
            "peak_source": peak_src,
            "per_launch_ms": dom_ms / max(dom_n, 1), "launches_timed": dom_n,
            "algorithmic_flop_per_point_this_kernel": 2 * F_FLOP,
            "whole_step": {"algorithmic_tflops": step_tf, "frac": step_tf / peak_tf, "flop_per_point": TRAIN_FLOP},
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
        },
        "extra": {"grid_queries_per_s": grid_qps * 1.0, "grid": f"{gn}^3 slab-sharded over {world} rank(s), max over ranks",
                  "grid_frac_of_tensor_peak": grid_qps * F_FLOP / 1e12 / peak_tf},
    }
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt = cpu_port_step(8192, 2, threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"2 x 8192-point IGRLOSS steps of the oracle port (torch CPU fp32, {dt:.1f} s)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
