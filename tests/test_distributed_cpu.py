"""CPU, world_size 2, gloo: the N>1 host path (row sharding + one sum-allreduce of [grads] and [loss sums]) gives the
single-process result.  The per-shard 'engine' here is the oracle, so this exercises exactly the host logic that
IGRLOSS.forward runs around the CUDA call on each rank."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import golden_params, load_golden, skip_of


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "..", "oracle"))
    sys.path.insert(0, os.path.join(here, ".."))
    sys.path.insert(0, here)
    import sdf_oracle as orc
    from sdf_representation_b200 import distributed as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    d = load_golden("softplus_skip_128x4.npz")
    W, b = golden_params(d)
    x = torch.from_numpy(d["x"]).double()[:332]  # divisible by 2: contiguous N/G rows per rank
    y = torch.from_numpy(d["y"]).double()[:332]
    lo, hi = D.shard_rows(x.shape[0])
    xs, ys = x[lo:hi], y[lo:hi]
    n_global = (hi - lo) * D.world_size()
    delta, tau, lam = 0.1, 1.0, 0.1
    # what IGRLOSS.forward does on each rank: global mean of clamped labels first ...
    t_sum = ys[:, 0].clamp(-delta, delta).sum().reshape(1)
    D.all_reduce_sum(t_sum)
    mean_t = t_sum / n_global
    # ... then per-shard seeds with GLOBAL normalisers, shard gradients, one sum-allreduce
    f, g = orc.forward_and_gradient(W, b, xs, skip_of(d), 100.0)
    _, _, fbar, gbar = orc.igr_terms(f, g, ys, delta, tau, lam, 1.0, "pairwise")
    scale = (hi - lo) / n_global
    p = f.clamp(-delta, delta)
    inside = ((f >= -delta) & (f <= delta)).double()
    fbar = (2.0 / n_global) * (p - mean_t) * inside
    gbar = gbar * scale
    dW, db = orc.param_grads(W, b, xs, skip_of(d), 100.0, fbar, gbar)
    flat = torch.cat([t.reshape(-1) for pair in zip(dW, db) for t in pair])
    D.all_reduce_sum(flat)
    if rank == 0:
        _, _, dWr, dbr = orc.igr_loss_and_grads(W, b, x, y, skip_of(d), 100.0, delta, tau, lam, 1.0, "pairwise")
        ref = torch.cat([t.reshape(-1) for pair in zip(dWr, dbr) for t in pair])
        ret["err"] = float((flat - ref).norm() / ref.norm())
        ret["world"] = D.world_size()
    dist.destroy_process_group()


def test_sharded_step_equals_single_process():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret["world"] == 2
    assert ret["err"] < 1e-10
