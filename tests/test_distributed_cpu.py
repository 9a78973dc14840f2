"""CPU, world_size 2, gloo: the PRODUCT host code of the data-parallel step (losses.IGRLOSS.forward /
PointCloudLoss.forward / WeightedSmoothL2Loss.forward: label statistics, one sum-allreduce of [gradients | loss sums], loss value, autograd hand-over;
distributed.shard_rows / broadcast_parameters) on UNEQUAL shards gives the single-process result.  The per-rank engine
is tests/fake_engine.py (oracle arithmetic): the CUDA engine needs a GPU, the host logic around it does not."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_losses(pkg, orc, FakeEngine, x, y, xp, lo, hi, plo, phi, D):
    torch.manual_seed(3)
    net = pkg.ImplicitNet(3, [64] * 3, (2,), 100.0, True).double()
    if D.is_on():
        with torch.no_grad():  # make rank 1 differ so that broadcast_parameters matters
            for p in net.parameters():
                p.add_(0.01 * D.rank())
        D.broadcast_parameters(net)
    fake = FakeEngine(net)
    net._engine_for = lambda _x: fake
    out = {}
    for name, crit, args in (
        ("igr_pairwise", pkg.IGRLOSS(0.1, 1.0, 0.1, 1.0), (x[lo:hi], y[lo:hi])),
        ("igr_pointwise_mask", pkg.IGRLOSS(0.1, 1.0, 0.1, 0.05, data_term="pointwise"), (x[lo:hi], y[lo:hi])),
        ("pcd", pkg.PointCloudLoss(0.1), (x[lo:hi], xp[plo:phi])),
        ("wsl2_pairwise", pkg.WeightedSmoothL2Loss(0.5, 0.1), (x[lo:hi], y[lo:hi])),
        ("wsl2_pointwise", pkg.WeightedSmoothL2Loss(0.5, 0.1, data_term="pointwise"), (x[lo:hi], y[lo:hi])),
    ):
        for p in net.parameters():
            p.grad = None
        loss = crit(*args, net, 0)
        loss.backward()
        terms = getattr(crit, "last_terms", None)
        out[name] = (float(loss.detach()), torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone(),
                     torch.zeros(1, dtype=torch.float64) if terms is None else terms.clone())
    return out


def _worker(rank, world, port, ret):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "..", "oracle"))
    sys.path.insert(0, os.path.join(here, ".."))
    sys.path.insert(0, here)
    import sdf_oracle as orc
    import sdf_representation_b200 as pkg
    from fake_engine import FakeEngine
    from sdf_representation_b200 import distributed as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    g = torch.Generator().manual_seed(1)
    n = 333                                   # odd: shards of 167 + 166 rows
    x = (torch.rand(n, 3, generator=g) * 2 - 1).double()
    y = orc.torus_labels(x).float().double()  # the product path hands float32 labels to the engine
    y[::17, 1:] = 0.0
    xp = pkg.get_points(x, generator=torch.Generator().manual_seed(3))  # 111 rows: 56 + 55
    lo, hi = D.shard_rows(n)
    plo, phi = D.shard_rows(xp.shape[0])
    assert (hi - lo) != n // world or rank == 1
    sharded = _run_losses(pkg, orc, FakeEngine, x, y, xp, lo, hi, plo, phi, D)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        single = _run_losses(pkg, orc, FakeEngine, x, y, xp, 0, n, 0, xp.shape[0], D)
        for name in single:
            ls, gs, ts = sharded[name]
            l1, g1, t1 = single[name]
            ret[name] = (abs(ls - l1), float((gs - g1).norm() / g1.norm()), float((ts - t1).abs().max()))
        # and the single-process product path agrees with the oracle's own loss functions
        torch.manual_seed(3)
        net = pkg.ImplicitNet(3, [64] * 3, (2,), 100.0, True).double()
        W = [getattr(net, f"lin{i}").weight.detach() for i in range(4)]
        b = [getattr(net, f"lin{i}").bias.detach() for i in range(4)]
        f32 = lambda v: torch.tensor(v, dtype=torch.float32).item()  # the C ABI carries the hyper-parameters as float
        lref, _, dW, db = orc.igr_loss_and_grads(W, b, x, y, (2,), 100.0, f32(0.1), 1.0, f32(0.1), f32(0.05), "pointwise")
        ref = torch.cat([t.reshape(-1) for pair in zip(dW, db) for t in pair])
        ret["oracle"] = (abs(single["igr_pointwise_mask"][0] - float(lref)), float((single["igr_pointwise_mask"][1] - ref).norm() / ref.norm()))
        # the fused WeightedSmoothL2Loss host path against the oracle's restatement of model/losses.py:50-69 (both data terms)
        for key, dt in (("wsl2_pairwise", "pairwise"), ("wsl2_pointwise", "pointwise")):
            lw, dWw, dbw = orc.wsl2_loss_and_grads(W, b, x, y, (2,), 100.0, 0.5, 0.1, dt)  # (python floats reach the fake engine unrounded)
            refw = torch.cat([t.reshape(-1) for pair in zip(dWw, dbw) for t in pair])
            ret["oracle_" + key] = (abs(single[key][0] - float(lw)), float((single[key][1] - refw).norm() / refw.norm()))


def test_sharded_step_equals_single_process():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for name in ("igr_pairwise", "igr_pointwise_mask", "pcd", "wsl2_pairwise", "wsl2_pointwise"):
        dl, dg, dt = ret[name]
        assert dl < 1e-12 and dg < 1e-10 and dt < 1e-12, (name, dl, dg, dt)
    assert ret["oracle"][0] < 1e-10 and ret["oracle"][1] < 1e-9
    for key in ("oracle_wsl2_pairwise", "oracle_wsl2_pointwise"):
        assert ret[key][0] < 1e-10 and ret[key][1] < 1e-9, (key, ret[key])
