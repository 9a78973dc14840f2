"""2+ ranks on real GPUs (torchrun): the sharded IGRLOSS step (NCCL sum-allreduce of the flat gradient) must equal the
single-GPU step over the whole batch.  Also times slab-sharded grid inference.  SURVEY.md §8e / cfg 3, cfg 4."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import torch.distributed as dist
import sdf_oracle as orc
import sdf_representation_b200 as pkg
from sdf_representation_b200 import distributed as D, inference

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
net = pkg.ImplicitNet(3, [512] * 8, (4,), 100.0, True, precision="fast").to(dev)
n = 1 << 18
g = torch.Generator().manual_seed(7)
x = torch.rand(n, 3, generator=g) * 2 - 1
y = orc.torus_labels(x).float()
crit = pkg.IGRLOSS(0.1, 1.0, 0.1, 1.0)
lo, hi = D.shard_rows(n)
net.zero_grad()
loss = crit(x[lo:hi].to(dev), y[lo:hi].to(dev), net, 0)
loss.backward()
flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()
if rank == 0:
    # single-process reference over the whole batch: leave the process group out of the loss
    saved = D.is_on
    D.is_on = lambda: False
    net.zero_grad()
    loss1 = crit(x.to(dev), y.to(dev), net, 0)
    loss1.backward()
    D.is_on = saved
    ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    err = float((flat - ref).norm() / ref.norm())
    print(f"dp_check world={world}: loss sharded {float(loss):.7f} single {float(loss1):.7f}  grad rel err {err:.3e}")
    assert abs(float(loss) - float(loss1)) < 1e-5 and err < 1e-3, "sharded step differs from the single-GPU step"
dist.barrier()
# slab-sharded dense grid (cfg 4 shape at reduced n): device time, max over ranks
gn = 512
ix0, ix1 = D.slab_bounds(gn)
inference.grid_sdf(net, gn, ix0, min(ix0 + 4, ix1))
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = inference.grid_sdf(net, gn, ix0, ix1); e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"grid {gn}^3 over {world} ranks: {float(ms):.2f} ms (max over ranks) = {gn**3/float(ms)/1e3:.1f} Mq/s")
dist.destroy_process_group()
