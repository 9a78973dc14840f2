"""Point-cloud ("distributed" executor) training step of BASELINE config 3 on this rank's share: N surface points +
N//3 perturbed points (executor/executordistributed.py:63-83, 108-123), Adam inside.  Device time by CUDA events."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import sdf_representation_b200 as pkg

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = pkg.ImplicitNet(3, [512] * 8, (4,), 100.0, True, precision="fast").to(dev)
crit = pkg.PointCloudLoss(lambda_g=0.1)
opt = torch.optim.Adam(net.parameters(), lr=4e-4)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
g = torch.Generator().manual_seed(2)
d = torch.randn(n, 3, generator=g)
x = (d / d.norm(dim=1, keepdim=True) * 0.6).to(dev)       # points on a sphere surface


def step(i):
    opt.zero_grad(set_to_none=True)
    xp = pkg.get_points(x, on_device=True)                   # x[randperm[:N//3]] + randn * 1e-4, drawn on the device
    loss = crit(x, xp, net, i)
    loss.backward()
    opt.step()
    return loss


for i in range(3):
    step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
e0.record()
for i in range(K):
    loss = step(i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
F = 3_671_040
print(json.dumps({"workload": "pcd_step_8x512", "surface_points": n, "perturbed_points": n // 3, "ms_per_step": ms,
                  "surface_points_per_s": n / ms * 1e3, "algorithmic_tflops": 5 * F * n / ms * 1e3 / 1e12, "loss": float(loss)}))
