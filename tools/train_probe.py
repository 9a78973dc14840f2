"""One IGRLOSS step at a small size through the public API (use with the SDF_CLOCKS measurement build)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import sdf_oracle as orc
import sdf_representation_b200 as pkg
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = pkg.ImplicitNet(3, [512] * 8, (4,), 100.0, True, precision="fast").to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 8
g = torch.Generator().manual_seed(1)
x = torch.rand(n, 3, generator=g) * 2 - 1
y = orc.torus_labels(x).float()
crit = pkg.IGRLOSS(0.1, 1.0, 0.1, 1.0)
for i in range(2):
    print(f"--- step {i}", file=sys.stderr)
    net.zero_grad()
    loss = crit(x.to(dev), y.to(dev), net, i)
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss))
