"""Forward-only (dense grid) timing of the chain kernel; SDFNET_DEBUG_FLAGS selects measurement-only variants."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdf_representation_b200 as pkg
from sdf_representation_b200 import inference
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = pkg.ImplicitNet(3, [512] * 8, (4,), 100.0, True, precision="fast").to(dev)
gn = int(sys.argv[1]) if len(sys.argv) > 1 else 384
inference.grid_sdf(net, gn, 0, 8)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); inference.grid_sdf(net, gn, 0, gn); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"flags={os.environ.get('SDFNET_DEBUG_FLAGS','0')} grid {gn}^3: {ms:.2f} ms, {gn**3/ms/1e3:.1f} Mq/s, {gn**3/128/148*8/ms*1e-3:.0f}... per-layer-step {ms*1e3/(gn**3/128/148*8):.2f} us")
