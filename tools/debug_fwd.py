"""Debug aid: forward / gradient error of the CUDA path against the oracle for a ladder of small configurations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import sdf_oracle as orc
import sdf_representation_b200 as pkg

dev = torch.device("cuda:0")
torch.set_printoptions(precision=5, linewidth=160)

def run(H, L, skip, beta, n=300, geo=True, dtype="fp16", precision="fast"):
    torch.manual_seed(0)
    net = pkg.ImplicitNet(3, [H] * L, (skip,) if skip else (), beta, geo, precision=precision, dtype=dtype).to(dev)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(n, 3, generator=g) * 2 - 1
    W = [getattr(net, f"lin{i}").weight.detach().double().cpu() for i in range(L + 1)]
    b = [getattr(net, f"lin{i}").bias.detach().double().cpu() for i in range(L + 1)]
    f_ref, g_ref = orc.forward_and_gradient(W, b, x.double(), (skip,) if skip else (), beta)
    try:
        with torch.no_grad():
            f = net(x.to(dev))[:, 0].cpu().double()
        ef = (f - f_ref).abs().max().item()
        f2, gr = net.sdf_and_gradient(x.to(dev))
        eg = (gr.cpu().double() - g_ref).abs().max().item()
        ef2 = (f2[:, 0].cpu().double() - f_ref).abs().max().item()
        print(f"H={H} L={L} skip={skip} beta={beta} {dtype}/{precision}: max|df| {ef:.3e} (fg path {ef2:.3e})  max|dg| {eg:.3e}")
        if ef > 5e-3 or eg > 2e-2:
            print("   f    :", f[:6].numpy()); print("   f_ref:", f_ref[:6].numpy())
            print("   g    :", gr[:2].cpu().numpy().ravel()); print("   g_ref:", g_ref[:2].numpy().ravel())
            bad = (f - f_ref).abs() > 5e-3
            print("   bad rows:", bad.nonzero().ravel()[:20].tolist(), "of", n)
    except Exception as e:
        print(f"H={H} L={L} skip={skip} beta={beta}: EXC {e}")

for dtype in ("fp16", "bf16"):
    run(64, 2, 0, 100.0, dtype=dtype)
    run(64, 2, 0, 0.0, dtype=dtype)
    run(128, 2, 0, 100.0, dtype=dtype)
    run(128, 4, 2, 100.0, dtype=dtype)
    run(256, 4, 2, 100.0, dtype=dtype)
    run(512, 8, 4, 100.0, dtype=dtype)
run(512, 8, 0, 0.0, geo=False)
run(512, 8, 4, 100.0, n=20000)
for dtype in ("fp16", "bf16"):
    run(64, 2, 0, 100.0, dtype=dtype, precision="parity")
    run(128, 4, 2, 100.0, dtype=dtype, precision="parity")
    run(512, 8, 4, 100.0, dtype=dtype, precision="parity")
run(512, 8, 0, 0.0, geo=False, precision="parity")
