"""Generates tests/golden/*.npz by running the UNMODIFIED reference (``/root/reference/model/networks.py``,
``model/losses.py``, loop bodies of ``executor/executor.py:174-176`` and ``executor/executordistributed.py:63-79``)
on CPU in the build container.  The reference cannot travel to the GPU box, the vectors can.

    python oracle/make_golden.py            # needs /root/reference

Test infrastructure only.
"""
import contextlib
import hashlib
import io
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SDF_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from model import losses as ref_losses  # noqa: E402
from model import networks as ref_networks  # noqa: E402

import sdf_oracle as orc  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def perturb(model, seed, scale):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(1 + scale * torch.randn(p.shape, generator=g))
            p.add_(0.02 * scale * torch.randn(p.shape, generator=g))


def params_np(model):
    return {k.replace(".", "_"): v.detach().numpy().copy() for k, v in model.state_dict().items()}


def grads_np(model, prefix):
    out = {}
    for k, p in model.named_parameters():
        out[prefix + k.replace(".", "_")] = p.grad.detach().numpy().copy()
        p.grad = None
    return out


def points(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 3, generator=g) * 2 - 1


def case_softplus():
    torch.manual_seed(7)
    model = quiet(ref_networks.ImplicitNet, 3, [128] * 4, (2,), 100.0, True)
    perturb(model, 11, 0.3)
    n = 333
    x = points(n, 1)
    y = orc.torus_labels(x)
    y[::17, 1:] = 0.0  # zero label normals: cosine-similarity eps path
    d = {"skip": np.int64(2), "beta": np.float64(100.0), "hidden": np.int64(128), "num_hidden": np.int64(4), "x": x.numpy(), "y": y.numpy()}
    d.update(params_np(model))
    # forward + compute_normal (model/networks.py:87-106, model/losses.py:283-296)
    f = model(x).detach()
    g = ref_losses.compute_normal(model, x.clone()).detach()
    d["f"], d["g"] = f.numpy()[:, 0], g.numpy()
    # IGRLOSS step (model/losses.py:109-137 + executor/executor.py:174-176)
    loss_fn = ref_losses.IGRLOSS(0.1, 1.0, 0.1, 1.0)
    loss = quiet(loss_fn, x.clone(), y, model, 0)
    loss.backward()
    d["igr_loss"] = np.float64(loss.item())
    d.update(grads_np(model, "igr_grad_"))
    # WeightedSmoothL2Loss (model/losses.py:50-69)
    loss = ref_losses.WeightedSmoothL2Loss(0.5, 0.1)(x.clone(), y, model, 0)
    loss.backward()
    d["wsl2_loss"] = np.float64(loss.item())
    d.update(grads_np(model, "wsl2_grad_"))
    # point-cloud step (executor/executordistributed.py:63-79)
    gen = torch.Generator().manual_seed(3)
    xp = (x[torch.randperm(n, generator=gen)[: n // 3]] + torch.randn(n // 3, 3, generator=gen) * 1e-4).clone()
    xp.requires_grad = True
    pred = model(x)
    pred_s = model(xp)
    surf = pred.abs().mean()
    ones = torch.ones_like(pred_s)
    grads = torch.autograd.grad(pred_s, xp, ones, create_graph=True, retain_graph=True, only_inputs=True)[0][:, -3:]
    eik = ((grads.norm(2, dim=-1) - 1) ** 2).mean()
    loss = surf + 0.1 * eik
    loss.backward()
    d["xp"] = xp.detach().numpy()
    d["pcd_loss"] = np.float64(loss.item())
    d["pcd_terms"] = np.array([surf.item(), eik.item()])
    d.update(grads_np(model, "pcd_grad_"))
    np.savez_compressed(os.path.join(OUT, "softplus_skip_128x4.npz"), **d)


def case_relu():
    # the shipped ini's model family: skip_connection = 0 => ReLU + tanh, geometric_init False (tests/test_config.ini)
    torch.manual_seed(5)
    model = quiet(ref_networks.ImplicitNet, 3, [128] * 3, (), 0, False)
    n = 257
    x = points(n, 2)
    y = orc.torus_labels(x)
    d = {"skip": np.int64(0), "beta": np.float64(0.0), "hidden": np.int64(128), "num_hidden": np.int64(3), "x": x.numpy(), "y": y.numpy()}
    d.update(params_np(model))
    f = model(x).detach()
    g = ref_losses.compute_normal(model, x.clone()).detach()
    d["f"], d["g"] = f.numpy()[:, 0], g.numpy()
    loss = ref_losses.WeightedSmoothL2Loss(0.5, 0.1)(x.clone(), y, model, 0)
    loss.backward()
    d["wsl2_loss"] = np.float64(loss.item())
    d.update(grads_np(model, "wsl2_grad_"))
    loss_fn = ref_losses.IGRLOSS(0.1, 1.0, 0.1, 1.0)
    loss = quiet(loss_fn, x.clone(), y, model, 0)
    loss.backward()
    d["igr_loss"] = np.float64(loss.item())
    d.update(grads_np(model, "igr_grad_"))
    np.savez_compressed(os.path.join(OUT, "relu_tanh_128x3.npz"), **d)


def case_full():
    # BASELINE configs 2-5 network: weights are NOT stored (7.36 MB); the drop-in's seed-identical init regenerates
    # them (hash checked) -- SURVEY.md §8d cfg 2.
    torch.manual_seed(0)
    model = quiet(ref_networks.ImplicitNet, 3, [512] * 8, (4,), 100.0, True)
    h = hashlib.sha256()
    for v in model.state_dict().values():
        h.update(v.numpy().tobytes())
    x = points(512, 1)
    y = orc.torus_labels(x)
    f = model(x).detach()
    g = ref_losses.compute_normal(model, x.clone()).detach()
    loss_fn = ref_losses.IGRLOSS(0.1, 1.0, 0.1, 1.0)
    loss = quiet(loss_fn, x.clone(), y, model, 0)
    loss.backward()
    gnorm = np.array([p.grad.norm().item() for p in model.parameters()])
    gsum = np.array([p.grad.double().sum().item() for p in model.parameters()])
    # a 9^3 grid through the reference's own coordinate construction (executor/executor.py:351-356)
    ax = torch.linspace(-1, 1, 9)
    xx, yy, zz = torch.meshgrid(ax, ax, ax, indexing="ij")
    coords = torch.stack((xx, yy, zz), dim=-1).reshape(-1, 3)
    grid = model(coords).detach().numpy().reshape(9, 9, 9)
    np.savez_compressed(
        os.path.join(OUT, "full_8x512_seed0.npz"), weights_sha256=np.array(h.hexdigest()), x=x.numpy(), y=y.numpy(),
        f=f.numpy()[:, 0], g=g.numpy(), igr_loss=np.float64(loss.item()), igr_grad_norms=gnorm, igr_grad_sums=gsum,
        grid9=grid,
    )


def ref_igr_pointwise(model, x, y, delta, tau, lambda_g, thr):
    """IGRLOSS with the per-point data term mean((p_i - t_i)^2) the authors presumably meant, built from the
    reference's own pieces: model/networks.py forward, model/losses.py:283-296 compute_normal and the expressions of
    model/losses.py:114-136 with ``predicted_sdf[:, 0]`` in place of the broadcasting ``predicted_sdf``."""
    pred = model(x)
    p = torch.clamp(pred, -delta, delta)[:, 0]
    t = torch.clamp(y[:, 0], -delta, delta)
    data = (p - t) ** 2
    g = ref_losses.compute_normal(model, x)
    n_true = y[:, 1:]
    cos = torch.nn.functional.cosine_similarity(n_true, g / torch.norm(g), dim=-1)
    reg = torch.where(torch.abs(t) < thr, (1 - cos) ** 2, torch.tensor(1e-8))
    eik = torch.where(torch.abs(t) < thr, (g.norm(2, dim=-1) - 1) ** 2, torch.tensor(1e-8))
    return (data + tau * reg + lambda_g * eik).mean()


def case_round2_small():
    """128x4 net: (1) the |t| < regularizer_threshold mask branch of model/losses.py:127-134 with thr = 0.05 < delta;
    (2) the pointwise data term; (3) a 60-step loss curve of the reference loop body executor/executor.py:167-178
    (zero_grad, IGRLOSS, backward, Adam.step, loss.item()) from identical seeded weights and identical batches."""
    torch.manual_seed(7)
    model = quiet(ref_networks.ImplicitNet, 3, [128] * 4, (2,), 100.0, True)
    perturb(model, 11, 0.3)
    n = 333
    x = points(n, 1)
    y = orc.torus_labels(x)
    y[::17, 1:] = 0.0
    d = {"skip": np.int64(2), "beta": np.float64(100.0), "hidden": np.int64(128), "num_hidden": np.int64(4), "x": x.numpy(), "y": y.numpy()}
    d.update(params_np(model))
    loss = quiet(ref_losses.IGRLOSS(0.1, 1.0, 0.1, 0.05), x.clone(), y, model, 0)
    loss.backward()
    d["mask_loss"] = np.float64(loss.item())
    d["mask_frac"] = np.float64((torch.clamp(y[:, 0], -0.1, 0.1).abs() < 0.05).float().mean().item())
    d.update(grads_np(model, "mask_grad_"))
    for thr, tag in ((1.0, "pw"), (0.05, "pwmask")):
        loss = ref_igr_pointwise(model, x.clone(), y, 0.1, 1.0, 0.1, thr)
        loss.backward()
        d[tag + "_loss"] = np.float64(loss.item())
        d.update(grads_np(model, tag + "_grad_"))
    # loss curve: geometric init (no perturbation), pool of 4096 points, step i trains on rows [1024 (i % 4), +1024)
    torch.manual_seed(21)
    model = quiet(ref_networks.ImplicitNet, 3, [128] * 4, (2,), 100.0, True)
    pool_x = points(4096, 5)
    pool_y = orc.torus_labels(pool_x)
    d["curve_x"], d["curve_y"] = pool_x.numpy(), pool_y.numpy()
    d.update({"curve_init_" + k: v for k, v in params_np(model).items()})
    opt = torch.optim.Adam(model.parameters(), lr=4e-4)
    crit = ref_losses.IGRLOSS(0.1, 1.0, 0.1, 1.0)
    curve = []
    for i in range(60):
        lo = 1024 * (i % 4)
        opt.zero_grad()
        loss = quiet(crit, pool_x[lo:lo + 1024].clone(), pool_y[lo:lo + 1024], model, i)
        loss.backward()
        opt.step()
        curve.append(loss.item())
    d["curve_loss"] = np.array(curve)
    d.update({"curve_final_" + k: v for k, v in params_np(model).items()})
    np.savez_compressed(os.path.join(OUT, "round2_128x4.npz"), **d)


def case_round2_full():
    """BASELINE cfg 2 network, 8192 points (the reference's native batch scale, [N,N] data term included): per-element
    parameter gradients of the unmodified reference, stored as a strided sample (every 61st element of each weight
    tensor, biases in full) -- 1.84 M gradients are 7 MB."""
    torch.manual_seed(0)
    model = quiet(ref_networks.ImplicitNet, 3, [512] * 8, (4,), 100.0, True)
    h = hashlib.sha256()
    for v in model.state_dict().values():
        h.update(v.numpy().tobytes())
    n = 8192
    x = points(n, 31)
    y = orc.torus_labels(x)
    y[::23, 1:] = 0.0
    loss = quiet(ref_losses.IGRLOSS(0.1, 1.0, 0.1, 1.0), x.clone(), y, model, 0)
    loss.backward()
    d = {"weights_sha256": np.array(h.hexdigest()), "x": x.numpy(), "y": y.numpy(), "igr_loss": np.float64(loss.item()), "stride": np.int64(61)}
    for k, p in model.named_parameters():
        g = p.grad.detach().numpy()
        key = "igr_grad_" + k.replace(".", "_")
        d[key] = g.copy() if g.ndim == 1 else g.reshape(-1)[::61].copy()
        d["igr_gnorm_" + k.replace(".", "_")] = np.float64(np.linalg.norm(g.astype(np.float64)))
        d["igr_gmax_" + k.replace(".", "_")] = np.float64(np.abs(g).max())
    np.savez_compressed(os.path.join(OUT, "round2_full_8x512_n8192.npz"), **d)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["round1", "round2"]
    if "round1" in which:
        case_softplus()
        case_relu()
        case_full()
    if "round2" in which:
        case_round2_small()
        case_round2_full()
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))
