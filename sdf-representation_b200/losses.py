"""Drop-ins for the loss callables of the reference's ``model/losses.py`` that sit on the hot path.

Protocol (reference ``model/losses.py:5``): ``loss(x_batch, y_batch, model, epoch) -> scalar`` with ``__name__()``
and, for IGRLOSS, the attribute ``lambda_g`` (read by ``executor/executordistributed.py:26``).

``IGRLOSS`` recognises the drop-in ``ImplicitNet`` and runs ONE fused training step in the CUDA library
(forward, input gradient, loss seeds + warp-shuffle loss reductions, tangent / adjoint sweeps, weight-gradient
GEMMs); the returned scalar carries a graph whose backward only hands the already computed parameter gradients to
autograd, so ``loss.backward(); optimizer.step()`` of ``executor/executor.py:174-177`` works unchanged.

Data term: the reference subtracts a ``[N]`` label from a ``[N,1]`` prediction, which broadcasts to
``[N,N]``; ``data_term="pairwise"`` (default) reproduces exactly that mean over all pairs in O(N) with moment sums,
``data_term="pointwise"`` is the per-point loss the authors presumably meant.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from .networks import ImplicitNet, _split_flat
from . import distributed as _dist


def _unwrap(model):
    m = model.module if isinstance(model, (nn.DataParallel, nn.parallel.DistributedDataParallel)) else model
    if not isinstance(m, ImplicitNet):
        raise TypeError(
            "sdf-representation_b200 losses drive the fused CUDA path of the drop-in ImplicitNet; got "
            f"{type(m).__name__}.  There is no generic eager fallback."
        )
    return m


class _FusedLoss(torch.autograd.Function):
    """loss value -> a scalar whose backward hands the gradient the fused step has already accumulated to autograd.

    The gradient lives in the engine's accumulators (summed over ranks).  If a ``FusedAdam`` (optim.py) drives the
    network, backward only records d(total)/d(loss) as a device scalar and leaves the accumulators to the optimizer's
    fused unpack + Adam + re-pack kernel (no per-tensor work at all, ``p.grad`` stays unset unless the optimizer was
    built with ``materialize_grads=True``).  Otherwise the flat gradient is unpacked and scattered to ``p.grad`` so that
    ``torch.optim.Adam`` (reference executor/executor.py:160,177) works unchanged."""

    @staticmethod
    def forward(ctx, loss_value, net, eng, *params):
        ctx.net, ctx.eng = net, eng
        ctx.params = params
        return loss_value.clone()

    @staticmethod
    def backward(ctx, gout):
        eng = ctx.eng
        opt = ctx.net._fused_optimizer()
        if opt is not None and not opt.materialize_grads:
            eng.pending_scale = gout.detach().to(torch.float32).reshape(1).contiguous()
            eng.acc_ready = True
            return (None, None, None, *([None] * len(ctx.params)))
        flat = eng.read_grads()
        gp = [g * gout for g in _split_flat(flat, ctx.params)]
        if opt is not None:
            eng.pending_scale = gout.detach().to(torch.float32).reshape(1).contiguous()
            eng.acc_ready = True
        return (None, None, None, *gp)


def _global_counts(counts, device):
    """Sum of per-rank row counts (host ints).  One small allreduce + host read; callers that know the global batch
    pass it instead (``loss.global_rows``)."""
    t = torch.tensor(list(counts), dtype=torch.int64, device=device)
    _dist.all_reduce_sum(t)
    return [int(v) for v in t.tolist()]


def compute_normal(model, pj):
    """grad_x f(pj), differentiable w.r.t. the weights (reference model/losses.py:283-296)."""
    net = _unwrap(model)
    pj.requires_grad = True  # the reference flips this in place (model/losses.py:285)
    return net.input_gradient(pj)[:, -3:]


def gradient(inputs, outputs):
    """reference executor/executordistributed.py:10-19 (generic autograd path through the drop-in module)."""
    d_points = torch.ones_like(outputs, requires_grad=False, device=outputs.device)
    return torch.autograd.grad(outputs=outputs, inputs=inputs, grad_outputs=d_points, create_graph=True, retain_graph=True,
                               only_inputs=True)[0][:, -3:]


class IGRLOSS(nn.Module):
    """reference model/losses.py:96-137"""

    def __init__(self, delta=0.1, tau=1, lambda_g=0.1, regularizer_threshold=1, data_term="pairwise"):
        super().__init__()
        self.delta = delta
        self.tau = tau
        self.lambda_g = lambda_g
        self.regularizer_threshold = regularizer_threshold
        if data_term not in ("pairwise", "pointwise"):
            raise ValueError("data_term must be 'pairwise' (reference-exact) or 'pointwise'")
        self.data_term = data_term
        self.last_terms = None  # (data, reg, eik) device tensor of the last call, for logging

    def __name__(self):
        return "IGRLOSS"

    def forward(self, x_batch, y_batch, model, epoch):
        net = _unwrap(model)
        eng = net._engine_for(x_batch)
        world = _dist.world_size()
        # {sum clamp(S), rows} of the whole batch, on the device: the pairwise data term needs the global label mean and
        # every term the global normaliser BEFORE the seeds exist.  Rows travel with the label sum, so unequal shards
        # (distributed.shard_rows with N % world != 0) use the right normaliser without any host value.
        y_batch = y_batch.detach()
        if y_batch.dtype != torch.float32 or not y_batch.is_contiguous():
            y_batch = y_batch.float().contiguous()
        stats = eng.label_stats(y_batch, self.delta)
        if world > 1:
            _dist.all_reduce_sum(stats)
        cfg = _lib.SdfnetIgrCfg(
            delta=float(self.delta), tau=float(self.tau), lambda_g=float(self.lambda_g),
            regularizer_threshold=float(self.regularizer_threshold),
            data_term=0 if self.data_term == "pairwise" else 1, mean_t=0.0, n_global=0,
            mean_t_dev=ctypes.c_void_p(0), label_stats_dev=ctypes.c_void_p(stats.data_ptr()),
        )
        eng.zero_grads()
        eng._sums.zero_()
        eng.train_step_igr(x_batch, y_batch, cfg, eng._sums)
        if world > 1:
            # ONE fp32 sum-allreduce: [gradient accumulators | loss sums] (replaces DataParallel's reduce into GPU 0)
            _dist.all_reduce_sum(eng.grad_buffer())
        out = eng.loss_value(0, tau=self.tau, lambda_g=self.lambda_g, data_term=0 if self.data_term == "pairwise" else 1)
        self.last_terms = out[1:4]
        return _FusedLoss.apply(out[0], net, eng, *net._ordered_params())


class WeightedSmoothL2Loss(nn.Module):
    """reference model/losses.py:50-69 (the loss the shipped ini selects).  First order only.  Fused like IGRLOSS: one
    forward that spills the activations, the seeds + loss sums (closed-form moments of the reference's [N,N] broadcast --
    "pairwise" -- or the per-point form), the adjoint sweep and the first term of the weight-gradient GEMMs (3F per point;
    the generic autograd route through ``model(x)`` runs the forward twice and pays ~30 small torch kernels per step).
    ``fused=False`` keeps that generic route (it also serves any model the fused engine does not)."""

    def __init__(self, weight_factor=0.5, delta=0.1, data_term="pairwise", fused=True):
        super().__init__()
        self.weight_factor = weight_factor
        self.delta = delta
        if data_term not in ("pairwise", "pointwise"):
            raise ValueError("data_term must be 'pairwise' (reference-exact) or 'pointwise'")
        self.data_term = data_term
        self.fused = bool(fused)

    def __name__(self):
        return "WeightedSmoothL2Loss"

    def forward(self, x_batch, y_batch, model, epoch):
        net = _unwrap(model)
        if not self.fused or not torch.is_grad_enabled():
            return self._generic(x_batch, y_batch, model)
        eng = net._engine_for(x_batch)
        world = _dist.world_size()
        y_batch = y_batch.detach()
        if y_batch.dtype != torch.float32 or not y_batch.is_contiguous():
            y_batch = y_batch.float().contiguous()
        dt = 0 if self.data_term == "pairwise" else 1
        stats = eng.wsl2_label_stats(y_batch, self.delta, self.weight_factor)
        if world > 1:
            _dist.all_reduce_sum(stats)
        eng.zero_grads()
        eng._sums.zero_()
        eng.train_step_wsl2(x_batch, y_batch, self.delta, self.weight_factor, dt, stats, eng._sums)
        if world > 1:
            _dist.all_reduce_sum(eng.grad_buffer())
        out = eng.loss_value(2, data_term=dt)
        return _FusedLoss.apply(out[0], net, eng, *net._ordered_params())

    def _generic(self, x_batch, y_batch, model):
        t = torch.clamp(y_batch[:, 0], -self.delta, self.delta)
        p = torch.clamp(model(x_batch), -self.delta, self.delta)[:, 0]
        w = 1.0 + self.weight_factor * torch.exp(-torch.abs(t))
        if self.data_term == "pairwise":
            # mean_{i,j} w_j (t_j - p_i)^2
            return (w * t * t).mean() - 2.0 * p.mean() * (w * t).mean() + (p * p).mean() * w.mean()
        return (w * (t - p) ** 2).mean()


def get_points(pc_input, local_sigma=0.0001, generator=None, on_device=False):
    """reference executor/executordistributed.py:108-123: a random third of the batch, perturbed by N(0, sigma^2).
    Default: CPU randperm / randn exactly like the reference (same seeds -> same points).  ``on_device=True`` draws
    the permutation and the noise with the device generator instead (no host work and no
    H2D copy per step; a different random stream, same distribution)."""
    sample_size = pc_input.shape[0]
    if on_device and pc_input.is_cuda:
        perm = torch.randperm(sample_size, device=pc_input.device)[: sample_size // 3]
        sample = pc_input[perm]
        return sample + torch.randn_like(sample) * local_sigma
    perm = torch.randperm(sample_size, generator=generator)[: sample_size // 3]
    sample = pc_input[perm.to(pc_input.device)]
    noise = torch.randn(sample.shape, generator=generator, dtype=sample.dtype).to(pc_input.device)
    return sample + noise * local_sigma


class PointCloudLoss(nn.Module):
    """The loss of the reference's point-cloud ("distributed") executor, executor/executordistributed.py:63-79:
    ``mean |f(x)| + lambda_g * mean((|grad f(x~)| - 1)^2)``, fused like IGRLOSS."""

    def __init__(self, lambda_g=0.1):
        super().__init__()
        self.lambda_g = lambda_g
        self.last_terms = None
        # (rows of x, rows of x~) over all ranks if the caller knows them (it sharded the batch); None: summed over ranks
        # with one small allreduce + host read per call
        self.global_rows = None

    def __name__(self):
        return "PointCloudLoss"

    def forward(self, x_batch, sampled_pts, model, epoch=0):
        net = _unwrap(model)
        eng = net._engine_for(x_batch)
        world = _dist.world_size()
        n, m = x_batch.shape[0], sampled_pts.shape[0]
        if self.global_rows is not None:
            n_global, m_global = (int(v) for v in self.global_rows)
        elif world > 1:
            n_global, m_global = _global_counts((n, m), x_batch.device)
        else:
            n_global, m_global = n, m
        eng.zero_grads()
        eng._sums.zero_()
        eng.train_step_pcd(x_batch, n_global, sampled_pts, m_global, self.lambda_g, eng._sums)
        if world > 1:
            _dist.all_reduce_sum(eng.grad_buffer())
        out = eng.loss_value(1, lambda_g=self.lambda_g, n_global=n_global, m_global=max(m_global, 1))
        self.last_terms = out[1:3]
        return _FusedLoss.apply(out[0], net, eng, *net._ordered_params())
