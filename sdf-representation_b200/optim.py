"""Fused optimizer tail: drop-in for ``torch.optim.Adam(model.parameters(), lr)`` of the reference's training loops
(``executor/executor.py:160,177``, ``executor/executordistributed.py:59-83``).

``loss.backward()`` of the fused losses (losses.py) leaves the step's gradient in the engine's accumulators;
``FusedAdam.step()`` then runs TWO kernels of libsdfnet.so: unpack + Adam update of all 18 tensors in place, and the
re-pack of the 16-bit operand images the next step reads.  The generic path
(``torch.optim.Adam`` on ``p.grad``) costs 1 unpack + 18 multiplies + the foreach Adam kernels + 26 pack launches.

Gradients that arrive as ``p.grad`` (any loss that goes through the generic autograd path of the drop-in module, e.g.
``WeightedSmoothL2Loss``) are applied with the library's flat Adam kernel per tensor.  Same update rule, state layout
and ``state_dict()`` keys (``step``, ``exp_avg``, ``exp_avg_sq``) as ``torch.optim.Adam`` without amsgrad / weight decay.
"""
import weakref

import torch
import torch.nn as nn

from .engine import adam_step
from .networks import ImplicitNet


def _unwrap(model):
    m = model.module if isinstance(model, (nn.DataParallel, nn.parallel.DistributedDataParallel)) else model
    if not isinstance(m, ImplicitNet):
        raise TypeError(f"FusedAdam drives the drop-in ImplicitNet; got {type(m).__name__}")
    return m


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, materialize_grads=False):
        net = _unwrap(model)
        params = net._ordered_params()
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.net = net
        self.materialize_grads = bool(materialize_grads)
        self._steps = 0
        self._step_t = torch.tensor(0.0)  # one host tensor shared by every parameter's state["step"]
        self._flat_m = self._flat_v = self._flat_g = None
        net.__dict__["_fused_opt"] = weakref.ref(self)

    # ------------------------------------------------------------------ state (torch.optim.Adam layout)
    def _ensure_state(self):
        params = self.net._ordered_params()
        dev = params[0].device
        if self._flat_m is not None and self._flat_m.device == dev:
            return params
        n = sum(p.numel() for p in params)
        old = [(self.state[p].get("exp_avg"), self.state[p].get("exp_avg_sq")) if p in self.state else (None, None) for p in params]
        self._flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        self._flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        self._flat_g = torch.zeros(n, dtype=torch.float32, device=dev) if self.materialize_grads else None
        o = 0
        for p, (m0, v0) in zip(params, old):
            k = p.numel()
            mv, vv = self._flat_m[o:o + k].view(p.shape), self._flat_v[o:o + k].view(p.shape)
            if m0 is not None:
                mv.copy_(m0)
                vv.copy_(v0)
            self.state[p] = {"step": self._step_t, "exp_avg": mv, "exp_avg_sq": vv}
            o += k
        return params

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        steps = [int(st["step"]) for st in self.state.values() if "step" in st]
        self._steps = max(steps) if steps else 0
        self._step_t = torch.tensor(float(self._steps))
        self._flat_m = None  # re-flatten from the loaded per-tensor state on the next step
        self._ensure_state()

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        params = self._ensure_state()
        group = self.param_groups[0]
        lr, (b1, b2), eps = group["lr"], group["betas"], group["eps"]
        self._steps += 1
        eng = self.net._engines.get(params[0].device.index)
        if eng is not None and eng.acc_ready:
            eng.fused_adam_step(params[0::2], params[1::2], self._flat_m, self._flat_v, lr, b1, b2, eps, self._steps,
                                grad_scale_dev=eng.pending_scale, grads_out=self._flat_g)
            eng.acc_ready = False
            # the library re-packed the images from the updated values; the raw-pointer update does not bump the
            # parameters' version counters, so the cache key is unchanged and now names the NEW values
            eng._packed_key = self.net._param_key(params)
            if self.materialize_grads:
                o = 0
                for p in params:
                    p.grad = self._flat_g[o:o + p.numel()].view(p.shape)
                    o += p.numel()
        else:
            o = 0
            touched = False
            for p in params:
                k = p.numel()
                if p.grad is not None:
                    g = p.grad.detach().to(torch.float32).contiguous().view(-1)
                    adam_step(p.detach().view(-1), g, self._flat_m[o:o + k], self._flat_v[o:o + k], lr, self._steps, b1, b2, eps)
                    touched = True
                o += k
            if not touched:
                raise RuntimeError("FusedAdam.step(): no gradient available (call loss.backward() first)")
            self.net.invalidate_packed()
        self._step_t.fill_(float(self._steps))
        return loss
