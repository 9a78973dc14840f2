"""Drop-in for the reference's ``model/networks.py:8-112`` ``ImplicitNet``.

Same constructor (positional order ``d_in, dims, skip_in, beta, geometric_init, radius_init, lipsitch`` as called by
``configgen/config_reader.py:35``), same sub-module / parameter names (``lin0..lin{L}``, each an ``nn.Linear``), same
RNG call order (``nn.Linear`` default init, then the geometric-init overwrite, ``model/networks.py:62-74``) so that
``torch.manual_seed(s)`` gives bit-identical initial weights, same ``state_dict`` layout.  ``forward`` does not run
PyTorch ops: it calls the fused layer-chain kernels of libsdfnet.so through the C ABI.  The op is differentiable to
second order in the way the reference uses it (``model/losses.py:283-296``,
``executor/executordistributed.py:10-19``): ``autograd.grad(f, x, create_graph=True)`` followed by a backward into the
weights.  There is no CPU fallback: CPU tensors raise.
"""
import numpy as np
import torch
import torch.nn as nn

from .engine import Engine

# keyword-only knobs that do not exist in the reference; the defaults keep the reference's numerics contract
_DEFAULT_DTYPE = "fp16"
_DEFAULT_PRECISION = "parity"


def _split_flat(flat, params):
    out, o = [], 0
    for p in params:
        n = p.numel()
        out.append(flat[o:o + n].view(p.shape))
        o += n
    return out


class _SdfEval(torch.autograd.Function):
    """f = net(x).  backward: dL/dx = fbar * grad_x f (differentiable), dL/dparams via the adjoint sweep + dW GEMMs."""

    @staticmethod
    def forward(ctx, net, x, *params):
        eng = net._engine_for(x)
        f = eng.forward(x)
        ctx.net = net
        ctx.save_for_backward(x, *params)
        return f.unsqueeze(-1)

    @staticmethod
    def backward(ctx, fbar):
        net = ctx.net
        x, *params = ctx.saved_tensors
        need_x = ctx.needs_input_grad[1]
        need_p = any(ctx.needs_input_grad[2:])
        gx = None
        if need_x:
            # differentiable when called under create_graph=True (reference model/losses.py:289-295)
            gx = _SdfInputGrad.apply(net, x, fbar, *params)
        gp = [None] * len(params)
        if need_p:
            eng = net._engine_for(x)
            eng.zero_grads()
            eng.backward(x, fbar=fbar.reshape(-1), gbar=None)
            gp = _split_flat(eng.read_grads(), params)
            gp = [g if need else None for g, need in zip(gp, ctx.needs_input_grad[2:])]
        return (None, gx, *gp)


class _SdfInputGrad(torch.autograd.Function):
    """xbar = fbar * grad_x f(x; params).  backward (the reference's double backward): given G = dL/dxbar,
    dL/dfbar = (grad_x f . G), dL/dparams = second-order sweep seeded with gbar = fbar * G.  The Hessian term
    dL/dx is not produced (the reference never uses x.grad)."""

    @staticmethod
    def forward(ctx, net, x, fbar, *params):
        eng = net._engine_for(x)
        _, g = eng.forward_grad(x)
        ctx.net = net
        ctx.save_for_backward(x, fbar, g, *params)
        return g * fbar.reshape(-1, 1)

    @staticmethod
    def backward(ctx, G):
        net = ctx.net
        x, fbar, g, *params = ctx.saved_tensors
        g_fbar = None
        if ctx.needs_input_grad[2]:
            g_fbar = (g * G).sum(-1, keepdim=True).reshape(fbar.shape)
        gp = [None] * len(params)
        if any(ctx.needs_input_grad[3:]):
            eng = net._engine_for(x)
            eng.zero_grads()
            eng.backward(x, fbar=None, gbar=(G * fbar.reshape(-1, 1)).contiguous())
            gp = _split_flat(eng.read_grads(), params)
            gp = [t if need else None for t, need in zip(gp, ctx.needs_input_grad[3:])]
        return (None, None, g_fbar, *gp)


class ImplicitNet(nn.Module):
    """Skip-in, geometric-init fully connected SDF network (reference ``model/networks.py:8-112``)."""

    def __init__(
        self,
        d_in,
        dims,
        skip_in=(),
        beta=100,
        geometric_init=True,
        radius_init=1,
        lipsitch=False,
        *,
        precision=_DEFAULT_PRECISION,
        dtype=_DEFAULT_DTYPE,
        deterministic=False,
    ):
        super().__init__()
        hidden = list(dims)
        dims = [d_in] + hidden + [1]
        self.num_layers = len(dims)
        self.skip_in = skip_in
        # (the reference prints skip_in here, model/networks.py:50; a drop-in keeps stdout clean for its callers)
        self.lipsitch = lipsitch
        self.d_in = d_in
        self.beta = beta
        self.precision = precision
        self.dtype = dtype
        self.deterministic = bool(deterministic)  # bit-identical gradients from run to run (Engine.set_deterministic)
        for layer in range(0, self.num_layers - 1):
            out_dim = dims[layer + 1] - d_in if layer + 1 in skip_in else dims[layer + 1]
            lin = nn.Linear(dims[layer], out_dim)
            if geometric_init:
                if layer == self.num_layers - 2:
                    torch.nn.init.normal_(lin.weight, mean=np.sqrt(np.pi) / np.sqrt(dims[layer]), std=0.00001)
                    torch.nn.init.constant_(lin.bias, -radius_init)
                else:
                    torch.nn.init.constant_(lin.bias, 0.0)
                    torch.nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
            setattr(self, "lin" + str(layer), lin)
        self.activation = nn.Softplus(beta=beta) if beta > 0 else nn.ReLU()
        # kernel-side description (validated when the first CUDA batch arrives)
        self._hidden = hidden
        object.__setattr__(self, "_engines", {})
        object.__setattr__(self, "_fused_opt", None)  # weakref to the FusedAdam driving this network (optim.py)
        object.__setattr__(self, "_pack_epoch", 0)    # bumped by invalidate_packed()

    def __name__(self):
        return "ImplicitNet"

    def __getstate__(self):
        # engines hold raw library handles: never pickled / deep-copied, rebuilt lazily
        state = dict(self.__dict__)
        state["_engines"] = {}
        state["_fused_opt"] = None
        return state

    def _fused_optimizer(self):
        ref = self.__dict__.get("_fused_opt")
        return ref() if ref is not None else None

    def invalidate_packed(self):
        """Force a re-pack of the 16-bit operand images on the next launch.  The packed-weight cache is keyed on each
        parameter's (data_ptr, version counter); writes that bypass the version counter (``p.data.copy_()``, a collective
        on ``p.data``, a raw-pointer update from another library) must call this."""
        object.__setattr__(self, "_pack_epoch", self.__dict__.get("_pack_epoch", 0) + 1)

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_packed()
        return out

    # ------------------------------------------------------------------ engine plumbing
    def _ordered_params(self):
        ps = []
        for layer in range(self.num_layers - 1):
            lin = getattr(self, "lin" + str(layer))
            ps += [lin.weight, lin.bias]
        return ps

    def _describe(self):
        hidden = self._hidden
        if self.d_in != 3:
            raise RuntimeError(f"sdfnet kernels support d_in = 3 only (got {self.d_in}); there is no fallback path")
        if len(set(hidden)) != 1:
            raise RuntimeError(f"sdfnet kernels need equal hidden widths (got {hidden}); there is no fallback path")
        skip = tuple(self.skip_in)
        if len(skip) > 1:
            raise RuntimeError(f"sdfnet kernels support at most one skip layer (got {skip}); there is no fallback path")
        return hidden[0], len(hidden), (skip[0] if skip else 0)

    def _engine_for(self, x):
        if not x.is_cuda:
            raise RuntimeError(
                "ImplicitNet (sdf-representation_b200) runs on CUDA tensors only: the hot path is hand-written sm_100a "
                "code with no CPU or eager-PyTorch fallback"
            )
        params = self._ordered_params()
        if params[0].device != x.device:
            raise RuntimeError(f"input is on {x.device} but the parameters are on {params[0].device}")
        key = x.device.index
        eng = self._engines.get(key)
        if eng is None:
            H, L, skip = self._describe()
            eng = Engine(H, L, skip, float(self.beta) if self.beta > 0 else 0.0, x.device, dtype=self.dtype, precision=self.precision)
            if getattr(self, "deterministic", False):
                eng.set_deterministic(True)
            self._engines[key] = eng
        eng.pack(params[0::2], params[1::2], key=self._param_key(params))
        return eng

    def _param_key(self, params=None):
        params = self._ordered_params() if params is None else params
        return (self.__dict__.get("_pack_epoch", 0),) + tuple((p.data_ptr(), p._version) for p in params)

    def engine(self, device=None):
        """The engine bound to ``device`` (default: the parameters' device), with current weights packed."""
        p = self._ordered_params()[0]
        probe = torch.empty(0, 3, device=p.device if device is None else device)
        return self._engine_for(probe)

    def gpu_launches(self):
        return sum(e.launches for e in self._engines.values())

    # ------------------------------------------------------------------ nn.Module protocol
    def forward(self, input):
        if input.dim() != 2 or input.shape[-1] != self.d_in:
            raise RuntimeError(f"expected input of shape [N, {self.d_in}], got {tuple(input.shape)}")
        x = input if input.dtype == torch.float32 else input.float()
        return _SdfEval.apply(self, x, *self._ordered_params())

    def sdf_and_gradient(self, input):
        """(f [N,1], grad_x f [N,3]) in one fused pass, without autograd graph
        (ops/conversion_test/test_loading.cpp:62,99 equivalent)."""
        eng = self._engine_for(input)
        f, g = eng.forward_grad(input)
        return f.unsqueeze(-1), g

    def input_gradient(self, input):
        """grad_x f [N,3], differentiable w.r.t. the parameters (reference model/losses.py:283-296 compute_normal)."""
        ones = torch.ones(input.shape[0], 1, device=input.device, dtype=torch.float32)
        return _SdfInputGrad.apply(self, input, ones, *self._ordered_params())
