"""Per-device engine: owns one ``sdfnet_t`` handle of libsdfnet.so and the scratch buffers its calls need.

This is plumbing only: PyTorch supplies device memory and the CUDA stream, every computation happens in the
hand-written sm_100a kernels behind the C ABI of ``include/sdfnet.h``.  There is no CPU or eager-PyTorch fallback:
non-CUDA tensors raise.
"""
import ctypes
import math
import os

import torch

from . import _lib

DTYPE_IDS = {"bf16": 0, "fp16": 1}
PRECISION_IDS = {"fast": 0, "parity": 1}

# cap of the spill workspace of one training chunk: SDFNET_WS_GB (GiB) if set, else 60 % of the device memory that is
# free when the workspace is (re)allocated, at most 64 GiB -- larger batches are chunked by the library
_WS_CAP_ENV = os.environ.get("SDFNET_WS_GB")
_WS_CAP_MAX = 64 << 30


def _ws_cap_bytes(device):
    if _WS_CAP_ENV:
        return int(float(_WS_CAP_ENV) * (1 << 30))
    free, _total = torch.cuda.mem_get_info(device)
    return max(min(int(0.6 * free), _WS_CAP_MAX), 1 << 26)


class _RawCudaBuffer:
    """Zero-copy torch view of library-owned device memory (``__cuda_array_interface__``)."""

    def __init__(self, ptr, n, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_points(x, cols, name):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the SDF hot path has no CPU fallback")
    if x.dim() != 2 or x.shape[1] != cols:
        raise RuntimeError(f"{name} must have shape [N, {cols}] (got {tuple(x.shape)})")
    if x.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32 (got {x.dtype})")
    return x.detach().contiguous()


class Engine:
    """One network description bound to one CUDA device."""

    def __init__(self, hidden, num_hidden, skip_layer, beta, device, dtype="fp16", precision="fast", d_in=3):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("sdfnet engines exist on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        desc = _lib.SdfnetDesc(
            d_in=int(d_in),
            hidden=int(hidden),
            num_hidden=int(num_hidden),
            skip_layer=int(skip_layer),
            beta=float(beta),
            dtype=DTYPE_IDS[dtype],
            precision=PRECISION_IDS[precision],
            device=self.device.index,
        )
        self.hidden, self.num_hidden, self.skip_layer, self.beta = int(hidden), int(num_hidden), int(skip_layer), float(beta)
        self.dtype, self.precision = dtype, precision
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sdfnet_create(ctypes.byref(desc), ctypes.byref(h)), "sdfnet_create")
        self.handle = h
        self.num_params = int(self.lib.sdfnet_num_params(self.handle))
        self._ws = None
        self._ws_capped = False
        self._packed_key = None
        self.launches = 0  # kernels launched through this engine (bench.py reports it as gpu_launches)
        # fused step state: persistent loss-sum / label-stat / loss-value buffers, the gradient buffer view, and the
        # hand-over of a finished backward to FusedAdam (optim.py)
        self._sums = torch.zeros(8, dtype=torch.float64, device=self.device)
        self._stats = torch.zeros(2, dtype=torch.float64, device=self.device)
        self._stats4 = torch.zeros(4, dtype=torch.float64, device=self.device)
        self._gbuf = None
        self.acc_ready = False
        self.pending_scale = None
        self.deterministic = False

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.sdfnet_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ instrumentation
    def launch_count(self):
        """Kernels of libsdfnet.so launched through this engine so far (counted inside the library)."""
        return int(self.lib.sdfnet_launch_count(self.handle))

    def profile(self, on=True):
        _lib.check(self.lib.sdfnet_profile_enable(self.handle, 1 if on else 0), "sdfnet_profile_enable")

    def profile_read(self):
        """{class: (milliseconds, launches)} measured with CUDA events on the launch stream."""
        ms = (ctypes.c_double * 4)()
        n = (ctypes.c_int64 * 4)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sdfnet_profile_read(self.handle, ms, n), "sdfnet_profile_read")
        names = ("chain_fwd_rev", "chain_tan_adj", "dw_gemm", "other")
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(names)}

    # ------------------------------------------------------------------ weights
    def pack(self, weights, biases, key=None):
        """Re-pack lin{i}.weight / lin{i}.bias into the UMMA operand images (skipped if ``key`` is unchanged)."""
        if key is not None and key == self._packed_key:
            return
        n = self.num_hidden + 1
        if len(weights) != n or len(biases) != n:
            raise RuntimeError(f"expected {n} weight and bias tensors")
        keep = []
        wp = (ctypes.c_void_p * n)()
        bp = (ctypes.c_void_p * n)()
        for i, (w, b) in enumerate(zip(weights, biases)):
            for t in (w, b):
                if not t.is_cuda or t.device != self.device or t.dtype != torch.float32:
                    raise RuntimeError("parameters must be float32 CUDA tensors on the engine's device (no CPU fallback)")
            w = w.detach().contiguous()
            b = b.detach().contiguous()
            keep += [w, b]
            wp[i] = w.data_ptr()
            bp[i] = b.data_ptr()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sdfnet_pack_weights(self.handle, wp, bp, _stream(self.device)), "sdfnet_pack_weights")
        # the library records the parameter pointers and packs a weight layout on first use (sdfnet.h): keep the tensors
        # (possibly contiguous temporaries) alive until the next pack
        self._pack_keep = keep
        self.launches += 2 * n + (n - 1) + 1
        self._packed_key = key

    # ------------------------------------------------------------------ scratch
    def release_workspace(self):
        """Give the spill workspace back to the caching allocator (e.g. before a large grid inference)."""
        self._ws = None
        self._ws_capped = False

    def workspace(self, n_points, mode):
        need = int(self.lib.sdfnet_workspace_bytes(self.handle, int(n_points), int(mode)))
        if self._ws is not None and (self._ws.numel() >= need or self._ws_capped):
            return self._ws
        held = 0 if self._ws is None else self._ws.numel()
        self._ws = None  # returned to the caching allocator before the cap is computed from the free memory
        floor = int(self.lib.sdfnet_workspace_bytes(self.handle, 256, int(mode))) + 4096  # 256 points = an even tile count
        target = max(min(need, _ws_cap_bytes(self.device)), floor, held)
        self._ws_capped = target < need  # the library chunks the batch; do not re-allocate on every call
        self._ws = torch.empty(target, dtype=torch.uint8, device=self.device)
        return self._ws

    def _chunks(self, n, mode):
        per_256 = int(self.lib.sdfnet_workspace_bytes(self.handle, 256, int(mode))) - 1024
        pts = max(1, (self._ws.numel() - 1024) // per_256) * 256
        return max(1, math.ceil(n / pts))

    # ------------------------------------------------------------------ inference
    def forward(self, x):
        x = _check_points(x, 3, "points")
        n = x.shape[0]
        out = torch.empty(n, dtype=torch.float32, device=x.device)
        if n:
            with torch.cuda.device(self.device):
                _lib.check(self.lib.sdfnet_forward(self.handle, _ptr(x), n, _ptr(out), _stream(self.device)), "sdfnet_forward")
            self.launches += 1
        return out

    def forward_grad(self, x):
        x = _check_points(x, 3, "points")
        n = x.shape[0]
        f = torch.empty(n, dtype=torch.float32, device=x.device)
        g = torch.empty(n, 3, dtype=torch.float32, device=x.device)
        if n:
            ws = self.workspace(n, 0)
            with torch.cuda.device(self.device):
                _lib.check(
                    self.lib.sdfnet_forward_grad(self.handle, _ptr(x), n, _ptr(f), _ptr(g), _ptr(ws), ws.numel(), _stream(self.device)),
                    "sdfnet_forward_grad",
                )
            self.launches += self._chunks(n, 0)
        return f, g

    def grid_eval(self, axis, first, count, out=None):
        if not axis.is_cuda or axis.dtype != torch.float32 or axis.dim() != 1:
            raise RuntimeError("axis must be a 1-D float32 CUDA tensor (no CPU fallback)")
        axis = axis.contiguous()
        if out is None:
            out = torch.empty(int(count), dtype=torch.float32, device=axis.device)
        elif out.numel() < count or out.dtype != torch.float32 or not out.is_contiguous():
            raise RuntimeError("grid_eval: output buffer too small or not contiguous float32")
        if count:
            with torch.cuda.device(self.device):
                _lib.check(
                    self.lib.sdfnet_grid_eval(self.handle, _ptr(axis), int(axis.numel()), int(first), int(count), _ptr(out), _stream(self.device)),
                    "sdfnet_grid_eval",
                )
            self.launches += math.ceil(count / (1 << 30))
        return out

    def octree_level(self, keys, level, band, want_flags=True):
        if not keys.is_cuda or keys.dtype not in (torch.int64, torch.uint64):
            raise RuntimeError("keys must be an int64 CUDA tensor of Morton codes (no CPU fallback)")
        keys = keys.contiguous()
        n = keys.numel()
        sdf = torch.empty(n, dtype=torch.float32, device=keys.device)
        flags = torch.empty(n, dtype=torch.uint8, device=keys.device) if want_flags else None
        if n:
            with torch.cuda.device(self.device):
                _lib.check(
                    self.lib.sdfnet_octree_level(self.handle, _ptr(keys), n, int(level), float(band), _ptr(sdf), _ptr(flags), _stream(self.device)),
                    "sdfnet_octree_level",
                )
            self.launches += 2 if want_flags else 1
        return sdf, flags

    # ------------------------------------------------------------------ gradients
    def zero_grads(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sdfnet_zero_grads(self.handle, _stream(self.device)), "sdfnet_zero_grads")
        self.acc_ready = False

    def grad_buffer(self):
        """Zero-copy fp32 view of [gradient accumulators | loss-sum tail] (library-owned): the one message a data-parallel
        step all-reduces (include/sdfnet.h sdfnet_grad_buffer)."""
        if self._gbuf is None:
            ptr = ctypes.c_void_p()
            n = ctypes.c_int64()
            _lib.check(self.lib.sdfnet_grad_buffer(self.handle, ctypes.byref(ptr), ctypes.byref(n)), "sdfnet_grad_buffer")
            self._gbuf = torch.as_tensor(_RawCudaBuffer(ptr.value, n.value), device=self.device)
        return self._gbuf

    def set_deterministic(self, on=True):
        """Fixed-order reductions everywhere (weight-gradient split-K through private slots + an ordered second pass, loss
        sums through per-block partials): bit-identical results from run to run, like the reference's SGEMM path.  Costs a
        135 MB memset + reduce per step (~0.1 ms at 1M points)."""
        _lib.check(self.lib.sdfnet_set_deterministic(self.handle, 1 if on else 0), "sdfnet_set_deterministic")
        self.deterministic = bool(on)

    def label_stats(self, y, delta):
        """device double[2] = {sum clamp(y[:,0], +-delta), rows} of this shard (persistent buffer, overwritten per call)."""
        if self.deterministic:
            # torch's sum is a fixed-shape tree: the same bits on every run
            self._stats[0] = y[:, 0].clamp(-delta, delta).sum(dtype=torch.float64)
            self._stats[1] = float(y.shape[0])
            return self._stats
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sdfnet_label_stats(_ptr(y), int(y.shape[0]), float(delta), _ptr(self._stats), _stream(self.device)),
                       "sdfnet_label_stats")
        self.launches += 1
        return self._stats

    def wsl2_label_stats(self, y, delta, weight_factor):
        """device double[4] = {sum w, sum w t, sum w t^2, rows} of this shard's labels (model/losses.py:50-69; persistent buffer)."""
        if self.deterministic:
            t = y[:, 0].clamp(-delta, delta).double()
            w = 1.0 + weight_factor * torch.exp(-t.abs())
            self._stats4[0], self._stats4[1], self._stats4[2] = w.sum(), (w * t).sum(), (w * t * t).sum()
            self._stats4[3] = float(y.shape[0])
            return self._stats4
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sdfnet_wsl2_label_stats(_ptr(y), int(y.shape[0]), float(delta), float(weight_factor), _ptr(self._stats4),
                                                        _stream(self.device)), "sdfnet_wsl2_label_stats")
        self.launches += 1
        return self._stats4

    def loss_value(self, kind, tau=0.0, lambda_g=0.0, data_term=0, n_global=0, m_global=0):
        """float32[4] = {total, term0, term1, term2} from the (all-reduced) tail of the gradient buffer."""
        out = torch.empty(4, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sdfnet_loss_value(self.handle, int(kind), float(tau), float(lambda_g), int(data_term), int(n_global),
                                                  int(m_global), _ptr(out), _stream(self.device)), "sdfnet_loss_value")
        self.launches += 1
        return out

    def fused_adam_step(self, weights, biases, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, grad_scale=1.0, grad_scale_dev=None,
                        grads_out=None):
        """Adam on the accumulated gradient + re-pack of the operand images, 2 kernels (include/sdfnet.h)."""
        n = self.num_hidden + 1
        wp = (ctypes.c_void_p * n)(*[w.data_ptr() for w in weights])
        bp = (ctypes.c_void_p * n)(*[b.data_ptr() for b in biases])
        with torch.cuda.device(self.device):
            _lib.check(
                self.lib.sdfnet_fused_adam_step(self.handle, wp, bp, _ptr(exp_avg), _ptr(exp_avg_sq), float(lr), float(beta1), float(beta2),
                                                float(eps), int(step), float(grad_scale), _ptr(grad_scale_dev), _ptr(grads_out),
                                                _stream(self.device)),
                "sdfnet_fused_adam_step",
            )
        self._pack_keep = list(weights) + list(biases)
        self.launches += 2

    def read_grads(self, out=None, scale=1.0):
        if out is None:
            out = torch.empty(self.num_params, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sdfnet_read_grads(self.handle, _ptr(out), float(scale), _stream(self.device)), "sdfnet_read_grads")
        self.launches += 1
        return out

    def backward(self, x, fbar=None, gbar=None):
        """Accumulate dL/dparams for seeds fbar = dL/df [N] and gbar = dL/d(grad_x f) [N,3]."""
        x = _check_points(x, 3, "points")
        n = x.shape[0]
        if n == 0 or (fbar is None and gbar is None):
            return
        if fbar is not None:
            fbar = fbar.detach().reshape(-1).to(torch.float32).contiguous()
            if fbar.numel() != n:
                raise RuntimeError("fbar must have one entry per point")
        if gbar is not None:
            gbar = _check_points(gbar, 3, "gbar")
            if gbar.shape[0] != n:
                raise RuntimeError("gbar must have one row per point")
        ws = self.workspace(n, 1)
        with torch.cuda.device(self.device):
            _lib.check(
                self.lib.sdfnet_backward(self.handle, _ptr(x), n, _ptr(fbar), _ptr(gbar), _ptr(ws), ws.numel(), _stream(self.device)),
                "sdfnet_backward",
            )
        self.launches += 4 * self._chunks(n, 1)

    def train_step_igr(self, x, y, cfg, loss_sums):
        x = _check_points(x, 3, "x_batch")
        y = _check_points(y, 4, "y_batch")
        if y.shape[0] != x.shape[0]:
            raise RuntimeError("x_batch and y_batch must have the same number of rows")
        n = x.shape[0]
        if n == 0:
            return
        ws = self.workspace(n, 1)
        with torch.cuda.device(self.device):
            _lib.check(
                self.lib.sdfnet_train_step_igr(
                    self.handle, _ptr(x), _ptr(y), n, ctypes.byref(cfg), _ptr(loss_sums), _ptr(ws), ws.numel(), _stream(self.device)
                ),
                "sdfnet_train_step_igr",
            )
        self.launches += 5 * self._chunks(n, 1)
        self.acc_ready = False

    def train_step_pcd(self, x, n_global, xp, m_global, lambda_g, loss_sums):
        x = _check_points(x, 3, "surface points")
        xp = _check_points(xp, 3, "perturbed points")
        n, m = x.shape[0], xp.shape[0]
        ws = self.workspace(max(n, m, 1), 1)
        with torch.cuda.device(self.device):
            _lib.check(
                self.lib.sdfnet_train_step_pcd(
                    self.handle, _ptr(x), n, int(n_global), _ptr(xp), m, int(m_global), float(lambda_g), _ptr(loss_sums), _ptr(ws),
                    ws.numel(), _stream(self.device),
                ),
                "sdfnet_train_step_pcd",
            )
        self.launches += 5 * (self._chunks(max(n, 1), 1) + self._chunks(max(m, 1), 1))


    def train_step_wsl2(self, x, y, delta, weight_factor, data_term, stats, loss_sums):
        """One fused WeightedSmoothL2Loss step (include/sdfnet.h sdfnet_train_step_wsl2); stats = the all-reduced
        wsl2_label_stats of the whole batch."""
        x = _check_points(x, 3, "x_batch")
        y = _check_points(y, 4, "y_batch")
        if y.shape[0] != x.shape[0]:
            raise ValueError("x_batch and y_batch must have the same number of rows")
        n = x.shape[0]
        ws = self.workspace(max(n, 1), 1)
        with torch.cuda.device(self.device):
            _lib.check(
                self.lib.sdfnet_train_step_wsl2(
                    self.handle, _ptr(x), _ptr(y), n, float(delta), float(weight_factor), int(data_term), _ptr(stats), _ptr(loss_sums),
                    _ptr(ws), ws.numel(), _stream(self.device),
                ),
                "sdfnet_train_step_wsl2",
            )
        self.launches += 5 * self._chunks(max(n, 1), 1)


def adam_step(params, grads, exp_avg, exp_avg_sq, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam(lr).step() on a flat fp32 parameter vector (executor/executor.py:160,177)."""
    lib = _lib.load()
    for t in (params, grads, exp_avg, exp_avg_sq):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("adam_step operates on contiguous float32 CUDA tensors (no CPU fallback)")
    with torch.cuda.device(params.device):
        _lib.check(
            lib.sdfnet_adam_step(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq), params.numel(), float(lr), float(beta1),
                                 float(beta2), float(eps), int(step), _stream(params.device)),
            "sdfnet_adam_step",
        )
