"""Test infrastructure: a CPU stand-in for ``engine.Engine`` whose arithmetic is the oracle.  It lets the gloo tests run
the PRODUCT host code of the multi-GPU step (``losses.IGRLOSS.forward`` / ``PointCloudLoss.forward`` /
``WeightedSmoothL2Loss.forward``: label statistics,
the one sum-allreduce of [gradients | loss sums], the loss value, the autograd hand-over) without a GPU.  It mirrors the
engine methods those functions call and nothing else; it is never importable from the package."""
import torch

import sdf_oracle as orc


class FakeEngine:
    def __init__(self, net):
        self.net = net
        self.skip = tuple(net.skip_in)
        self.beta = float(net.beta)
        self.num_params = sum(p.numel() for p in net.parameters())
        self._sums = torch.zeros(8, dtype=torch.float64)
        self._stats = torch.zeros(2, dtype=torch.float64)
        self._gbuf = torch.zeros(self.num_params + 32, dtype=torch.float64)
        self.acc_ready = False
        self.pending_scale = None
        self.allreduce_calls = 0

    def _params(self):
        L = self.net.num_layers - 1
        W = [getattr(self.net, f"lin{i}").weight.detach().double() for i in range(L)]
        b = [getattr(self.net, f"lin{i}").bias.detach().double() for i in range(L)]
        return W, b

    def _add_grads(self, dW, db):
        flat = torch.cat([t.reshape(-1) for pair in zip(dW, db) for t in pair])
        self._gbuf[: self.num_params] += flat

    def _mirror(self):
        self._gbuf[self.num_params: self.num_params + 16:2] = self._sums  # "hi" slots; the lo slots stay 0 in fp64
        self._gbuf[self.num_params + 1: self.num_params + 16:2] = 0

    # ---- the Engine surface used by losses.py
    def zero_grads(self):
        self._gbuf.zero_()
        self.acc_ready = False

    def label_stats(self, y, delta):
        self._stats[0] = y[:, 0].double().clamp(-delta, delta).sum()
        self._stats[1] = y.shape[0]
        return self._stats

    def grad_buffer(self):
        return self._gbuf

    def train_step_igr(self, x, y, cfg, sums):
        W, b = self._params()
        x, y = x.double(), y.double()
        n_global = float(self._stats[1])
        mean_t = float(self._stats[0]) / n_global
        f, g = orc.forward_and_gradient(W, b, x, self.skip, self.beta)
        _, _, fbar, gbar = orc.igr_terms(f, g, y, cfg.delta, cfg.tau, cfg.lambda_g, cfg.regularizer_threshold,
                                         "pairwise" if cfg.data_term == 0 else "pointwise")
        n = x.shape[0]
        p = f.clamp(-cfg.delta, cfg.delta)
        t = y[:, 0].clamp(-cfg.delta, cfg.delta)
        inside = ((f >= -cfg.delta) & (f <= cfg.delta)).double()
        fbar = (2.0 / n_global) * (p - (mean_t if cfg.data_term == 0 else t)) * inside
        gbar = gbar * (n / n_global)
        dW, db = orc.param_grads(W, b, x, self.skip, self.beta, fbar, gbar)
        self._add_grads(dW, db)
        m = t.abs() < cfg.regularizer_threshold
        gn = g.norm(2, dim=-1)
        nn_ = y[:, 1:4].norm(2, dim=-1)
        cos = (y[:, 1:4] * g).sum(-1) / (nn_.clamp(min=1e-8) * gn.clamp(min=1e-8))
        eps = torch.full_like(f, 1e-8)
        sums += torch.stack([p.sum(), (p * p).sum(), t.sum(), (t * t).sum(), ((p - t) ** 2).sum(),
                             torch.where(m, (1 - cos) ** 2, eps).sum(), torch.where(m, (gn - 1) ** 2, eps).sum(),
                             torch.tensor(float(n), dtype=torch.float64)])
        self._mirror()

    def train_step_pcd(self, x, n_global, xp, m_global, lambda_g, sums):
        W, b = self._params()
        x, xp = x.double(), xp.double()
        f = orc.forward(W, b, x, self.skip, self.beta)
        dW1, db1 = orc.param_grads(W, b, x, self.skip, self.beta, torch.sign(f) / n_global, None)
        _, g = orc.forward_and_gradient(W, b, xp, self.skip, self.beta)
        gn = g.norm(2, dim=-1)
        gbar = (lambda_g / m_global) * 2 * ((gn - 1) / gn).unsqueeze(-1) * g
        dW2, db2 = orc.param_grads(W, b, xp, self.skip, self.beta, None, gbar)
        self._add_grads([a + c for a, c in zip(dW1, dW2)], [a + c for a, c in zip(db1, db2)])
        sums[0] += f.abs().sum()
        sums[1] += ((gn - 1) ** 2).sum()
        self._mirror()

    def wsl2_label_stats(self, y, delta, weight_factor):
        t = y[:, 0].double().clamp(-delta, delta)
        w = 1.0 + weight_factor * torch.exp(-t.abs())
        self._stats4 = torch.stack([w.sum(), (w * t).sum(), (w * t * t).sum(), torch.tensor(float(y.shape[0]), dtype=torch.float64)])
        return self._stats4

    def train_step_wsl2(self, x, y, delta, weight_factor, data_term, stats, sums):
        # model/losses.py:50-69 differentiated by hand: fbar = [|f| <= delta] (2/N)(p mean(w) - mean(w t))
        W, b = self._params()
        x, y = x.double(), y.double()
        N = float(stats[3])
        f = orc.forward(W, b, x, self.skip, self.beta)
        p = f.clamp(-delta, delta)
        t = y[:, 0].clamp(-delta, delta)
        w = 1.0 + weight_factor * torch.exp(-t.abs())
        inside = ((f >= -delta) & (f <= delta)).double()
        if data_term == 0:
            fbar = (2.0 / N) * (p * float(stats[0]) / N - float(stats[1]) / N) * inside
        else:
            fbar = (2.0 / N) * w * (p - t) * inside
        dW, db = orc.param_grads(W, b, x, self.skip, self.beta, fbar, None)
        self._add_grads(dW, db)
        sums += torch.stack([p.sum(), (p * p).sum(), (w * (t - p) ** 2).sum(), w.sum(), (w * t).sum(), (w * t * t).sum(),
                             torch.zeros((), dtype=torch.float64), torch.tensor(float(x.shape[0]), dtype=torch.float64)])
        self._mirror()

    def loss_value(self, kind, tau=0.0, lambda_g=0.0, data_term=0, n_global=0, m_global=0):
        s = self._gbuf[self.num_params: self.num_params + 16:2]
        if kind == 2:
            N = float(s[7])
            v = s[5] / N - 2 * (s[0] / N) * (s[4] / N) + (s[1] / N) * (s[3] / N) if data_term == 0 else s[2] / N
            z = torch.zeros((), dtype=torch.float64)
            return torch.stack([v, v, z, z])
        if kind == 0:
            N = float(n_global) if n_global > 0 else float(s[7])
            data = s[1] / N + s[3] / N - 2 * (s[0] / N) * (s[2] / N) if data_term == 0 else s[4] / N
            reg, eik = s[5] / N, s[6] / N
            return torch.stack([data + tau * reg + lambda_g * eik, data, reg, eik])
        surf, eik = s[0] / float(n_global), s[1] / float(m_global)
        return torch.stack([surf + lambda_g * eik, surf, eik, torch.zeros((), dtype=torch.float64)])

    def read_grads(self):
        return self._gbuf[: self.num_params].clone()
