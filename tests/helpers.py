"""Shared helpers of the test-suite (test infrastructure)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def golden_params(d, dtype=torch.float64):
    L = int(d["num_hidden"])
    W = [torch.from_numpy(d[f"lin{i}_weight"]).to(dtype) for i in range(L + 1)]
    b = [torch.from_numpy(d[f"lin{i}_bias"]).to(dtype) for i in range(L + 1)]
    return W, b


def skip_of(d):
    s = int(d["skip"])
    return (s,) if s else ()


def golden_grads(d, prefix):
    L = int(d["num_hidden"])
    return ([torch.from_numpy(d[f"{prefix}lin{i}_weight"]) for i in range(L + 1)],
            [torch.from_numpy(d[f"{prefix}lin{i}_bias"]) for i in range(L + 1)])


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / max(float(b.norm()), 1e-30))
