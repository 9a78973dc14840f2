"""CPU: host-side drop-in logic (no compute calls without a GPU) and the C-ABI library's exported symbols."""
import ctypes
import hashlib
import os
import pickle
import re

import numpy as np
import pytest
import torch

import sdf_oracle as orc
from helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import sdf_representation_b200 as pkg

    header = open(os.path.join(ROOT, "include", "sdfnet.h")).read()
    declared = set(re.findall(r"\b(sdfnet_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 18
    lib = ctypes.CDLL(pkg._lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sdfnet.h but not exported"
    assert declared == set(pkg._lib.SIGNATURES), "ctypes binding and header disagree"
    pkg._lib.load()


def test_create_without_gpu_fails_loudly():
    import sdf_representation_b200 as pkg

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = pkg._lib.load()
    desc = pkg._lib.SdfnetDesc(d_in=3, hidden=128, num_hidden=4, skip_layer=2, beta=100.0, dtype=1, precision=0, device=0)
    h = ctypes.c_void_p()
    rc = lib.sdfnet_create(ctypes.byref(desc), ctypes.byref(h))
    assert rc != 0 and len(lib.sdfnet_last_error()) > 0


def test_seed_identical_init_and_state_dict_layout():
    import sdf_representation_b200 as pkg

    d = load_golden("full_8x512_seed0.npz")
    torch.manual_seed(0)
    net = pkg.ImplicitNet(3, [512] * 8, (4,), 100.0, True)
    sd = net.state_dict()
    keys = list(sd.keys())
    assert keys == [f"lin{i}.{k}" for i in range(9) for k in ("weight", "bias")]
    assert sd["lin3.weight"].shape == (509, 512) and sd["lin8.weight"].shape == (1, 512) and sd["lin0.weight"].shape == (512, 3)
    assert sum(v.numel() for v in sd.values()) == 1_839_614
    h = hashlib.sha256()
    for v in sd.values():
        h.update(v.numpy().tobytes())
    assert h.hexdigest() == str(d["weights_sha256"]), "init is not bit-identical to the reference's for the same seed"
    assert net._get_name() == "ImplicitNet" and net.__name__() == "ImplicitNet"
    assert net.num_layers == 10 and net.skip_in == (4,)
    assert isinstance(net.activation, torch.nn.Softplus)


def test_relu_mode_and_cpu_input_raises():
    import sdf_representation_b200 as pkg

    net = pkg.ImplicitNet(3, [128] * 3, (), 0, False)
    assert isinstance(net.activation, torch.nn.ReLU)
    with pytest.raises(RuntimeError, match="no CPU"):
        net(torch.zeros(2, 3))


def test_checkpoint_layout_round_trip(tmp_path):
    import sdf_representation_b200 as pkg
    from sdf_representation_b200 import checkpoint

    torch.manual_seed(1)
    net = pkg.ImplicitNet(3, [64] * 4, (2,), 100.0, True)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    p1 = checkpoint.save_model(net, opt, [0.5], 3, 0.4, [0.6], str(tmp_path), best=True)
    p2 = checkpoint.save_model(net, opt, [0.5], 3, 0.4, [0.6], str(tmp_path), best=False)
    assert os.path.basename(p1) == "best_model.pkl" and os.path.basename(p2) == "model_epoch3.pkl"
    with open(p1, "rb") as f:
        data = pickle.load(f)
    assert list(data.keys()) == ["epoch", "loss_per_epoch", "best_val_loss", "val_loss_per_epoch", "model_state_dict", "optimizer_state_dict"]
    with open(p2, "rb") as f:
        assert list(pickle.load(f).keys()) == ["epoch", "model_state_dict"]
    net2 = pkg.ImplicitNet(3, [64] * 4, (2,), 100.0, True)
    net2, epoch = checkpoint.load_model(net2, None, str(tmp_path), best=False)
    assert epoch == 4
    for a, b in zip(net.state_dict().values(), net2.state_dict().values()):
        assert torch.equal(a, b)
    # 'module.' prefixed state (saved under DataParallel) loads into a bare module
    pref = {"module." + k: v for k, v in net.state_dict().items()}
    checkpoint.model_device_handler(net2, pref)
    comp = checkpoint.to_compatible_keys(net.state_dict())
    assert list(comp.keys())[0] == "layers.0.weight"


def test_configuration_reads_reference_keys():
    import sdf_representation_b200 as pkg

    cfg = pkg.Configuration(os.path.join(ROOT, "tests", "test.ini"))
    # same expectations as the reference's tests/test_config.py, on the corrected ini
    assert cfg.model.__class__.__name__ == "ImplicitNet"
    assert cfg.hidden_dim == 512 and cfg.num_hidden_layers == 8 and cfg.input_dim == 3
    assert cfg.skip_connection == () and cfg.beta == 0 and cfg.geometric_init is False
    assert cfg.loss.__class__.__name__ == "WeightedSmoothL2Loss"
    assert cfg.loss.weight_factor == 0.5 and cfg.loss.delta == 0.1
    assert cfg.lr == 0.00001 and cfg.epochs == 20000 and cfg.minepochs == 400 and cfg.batchsize == 4096
    assert cfg.uniform_points == 100000 and cfg.surface == 15 and cfg.narrowband == 15 and cfg.narrowband_width == 0.1
    assert cfg.train_test_split == 0.1 and cfg.ppo is False and cfg.cubesize == 256 and cfg.ppbatchsize == 6460000
    assert cfg.model._get_name() == "ImplicitNet" and cfg.loss._get_name() == "WeightedSmoothL2Loss"


def test_loss_protocol():
    import sdf_representation_b200 as pkg

    loss = pkg.IGRLOSS(delta=0.1, tau=1.0, lambda_g=0.1, regularizer_threshold=1.0)
    assert loss.__name__() == "IGRLOSS" and loss._get_name() == "IGRLOSS" and loss.lambda_g == 0.1
    with pytest.raises(TypeError):
        loss(torch.zeros(2, 3), torch.zeros(2, 4), torch.nn.Linear(3, 1), 0)


def test_get_points_shape_and_scale():
    import sdf_representation_b200 as pkg

    g = torch.Generator().manual_seed(0)
    x = torch.rand(100, 3, generator=g)
    s = pkg.get_points(x, generator=g)
    assert s.shape == (33, 3)
    # every perturbed point sits within a few sigma of one of the inputs
    dmin = torch.cdist(s, x).min(dim=1).values
    assert float(dmin.max()) < 1e-3


def test_slab_and_row_sharding_cover_exactly():
    from sdf_representation_b200 import distributed as D

    for n, w in [(1024, 8), (1000, 3), (7, 8), (128, 1)]:
        spans = [D.shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    assert D.slab_bounds(1024, 3, 8) == (384, 512)
    assert D.world_size() == 1 and D.rank() == 0


def test_export_flat_layout(tmp_path):
    """SDFW0001 flat export read by the native query tool (csrc/tools/sdf_query.cpp): header + per-layer arrays."""
    import struct

    import numpy as np
    import sdf_representation_b200 as pkg
    from sdf_representation_b200 import checkpoint

    torch.manual_seed(1)
    net = pkg.ImplicitNet(3, [64] * 3, (2,), 100.0, True)
    path = checkpoint.export_flat(net, str(tmp_path / "n.sdfw"))
    raw = open(path, "rb").read()
    assert raw[:8] == b"SDFW0001"
    d_in, hidden, nh, skip, beta = struct.unpack("<4if", raw[8:28])
    assert (d_in, hidden, nh, skip) == (3, 64, 3, 2) and abs(beta - 100.0) < 1e-6
    off = 28
    for i in range(4):
        lin = getattr(net, f"lin{i}")
        for t in (lin.weight, lin.bias):
            a = np.frombuffer(raw, dtype="<f4", count=t.numel(), offset=off)
            assert np.array_equal(a, t.detach().numpy().reshape(-1))
            off += 4 * t.numel()
    assert off == len(raw)
