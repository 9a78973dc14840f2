"""CPU: the oracle restatement against golden vectors produced by the unmodified reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

import sdf_oracle as orc
from helpers import golden_grads, golden_params, load_golden, rel_err, skip_of

CASES = ["softplus_skip_128x4.npz", "relu_tanh_128x3.npz"]


@pytest.mark.parametrize("name", CASES)
def test_forward_and_gradient_match_reference(name):
    d = load_golden(name)
    W, b = golden_params(d)
    x = torch.from_numpy(d["x"]).double()
    f, g = orc.forward_and_gradient(W, b, x, skip_of(d), float(d["beta"]))
    # the reference ran in fp32: its own rounding error is ~1e-6; reference python<->libtorch agreed to 4.8e-7
    assert np.abs(f.numpy() - d["f"]).max() < 5e-6
    assert np.abs(g.numpy() - d["g"]).max() < 2e-5


@pytest.mark.parametrize("name", CASES)
def test_igr_loss_and_grads_match_reference(name):
    d = load_golden(name)
    W, b = golden_params(d)
    x = torch.from_numpy(d["x"]).double()
    y = torch.from_numpy(d["y"]).double()
    loss, _, dW, db = orc.igr_loss_and_grads(W, b, x, y, skip_of(d), float(d["beta"]), 0.1, 1.0, 0.1, 1.0, "pairwise")
    assert abs(float(loss) - float(d["igr_loss"])) < 2e-6 * max(1.0, abs(float(d["igr_loss"])))
    gW, gb = golden_grads(d, "igr_grad_")
    for l in range(len(W)):
        assert rel_err(dW[l], gW[l]) < 2e-4, f"dW{l}"
        assert rel_err(db[l], gb[l]) < 2e-4, f"db{l}"


@pytest.mark.parametrize("name", CASES)
def test_wsl2_loss_and_grads_match_reference(name):
    d = load_golden(name)
    W, b = golden_params(d)
    x = torch.from_numpy(d["x"]).double()
    y = torch.from_numpy(d["y"]).double()
    loss, dW, db = orc.wsl2_loss_and_grads(W, b, x, y, skip_of(d), float(d["beta"]), 0.5, 0.1, "pairwise")
    assert abs(float(loss) - float(d["wsl2_loss"])) < 2e-6 * max(1.0, abs(float(d["wsl2_loss"])))
    gW, gb = golden_grads(d, "wsl2_grad_")
    for l in range(len(W)):
        assert rel_err(dW[l], gW[l]) < 2e-4, f"dW{l}"
        assert rel_err(db[l], gb[l]) < 2e-4 or float(gb[l].norm()) < 1e-12, f"db{l}"


def test_pcd_loss_and_grads_match_reference():
    d = load_golden("softplus_skip_128x4.npz")
    W, b = golden_params(d)
    x = torch.from_numpy(d["x"]).double()
    xp = torch.from_numpy(d["xp"]).double()
    loss, (surf, eik), dW, db = orc.pcd_loss_and_grads(W, b, x, xp, skip_of(d), float(d["beta"]), 0.1)
    assert abs(float(loss) - float(d["pcd_loss"])) < 2e-6
    assert abs(float(surf) - d["pcd_terms"][0]) < 2e-6 and abs(float(eik) - d["pcd_terms"][1]) < 2e-6
    gW, gb = golden_grads(d, "pcd_grad_")
    for l in range(len(W)):
        assert rel_err(dW[l], gW[l]) < 2e-4, f"dW{l}"
        assert rel_err(db[l], gb[l]) < 2e-4, f"db{l}"


def test_grid_points_order():
    # executor/executor.py:351-356: flat index (ix*n + iy)*n + iz, x slowest
    n = 5
    pts = orc.grid_points(n)
    ax = torch.linspace(-1, 1, n)
    assert pts.shape == (n ** 3, 3)
    idx = (3 * n + 1) * n + 4
    assert torch.equal(pts[idx], torch.stack([ax[3], ax[1], ax[4]]))
    assert torch.equal(orc.grid_points(n, 2, 4), pts[2 * n * n: 4 * n * n])


def test_mc_cases_sphere():
    n = 12
    ax = np.linspace(-1, 1, n, dtype=np.float32)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    sdf = np.sqrt(X * X + Y * Y + Z * Z) - np.float32(0.6)
    cases, active = orc.mc_cases(sdf)
    assert cases.shape == (n - 1, n - 1, n - 1)
    # brute force
    for (i, j, k) in [(0, 0, 0), (5, 5, 5), (3, 7, 2), (8, 5, 5)]:
        m = 0
        for c, (dx, dy, dz) in enumerate(orc._MC_CORNERS):
            m |= int(sdf[i + dx, j + dy, k + dz] < 0) << c
        assert cases[i, j, k] == m
    flat = cases.reshape(-1)
    assert np.array_equal(active, np.nonzero((flat != 0) & (flat != 255))[0])
    assert 0 < active.size < flat.size


def test_mc_edge_vertices_sphere():
    """The interpolated edge points of a sphere's active cells lie on the sphere (to the linear-interpolation error)."""
    n = 24
    ax = np.linspace(-1, 1, n, dtype=np.float32)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    sdf = (np.sqrt(X * X + Y * Y + Z * Z) - np.float32(0.6)).astype(np.float32)
    _, active = orc.mc_cases(sdf)
    h = 2.0 / (n - 1)
    verts, mask = orc.mc_edge_vertices(sdf, active, (h, h, h))
    assert verts.shape == (active.size, 12, 3) and mask.dtype == np.uint16
    assert (mask != 0).all()
    on = ((mask[:, None] >> np.arange(12)[None, :]) & 1).astype(bool)
    p = verts[on] - 1.0  # index space * spacing has its origin at the (-1,-1,-1) corner
    r = np.sqrt((p * p).sum(-1))
    assert np.abs(r - 0.6).max() < 0.5 * h * h / 0.6 + 1e-5
    assert (verts[~on] == 0).all()
    # one cell by hand
    c = int(active[len(active) // 2])
    k, j, i = c % (n - 1), (c // (n - 1)) % (n - 1), c // ((n - 1) * (n - 1))
    for e, (a, b) in enumerate(orc._MC_EDGES):
        ia = np.array([i, j, k]) + np.array(orc._MC_CORNERS[a])
        ib = np.array([i, j, k]) + np.array(orc._MC_CORNERS[b])
        fa, fb = sdf[tuple(ia)], sdf[tuple(ib)]
        if (fa < 0) != (fb < 0):
            t = np.float32(fa / np.float32(fa - fb))
            want = (ia.astype(np.float32) + t * (ib - ia).astype(np.float32)) * np.float32(h)
            assert np.allclose(verts[len(active) // 2, e], want, rtol=0, atol=1e-6)


def test_octree_keys_and_refinement():
    lv = orc.octree_build(lambda c: np.sqrt((c * c).sum(-1)) - np.float32(0.5), 5, 2, np.sqrt(3.0) / 2)
    assert [len(l["keys"]) for l in lv[:3]] == [1, 8, 64]
    # centres of level 1: +-0.5 in x-major bit order
    c1 = orc.octree_centers(lv[1]["keys"], 1)
    assert np.array_equal(c1[0], np.float32([-0.5, -0.5, -0.5])) and np.array_equal(c1[4], np.float32([0.5, -0.5, -0.5]))
    assert np.array_equal(c1[1], np.float32([-0.5, -0.5, 0.5]))
    # children are inside their parent
    p = orc.octree_centers(lv[3]["keys"], 3)
    kids = orc.octree_children(lv[3]["keys"], np.ones(len(lv[3]["keys"]), np.uint8))
    ck = orc.octree_centers(kids, 4).reshape(-1, 8, 3)
    assert np.abs(ck - p[:, None, :]).max() == np.float32(2.0 ** -4)
    # a 1-Lipschitz band never drops a cell that contains the surface
    for l in lv[2:]:
        h = 2.0 ** (1 - l["level"])
        assert np.all((np.abs(l["sdf"]) <= np.float32(np.float32(np.sqrt(3.0) / 2) * np.float32(h))) == (l["flags"] != 0)) or l["level"] < 2


def test_adam_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(100, dtype=torch.float64)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn(100, dtype=torch.float64)
        ref.grad = g.clone()
        opt.step()
        p, m, v = orc.adam_step(p, g, m, v, 1e-3, step)
    assert torch.allclose(p, ref.detach(), atol=1e-12)
