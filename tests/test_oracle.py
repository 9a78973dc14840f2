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


# ------------------------------------------------------------------------------------------------ round 2 vectors
def test_igr_mask_branch_matches_reference():
    """regularizer_threshold (0.05) below delta (0.1): the where(|t| < thr, ., 1e-8) branch of model/losses.py:127-134."""
    d = load_golden("round2_128x4.npz")
    assert 0.05 < float(d["mask_frac"]) < 0.95  # both branches are exercised
    W, b = golden_params(d)
    x = torch.from_numpy(d["x"]).double()
    y = torch.from_numpy(d["y"]).double()
    loss, _, dW, db = orc.igr_loss_and_grads(W, b, x, y, skip_of(d), float(d["beta"]), 0.1, 1.0, 0.1, 0.05, "pairwise")
    assert abs(float(loss) - float(d["mask_loss"])) < 2e-6 * max(1.0, abs(float(d["mask_loss"])))
    gW, gb = golden_grads(d, "mask_grad_")
    for l in range(len(W)):
        assert rel_err(dW[l], gW[l]) < 2e-4, f"dW{l}"
        assert rel_err(db[l], gb[l]) < 2e-4, f"db{l}"


@pytest.mark.parametrize("tag,thr", [("pw", 1.0), ("pwmask", 0.05)])
def test_igr_pointwise_matches_reference_arithmetic(tag, thr):
    """data_term="pointwise": mean((p_i - t_i)^2) assembled from the reference's own forward / compute_normal
    (oracle/make_golden.py ref_igr_pointwise)."""
    d = load_golden("round2_128x4.npz")
    W, b = golden_params(d)
    x = torch.from_numpy(d["x"]).double()
    y = torch.from_numpy(d["y"]).double()
    loss, _, dW, db = orc.igr_loss_and_grads(W, b, x, y, skip_of(d), float(d["beta"]), 0.1, 1.0, 0.1, thr, "pointwise")
    assert abs(float(loss) - float(d[tag + "_loss"])) < 2e-6 * max(1.0, abs(float(d[tag + "_loss"])))
    gW, gb = golden_grads(d, tag + "_grad_")
    for l in range(len(W)):
        assert rel_err(dW[l], gW[l]) < 2e-4, f"dW{l}"
        assert rel_err(db[l], gb[l]) < 2e-4, f"db{l}"


def test_loss_curve_matches_reference_loop():
    """60 steps of executor/executor.py:167-178 (IGRLOSS, Adam lr 4e-4) from identical weights and batches: the oracle's
    sweeps + written-out Adam track the unmodified reference's loss curve and end at the same weights."""
    d = load_golden("round2_128x4.npz")
    L = int(d["num_hidden"])
    W = [torch.from_numpy(d[f"curve_init_lin{i}_weight"]).double() for i in range(L + 1)]
    b = [torch.from_numpy(d[f"curve_init_lin{i}_bias"]).double() for i in range(L + 1)]
    px, py = torch.from_numpy(d["curve_x"]).double(), torch.from_numpy(d["curve_y"]).double()
    mW = [torch.zeros_like(w) for w in W]; vW = [torch.zeros_like(w) for w in W]
    mb = [torch.zeros_like(w) for w in b]; vb = [torch.zeros_like(w) for w in b]
    curve = []
    for i in range(60):
        lo = 1024 * (i % 4)
        loss, _, dW, db = orc.igr_loss_and_grads(W, b, px[lo:lo + 1024], py[lo:lo + 1024], skip_of(d), float(d["beta"]))
        curve.append(float(loss))
        for l in range(L + 1):
            W[l], mW[l], vW[l] = orc.adam_step(W[l], dW[l], mW[l], vW[l], 4e-4, i + 1)
            b[l], mb[l], vb[l] = orc.adam_step(b[l], db[l], mb[l], vb[l], 4e-4, i + 1)
    ref = d["curve_loss"]
    rel = np.abs(np.array(curve) / ref - 1).max()
    assert ref[-1] < 0.7 * ref[0]  # the curve actually descends
    # fp32 reference vs fp64 oracle through 60 Adam steps (Adam's g / sqrt(v) amplifies rounding of tiny gradients)
    assert rel < 2e-3, rel
    for l in range(L + 1):
        assert rel_err(W[l], torch.from_numpy(d[f"curve_final_lin{l}_weight"])) < 2e-3


def test_full_network_per_element_gradients_at_8192():
    """BASELINE cfg 2 network at the reference's native batch scale (8192 points, [N,N] data term): every sampled
    gradient element of all 18 tensors against the unmodified reference's autograd."""
    d = load_golden("round2_full_8x512_n8192.npz")
    shapes = orc.layer_shapes(3, [512] * 8, (4,))
    import hashlib

    torch.manual_seed(0)
    W, b = [], []
    for l, (o, i) in enumerate(shapes):  # replays model/networks.py:62-74 (nn.Linear default init, then geometric init)
        lin = torch.nn.Linear(i, o)
        if l == len(shapes) - 1:
            torch.nn.init.normal_(lin.weight, mean=np.sqrt(np.pi) / np.sqrt(i), std=0.00001)
            torch.nn.init.constant_(lin.bias, -1)
        else:
            torch.nn.init.constant_(lin.bias, 0.0)
            torch.nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(o))
        W.append(lin.weight.detach())
        b.append(lin.bias.detach())
    h = hashlib.sha256()
    for w_, b_ in zip(W, b):
        h.update(w_.numpy().tobytes())
        h.update(b_.numpy().tobytes())
    if h.hexdigest() != str(d["weights_sha256"]):
        pytest.skip("torch build generates different seeded weights than the container that made the golden file")
    x = torch.from_numpy(d["x"])
    y = torch.from_numpy(d["y"])
    loss, _, dW, db = orc.igr_loss_and_grads(W, b, x, y, (4,), 100.0)
    assert abs(float(loss) - float(d["igr_loss"])) < 5e-6 * max(1.0, abs(float(d["igr_loss"])))
    st = int(d["stride"])
    for l in range(len(W)):
        gw = dW[l].reshape(-1)[::st].numpy()
        ew = np.abs(gw - d[f"igr_grad_lin{l}_weight"]).max() / float(d[f"igr_gmax_lin{l}_weight"])
        eb = np.abs(db[l].numpy() - d[f"igr_grad_lin{l}_bias"]).max() / max(float(d[f"igr_gmax_lin{l}_bias"]), 1e-30)
        assert ew < 1e-3 and eb < 1e-3, (l, ew, eb)  # fp32 vs fp32, different summation orders over 8192 points
