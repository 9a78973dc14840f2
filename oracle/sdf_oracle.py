"""ORACLE -- test infrastructure, not product code.

CPU restatement of the reference's neural-SDF hot path (newton-raphson/sdf-representation).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module; the
product package (``sdf-representation_b200/``) never does.

Everything is written out by hand (explicit layer loops, explicit reverse / tangent / adjoint sweeps, closed-form
loss moments) with torch CPU tensors as the array type, so it restates the reference's *algorithm* rather than calling
its autograd.  Each function cites the reference lines it follows.  Parity pinning: ``oracle/make_golden.py`` runs the
UNMODIFIED reference modules (``/root/reference/model/networks.py``, ``model/losses.py``) in the build container and
stores their inputs/outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks this file against those vectors.
The reference's own tests hold no golden vector for this path -- the octree refinement rule and the
marching-cubes cell classification have no reference counterpart at all and are "parity unpinned" (build-defined).
"""
import math
import os

import numpy as np
import torch

C_SKIP = 1.0 / math.sqrt(2.0)


# ------------------------------------------------------------------------------------------------ activation
def act(z, beta):
    """torch.nn.Softplus(beta, threshold=20) / ReLU -- reference model/networks.py:79-83."""
    if beta > 0:
        bz = beta * z
        return torch.where(bz > 20, z, torch.log1p(torch.exp(torch.clamp(bz, max=20.0))) / beta)
    return torch.clamp(z, min=0)


def act_d1(z, beta):
    if beta > 0:
        bz = beta * z
        return torch.where(bz > 20, torch.ones_like(z), torch.sigmoid(bz))
    return (z > 0).to(z.dtype)


def act_d2(z, beta):
    if beta > 0:
        bz = beta * z
        s = torch.sigmoid(bz)
        return torch.where(bz > 20, torch.zeros_like(z), beta * s * (1 - s))
    return torch.zeros_like(z)


# ------------------------------------------------------------------------------------------------ network
def layer_shapes(d_in, hidden, skip_in):
    """(out, in) of lin0..linL -- reference model/networks.py:41,55-62."""
    dims = [d_in] + list(hidden) + [1]
    shapes = []
    for layer in range(len(dims) - 1):
        out_dim = dims[layer + 1] - d_in if (layer + 1) in skip_in else dims[layer + 1]
        shapes.append((out_dim, dims[layer]))
    return shapes


def forward_states(W, b, x, skip_in, beta):
    """Sweep 1 -- reference model/networks.py:87-106.  Returns (f [N], h list, z list)."""
    L = len(W) - 1
    hs, zs = [], []
    a = x
    for l in range(L + 1):
        h = torch.cat([a, x], -1) * C_SKIP if l in skip_in else a
        z = h @ W[l].T + b[l]
        hs.append(h)
        zs.append(z)
        if l < L:
            a = act(z, beta)
    f = zs[L][:, 0]
    if beta <= 0:
        f = torch.tanh(f)  # model/networks.py:104-105
    return f, hs, zs


def forward(W, b, x, skip_in, beta):
    return forward_states(W, b, x, skip_in, beta)[0]


def reverse_states(W, x, zs, skip_in, beta):
    """Sweep 2 -- the input gradient the reference obtains with autograd.grad (model/losses.py:283-296).
    Returns (G = d z_L / d x [N,3], q list with q_l = d z_L / d z_l, s list with s_l = d z_L / d a_l)."""
    L = len(W) - 1
    d_in = x.shape[1]
    N = x.shape[0]
    q = [None] * (L + 1)
    s = [None] * L
    q[L] = torch.ones(N, 1, dtype=x.dtype)
    G = torch.zeros(N, d_in, dtype=x.dtype)
    for l in range(L, 0, -1):
        r = q[l] @ W[l]
        if l in skip_in:
            G = G + C_SKIP * r[:, -d_in:]
            s[l - 1] = C_SKIP * r[:, :-d_in]
        else:
            s[l - 1] = r
        q[l - 1] = s[l - 1] * act_d1(zs[l - 1], beta)
    G = G + q[0] @ W[0]
    return G, q, s


def forward_and_gradient(W, b, x, skip_in, beta):
    """(f [N], grad_x f [N,3]) -- model/networks.py:87-106 + model/losses.py:283-296."""
    f, hs, zs = forward_states(W, b, x, skip_in, beta)
    G, _, _ = reverse_states(W, x, zs, skip_in, beta)
    if beta <= 0:
        G = G * (1 - f * f).unsqueeze(-1)
    return f, G


def param_grads(W, b, x, skip_in, beta, fbar, gbar):
    """Sweeps 3-4 + parameter gradients for seeds fbar = dL/df [N], gbar = dL/d(grad_x f) [N,3] (either may be None):
    what ``loss.backward()`` (executor/executor.py:176) accumulates through the graphs built by
    model/networks.py:87-106 and model/losses.py:283-296.  Returns (dW list, db list)."""
    L = len(W) - 1
    d_in = x.shape[1]
    N = x.shape[0]
    f, hs, zs = forward_states(W, b, x, skip_in, beta)
    G, q, s = reverse_states(W, x, zs, skip_in, beta)
    fbar = torch.zeros(N, dtype=x.dtype) if fbar is None else fbar
    have_g = gbar is not None
    gbar = torch.zeros(N, d_in, dtype=x.dtype) if gbar is None else gbar
    if beta <= 0:
        # f = tanh(z_L), g = (1 - f^2) G: move the seeds onto (z_L, G)
        d = 1 - f * f
        zbar = fbar * d - 2 * f * (gbar * (G * d.unsqueeze(-1))).sum(-1)
        fbar, gbar = zbar, gbar * d.unsqueeze(-1)
    # sweep 3: forward tangent seeded with gbar
    hd = [None] * (L + 1)
    zd = [None] * (L + 1)
    ad = gbar
    for l in range(L + 1):
        hd[l] = torch.cat([ad, gbar], -1) * C_SKIP if l in skip_in else ad
        zd[l] = hd[l] @ W[l].T
        if l < L:
            ad = act_d1(zs[l], beta) * zd[l]
    # sweep 4: reverse total adjoint
    Zb = [None] * (L + 1)
    Zb[L] = fbar.unsqueeze(-1)
    for l in range(L - 1, -1, -1):
        r = Zb[l + 1] @ W[l + 1]
        abar = C_SKIP * r[:, :-d_in] if (l + 1) in skip_in else r
        Zb[l] = act_d1(zs[l], beta) * abar
        if have_g:
            Zb[l] = Zb[l] + act_d2(zs[l], beta) * s[l] * zd[l]
    dW = [Zb[l].T @ hs[l] + q[l].T @ hd[l] for l in range(L + 1)]
    db = [Zb[l].sum(0) for l in range(L + 1)]
    return dW, db


# ------------------------------------------------------------------------------------------------ losses
def igr_terms(f, g, y, delta, tau, lambda_g, thr, data_term="pairwise"):
    """IGRLOSS value and seeds -- reference model/losses.py:109-137 (data term: SURVEY.md T1).
    Returns (loss, (data, reg, eik), fbar [N], gbar [N,3])."""
    N = f.shape[0]
    p = torch.clamp(f, -delta, delta)
    t = torch.clamp(y[:, 0], -delta, delta)
    n = y[:, 1:4]
    m = t.abs() < thr
    gn = g.norm(2, dim=-1)
    nn_ = n.norm(2, dim=-1)
    gnc = torch.clamp(gn, min=1e-8)
    nnc = torch.clamp(nn_, min=1e-8)
    cos = (n * g).sum(-1) / (nnc * gnc)
    eps = torch.full_like(f, 1e-8)
    reg = torch.where(m, (1 - cos) ** 2, eps)
    eik = torch.where(m, (gn - 1) ** 2, eps)
    if data_term == "pairwise":
        data = (p * p).mean() + (t * t).mean() - 2 * p.mean() * t.mean()
    else:
        data = ((p - t) ** 2).mean()
    loss = data + tau * reg.mean() + lambda_g * eik.mean()
    inside = ((f >= -delta) & (f <= delta)).to(f.dtype)
    if data_term == "pairwise":
        fbar = (2.0 / N) * (p - t.mean()) * inside
    else:
        fbar = (2.0 / N) * (p - t) * inside
    k1 = (-2.0 * tau / N) * (1 - cos)
    dcos = n / (nnc * gnc).unsqueeze(-1) - (cos / (gnc * gnc)).unsqueeze(-1) * g
    k3 = (2.0 * lambda_g / N) * (gn - 1) / gnc
    gbar = m.to(f.dtype).unsqueeze(-1) * (k1.unsqueeze(-1) * dcos + k3.unsqueeze(-1) * g)
    return loss, (data, reg.mean(), eik.mean()), fbar, gbar


def igr_loss_and_grads(W, b, x, y, skip_in, beta, delta=0.1, tau=1.0, lambda_g=0.1, thr=1.0, data_term="pairwise"):
    """One IGRLOSS training step's loss and parameter gradients (model/losses.py:109-137 + executor/executor.py:174-176)."""
    f, g = forward_and_gradient(W, b, x, skip_in, beta)
    loss, terms, fbar, gbar = igr_terms(f, g, y, delta, tau, lambda_g, thr, data_term)
    dW, db = param_grads(W, b, x, skip_in, beta, fbar, gbar)
    return loss, terms, dW, db


def wsl2_loss_and_grads(W, b, x, y, skip_in, beta, weight_factor=0.5, delta=0.1, data_term="pairwise"):
    """WeightedSmoothL2Loss -- reference model/losses.py:50-69: mean_{i,j} w_j (t_j - p_i)^2 in closed form."""
    N = x.shape[0]
    f = forward(W, b, x, skip_in, beta)
    t = torch.clamp(y[:, 0], -delta, delta)
    p = torch.clamp(f, -delta, delta)
    w = 1.0 + weight_factor * torch.exp(-t.abs())
    inside = ((f >= -delta) & (f <= delta)).to(f.dtype)
    if data_term == "pairwise":
        loss = (w * t * t).mean() - 2 * p.mean() * (w * t).mean() + (p * p).mean() * w.mean()
        fbar = (2.0 / N) * (p * w.mean() - (w * t).mean()) * inside
    else:
        loss = (w * (t - p) ** 2).mean()
        fbar = (2.0 / N) * w * (p - t) * inside
    dW, db = param_grads(W, b, x, skip_in, beta, fbar, None)
    return loss, dW, db


def pcd_loss_and_grads(W, b, x, xp, skip_in, beta, lambda_g):
    """Point-cloud step -- reference executor/executordistributed.py:63-79:
    mean |f(x)| + lambda_g mean((|grad f(xp)| - 1)^2)."""
    N, M = x.shape[0], xp.shape[0]
    f = forward(W, b, x, skip_in, beta)
    surf = f.abs().mean()
    dW1, db1 = param_grads(W, b, x, skip_in, beta, torch.sign(f) / N, None)
    _, g = forward_and_gradient(W, b, xp, skip_in, beta)
    gn = g.norm(2, dim=-1)
    eik = ((gn - 1) ** 2).mean()
    gbar = ((2.0 * lambda_g / M) * (gn - 1) / torch.clamp(gn, min=1e-30)).unsqueeze(-1) * g
    dW2, db2 = param_grads(W, b, xp, skip_in, beta, None, gbar)
    loss = surf + lambda_g * eik
    return loss, (surf, eik), [a + c for a, c in zip(dW1, dW2)], [a + c for a, c in zip(db1, db2)]


def adam_step(p, g, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (executor/executor.py:160,177), written out."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    p = p - (lr / bc1) * m / (v.sqrt() / math.sqrt(bc2) + eps)
    return p, m, v


# ------------------------------------------------------------------------------------------------ grid
def grid_points(n, ix0=0, ix1=None):
    """Coordinates of planes [ix0, ix1) of the reference's flattened grid -- executor/executor.py:351-356
    (torch.linspace itself is called so that coordinates are bit-identical)."""
    ix1 = n if ix1 is None else ix1
    ax = torch.linspace(-1, 1, n)
    xx, yy, zz = torch.meshgrid(ax[ix0:ix1], ax, ax, indexing="ij")
    return torch.stack((xx, yy, zz), dim=-1).reshape(-1, 3)


# ------------------------------------------------------------------------------------------------ marching cubes
_MC_CORNERS = ((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1))


def mc_cases(sdf):
    """Cell classification at level 0 (build-defined restatement of the first stage of
    skimage.measure.marching_cubes, called at executor/executor.py:388; scikit-image 0.22.0 is not installed and has
    no vector in the reference -- parity unpinned).  bit c of the case index is set iff corner c (Lorensen-Cline
    order) is < 0.  Returns (cases uint8 [(nx-1),(ny-1),(nz-1)], active flat indices int64)."""
    sdf = np.asarray(sdf, dtype=np.float32)
    nx, ny, nz = sdf.shape
    cases = np.zeros((nx - 1, ny - 1, nz - 1), dtype=np.uint8)
    for c, (dx, dy, dz) in enumerate(_MC_CORNERS):
        cases |= ((sdf[dx:nx - 1 + dx, dy:ny - 1 + dy, dz:nz - 1 + dz] < 0).astype(np.uint8) << c).astype(np.uint8)
    flat = cases.reshape(-1)
    active = np.nonzero((flat != 0) & (flat != 255))[0].astype(np.int64)
    return cases, active


_MC_EDGES = ((0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7))


def mc_edge_vertices(sdf, active, spacing=(1.0, 1.0, 1.0)):
    """Vertex interpolation of the active cells at level 0 (second stage of skimage.measure.marching_cubes as called at
    executor/executor.py:388 with spacing=...; build-defined restatement, parity unpinned like mc_cases): edge e joins
    corners _MC_EDGES[e]; a crossing edge gets t = f_a / (f_a - f_b) and (idx_a + t (idx_b - idx_a)) * spacing, every
    operation a single fp32 rounding.  Returns (verts float32 [n, 12, 3], mask uint16 [n])."""
    sdf = np.asarray(sdf, dtype=np.float32)
    nx, ny, nz = sdf.shape
    active = np.asarray(active, dtype=np.int64)
    cy, cz = ny - 1, nz - 1
    k = active % cz
    j = (active // cz) % cy
    i = active // (cz * cy)
    base = np.stack([i, j, k], 1)
    sp = np.asarray(spacing, dtype=np.float32)
    verts = np.zeros((active.shape[0], 12, 3), dtype=np.float32)
    mask = np.zeros(active.shape[0], dtype=np.uint16)
    corners = np.asarray(_MC_CORNERS, dtype=np.int64)
    for e, (a, b) in enumerate(_MC_EDGES):
        ia = base + corners[a]
        ib = base + corners[b]
        fa = sdf[ia[:, 0], ia[:, 1], ia[:, 2]]
        fb = sdf[ib[:, 0], ib[:, 1], ib[:, 2]]
        cross = (fa < 0) != (fb < 0)
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (fa / (fa - fb)).astype(np.float32)
        pa = ia.astype(np.float32)
        pb = ib.astype(np.float32)
        v = ((pa + (t[:, None] * (pb - pa)).astype(np.float32)).astype(np.float32) * sp[None, :]).astype(np.float32)
        verts[:, e, :] = np.where(cross[:, None], v, np.float32(0))
        mask |= (cross.astype(np.uint16) << np.uint16(e)).astype(np.uint16)
    return verts, mask


_MC_TABLE = None


def mc_tri_table():
    """(table [256, 15] uint8 edge ids padded with 255, count [256]) -- generated by tools/make_mc_table.py from a rule that
    is watertight by construction (third-party skimage tables are not in /root/reference: parity unpinned)."""
    global _MC_TABLE
    if _MC_TABLE is None:
        d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mc_tri_table.npz"))
        _MC_TABLE = (d["table"], d["count"])
    return _MC_TABLE


def mc_triangles(cases, active):
    """Faces of marching_cubes(sdf, 0.0) (executor/executor.py:388) for the active cells, in cell order then table order:
    int64 [n_tri, 3] indices r * 12 + edge into the [n_active, 12] edge-vertex array of ``mc_edge_vertices``."""
    table, count = mc_tri_table()
    cs = np.asarray(cases).reshape(-1)[np.asarray(active, dtype=np.int64)]
    out = []
    for r, c in enumerate(cs):
        for t in range(int(count[c])):
            out.append([r * 12 + int(table[c, 3 * t + k]) for k in range(3)])
    return np.asarray(out, dtype=np.int64).reshape(-1, 3)


# ------------------------------------------------------------------------------------------------ octree
def octree_centers(keys, level):
    """Cell centres of 3-bit-per-level Morton keys (x bit highest, finest level in the lowest bits) over [-1,1]^3:
    c = -1 + (2 i + 1) 2^-level, exact in fp32.  Build-defined (the reference's
    evaluations/compare_octree_dl.py:20-36 only loops over externally supplied nodes)."""
    keys = np.asarray(keys, dtype=np.uint64)
    ci = np.zeros((keys.shape[0], 3), dtype=np.uint64)
    for bit in range(level):
        for ax in range(3):
            ci[:, ax] |= ((keys >> np.uint64(3 * bit + 2 - ax)) & np.uint64(1)) << np.uint64(bit)
    hs = np.float32(2.0 ** -level)
    return (np.float32(-1.0) + (np.float32(2.0) * ci.astype(np.float32) + np.float32(1.0)) * hs).astype(np.float32)


def octree_flags(sdf, level, band):
    """refine iff |f| <= fp32(band) * fp32(2 / 2^level): one fp32 rounding of the product."""
    thresh = np.float32(np.float32(band) * np.float32(2.0 ** (1 - level)))
    return (np.abs(np.asarray(sdf, dtype=np.float32)) <= thresh).astype(np.uint8)


def octree_children(keys, flags):
    keys = np.asarray(keys, dtype=np.uint64)
    sel = keys[np.asarray(flags) != 0]
    return ((sel[:, None] << np.uint64(3)) | np.arange(8, dtype=np.uint64)[None, :]).reshape(-1)


def octree_build(sdf_fn, max_level, refine_lvl_uni, band):
    """Level loop; ``sdf_fn(centers [n,3] float32) -> sdf [n] float32``."""
    keys = np.zeros(1, dtype=np.uint64)
    out = []
    for level in range(max_level + 1):
        sdf = np.asarray(sdf_fn(octree_centers(keys, level)), dtype=np.float32)
        flags = octree_flags(sdf, level, band)
        if level < refine_lvl_uni:
            flags = np.ones_like(flags)
        out.append({"level": level, "keys": keys, "sdf": sdf, "flags": flags})
        if level == max_level:
            break
        keys = octree_children(keys, flags)
        if keys.size == 0:
            break
    return out


# ------------------------------------------------------------------------------------------------ synthetic labels
def torus_labels(x, R=0.5, r=0.2):
    """Analytic torus SDF and unit normal, label layout [S, nx, ny, nz] of dataloader/load_data.py:61-62
    (stands in for igl.signed_distance, datagenerator/data_generator.py:880-906, which is not installed)."""
    rho = torch.sqrt(x[:, 0] ** 2 + x[:, 1] ** 2)
    a = rho - R
    d = torch.sqrt(a * a + x[:, 2] ** 2)
    S = d - r
    rho_c = torch.clamp(rho, min=1e-12)
    d_c = torch.clamp(d, min=1e-12)
    n = torch.stack([a / d_c * x[:, 0] / rho_c, a / d_c * x[:, 1] / rho_c, x[:, 2] / d_c], -1)
    return torch.cat([S.unsqueeze(-1), n], -1)
