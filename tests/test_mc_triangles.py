"""Marching-cubes faces (the step after the path, executor/executor.py:388).  CPU: the committed triangle table is what
tools/make_mc_table.py derives, satisfies its invariants, and the oracle's meshes of a sphere and a torus are closed,
consistently oriented 2-manifolds of the right volume.  GPU: the kernel reproduces the oracle's faces bit for bit and
the slab-sharded reconstruction covers every cell exactly once."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import sdf_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _generator():
    spec = importlib.util.spec_from_file_location("make_mc_table", os.path.join(ROOT, "tools", "make_mc_table.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _sphere(n, r=0.6):
    ax = np.linspace(-1, 1, n, dtype=np.float32)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    return (np.sqrt(x * x + y * y + z * z) - np.float32(r)).astype(np.float32)


def _torus(n, R=0.5, r=0.23):
    ax = np.linspace(-1, 1, n, dtype=np.float32)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    return (np.sqrt((np.sqrt(x * x + y * y) - R) ** 2 + z * z) - np.float32(r)).astype(np.float32)


def _oracle_mesh(sdf, spacing):
    cases, active = orc.mc_cases(sdf)
    verts, _ = orc.mc_edge_vertices(sdf, active, spacing)
    faces = orc.mc_triangles(cases, active)
    return verts.reshape(-1, 3), faces, active


_CORNERS = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])
_EDGES = np.array([(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)])


def _weld(shape, active, faces):
    """Per-cell vertex index r * 12 + e -> id of the GRID edge it lies on (the up to 4 cells sharing a grid edge each carry
    their own copy of its vertex; Lorensen-Cline edges 2, 3, 6, 7 run against the axis, so the copies may differ in the
    last fp32 bit and cannot be matched by position)."""
    nx, ny, nz = shape
    cy, cz = ny - 1, nz - 1
    r, e = faces // 12, faces % 12
    cell = active[r]
    base = np.stack([cell // (cz * cy), (cell // cz) % cy, cell % cz], -1)
    ca = base + _CORNERS[_EDGES[e][..., 0]]
    cb = base + _CORNERS[_EDGES[e][..., 1]]
    fa = (ca[..., 0] * ny + ca[..., 1]) * nz + ca[..., 2]
    fb = (cb[..., 0] * ny + cb[..., 1]) * nz + cb[..., 2]
    return np.minimum(fa, fb) * (nx * ny * nz) + np.maximum(fa, fb)


def test_table_is_what_the_generator_derives_and_is_consistent():
    gen = _generator()
    tris = gen.build()
    assert gen.check(tris) == 5
    table, count = orc.mc_tri_table()
    assert table.shape == (256, 15) and int(count.sum()) == 820
    for case, t in enumerate(tris):
        assert int(count[case]) == len(t)
        assert [tuple(int(v) for v in table[case, 3 * i:3 * i + 3]) for i in range(len(t))] == t
    # hand-checkable rows: nothing to draw, one corner cut off (corner 0 touches edges 0, 3, 8), the complement
    assert count[0] == 0 and count[255] == 0
    assert set(int(v) for v in table[1, :3]) == {0, 3, 8} and count[1] == 1
    assert set(int(v) for v in table[254, :3]) == {0, 3, 8} and count[254] == 1
    assert tuple(table[1, :3]) != tuple(table[254, :3])  # opposite orientation
    # the CUDA header holds the same numbers
    hdr = open(os.path.join(ROOT, "sdf-representation_b200", "csrc", "mc_tables.h")).read()
    row = "{" + ", ".join(f"{int(v):3d}" for v in table[105]) + "},"
    assert row in hdr


@pytest.mark.parametrize("shape", ["sphere", "torus"])
def test_oracle_mesh_is_closed_oriented_and_has_the_right_volume(shape):
    n = 24
    sdf = _sphere(n) if shape == "sphere" else _torus(n)
    sp = (2.0 / (n - 1),) * 3
    verts, faces, active = _oracle_mesh(sdf, sp)
    assert faces.shape[0] > 500
    w = _weld(sdf.shape, active, faces)
    # every directed edge appears once and its reverse once: closed, consistently oriented 2-manifold (no holes at the
    # ambiguous faces, which is what the face rule of the table guarantees)
    e = np.concatenate([w[:, [0, 1]], w[:, [1, 2]], w[:, [2, 0]]])
    assert np.all(e[:, 0] != e[:, 1])
    fwd = set(map(tuple, e.tolist()))
    assert len(fwd) == e.shape[0]
    assert all((b, a) in fwd for a, b in fwd)
    # outward normals: signed volume is positive and close to the analytic one
    p = verts[faces].astype(np.float64) - 1.0  # index space * spacing -> coordinates in [-1, 1]
    vol = np.einsum("ij,ij->i", p[:, 0], np.cross(p[:, 1], p[:, 2])).sum() / 6.0
    exact = 4.0 / 3.0 * np.pi * 0.6 ** 3 if shape == "sphere" else 2 * np.pi ** 2 * 0.5 * 0.23 ** 2
    assert vol > 0 and abs(vol / exact - 1) < 0.05, (vol, exact)
    # Euler characteristic: sphere 2, torus 0
    V, E, F = len(np.unique(w)), len(fwd) // 2, faces.shape[0]
    assert V - E + F == (2 if shape == "sphere" else 0)


@pytest.mark.gpu
def test_gpu_faces_match_oracle_bit_exact():
    from sdf_representation_b200 import inference

    dev = torch.device("cuda:0")
    for sdf in (_sphere(33), _torus(40), np.random.default_rng(0).standard_normal((9, 17, 13)).astype(np.float32)):
        sp = (0.5, 0.25, 2.0)
        verts_o, faces_o, _ = _oracle_mesh(sdf, sp)
        block = torch.from_numpy(sdf).to(dev)
        verts, faces = inference.marching_cubes(block, spacing=sp)
        assert np.array_equal(faces.cpu().numpy(), faces_o)
        assert np.array_equal(verts.cpu().numpy(), verts_o)
    # empty volume
    v, f = inference.marching_cubes(torch.ones(5, 5, 5, device=dev))
    assert f.shape == (0, 3)


@pytest.mark.gpu
def test_gpu_slab_meshes_cover_the_volume_once():
    """BASELINE configs[3] composition: slab of the slowest axis + 1 halo plane per rank -> classification, vertices and
    faces on the slab.  The union of the slab meshes has the faces of the whole-volume mesh (each cell on one rank)."""
    from sdf_representation_b200 import inference

    dev = torch.device("cuda:0")
    n = 40
    sdf = torch.from_numpy(_torus(n)).to(dev)
    sp = (2.0 / n,) * 3
    v_all, f_all = inference.marching_cubes(sdf, spacing=sp)
    tri_all = v_all[f_all].cpu().numpy()
    parts = []
    world = 3
    for r in range(world):
        base, rem = divmod(n, world)
        ix0 = r * base + min(r, rem)
        ix1 = ix0 + base + (1 if r < rem else 0)
        hi = min(n, ix1 + 1)
        v, f = inference.marching_cubes(sdf[ix0:hi].contiguous(), spacing=sp, index_offset=(ix0, 0, 0))
        parts.append(v[f].cpu().numpy())
    tri = np.concatenate(parts)
    assert tri.shape == tri_all.shape
    assert np.allclose(tri, tri_all, atol=1e-6)  # same cells in the same order; offsets add one fp32 rounding
