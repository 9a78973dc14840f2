#!/usr/bin/env python
"""Generates the marching-cubes triangle table used by the GPU kernel (csrc/mc_tables.h) and by the oracle
(oracle/mc_tri_table.npz) -- the step after classification / vertex interpolation in
``skimage.measure.marching_cubes(sdf, level=0.0)`` (reference executor/executor.py:388).

skimage 0.22 uses Lewiner's 33-case tables with interior ambiguity tests; that code is third-party and absent from
/root/reference (parity unpinned).  This table is therefore DERIVED, not transcribed, by a rule
that can be checked by hand and that makes the surface watertight across cells by construction:

* corner / edge numbering: Lorensen-Cline (the numbering of sdfnet_mc_classify / sdfnet_mc_edge_vertices);
* case index bit c is set iff corner c is negative (inside);
* on every cube face, walking its 4 corners counter-clockwise as seen from OUTSIDE the cube, each maximal run of
  negative corners is cut off by ONE segment from the edge where the run starts (+ -> -) to the edge where it ends
  (- -> +).  An ambiguous face (diagonal corners negative) thus always separates its two negative corners -- the two
  cells sharing the face see the same corner signs and draw the same two segments, so no holes appear;
* the directed segments of the 6 faces chain into closed loops (every crossed edge lies on exactly two faces, once as a
  segment start and once as a segment end); each loop is a polygon of 3..7 crossed edges, fan-triangulated from its
  smallest edge id.  With the segment direction above the triangles' normals point from the negative to the positive
  side (outward for an SDF), checked at the bottom for every case.

    python tools/make_mc_table.py        # rewrites csrc/mc_tables.h and oracle/mc_tri_table.npz (deterministic)
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CORNER = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])
EDGE_A = [0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3]
EDGE_B = [1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7]
EDGE_OF = {}
for e, (a, b) in enumerate(zip(EDGE_A, EDGE_B)):
    EDGE_OF[(a, b)] = e
    EDGE_OF[(b, a)] = e


def faces_ccw_from_outside():
    """The 6 faces as corner cycles, counter-clockwise when looked at from outside the cube."""
    faces = []
    for axis in range(3):
        for side in (0, 1):
            cs = [c for c in range(8) if CORNER[c][axis] == side]
            ctr = CORNER[cs].mean(0)
            normal = np.zeros(3)
            normal[axis] = 1.0 if side else -1.0
            # sort corners by angle around the outward normal
            u = np.roll(np.eye(3), 1, axis=0)[axis]  # any in-plane axis
            v = np.cross(normal, u)
            ang = [np.arctan2(np.dot(CORNER[c] - ctr, v), np.dot(CORNER[c] - ctr, u)) for c in cs]
            order = [c for _, c in sorted(zip(ang, cs))]
            faces.append(order)
    return faces


FACES = faces_ccw_from_outside()


def case_loops(case):
    neg = [(case >> c) & 1 for c in range(8)]
    nxt = {}  # directed segments: start edge -> end edge
    for cyc in FACES:
        s = [neg[c] for c in cyc]
        if sum(s) in (0, 4):
            continue
        for i in range(4):
            # a negative run starts at corner i if corner i is negative and its predecessor is positive
            if s[i] and not s[(i - 1) % 4]:
                j = i
                while s[(j + 1) % 4]:
                    j = (j + 1) % 4
                e_start = EDGE_OF[(cyc[(i - 1) % 4], cyc[i])]   # + -> - : the run starts behind this edge
                e_end = EDGE_OF[(cyc[j], cyc[(j + 1) % 4])]     # - -> + : the run ends at this edge
                assert e_start not in nxt
                nxt[e_start] = e_end
    loops, seen = [], set()
    for e0 in sorted(nxt):
        if e0 in seen:
            continue
        loop, e = [], e0
        while e not in seen:
            seen.add(e)
            loop.append(e)
            e = nxt[e]
        assert e == e0, "segments do not close"
        loops.append(loop)
    return loops


def build():
    tris = []
    for case in range(256):
        t = []
        for loop in case_loops(case):
            k = loop.index(min(loop))
            loop = loop[k:] + loop[:k]
            for i in range(1, len(loop) - 1):
                t.append((loop[0], loop[i], loop[i + 1]))
        tris.append(t)
    return tris


def check(tris):
    mid = (CORNER[EDGE_A] + CORNER[EDGE_B]) / 2.0
    for case, t in enumerate(tris):
        neg = [(case >> c) & 1 for c in range(8)]
        crossed = {e for e in range(12) if neg[EDGE_A[e]] != neg[EDGE_B[e]]}
        used = {e for tri in t for e in tri}
        assert used == crossed, case
        for tri in t:
            assert len(set(tri)) == 3
        # every polygon edge that lies inside the cube volume is shared by two triangles of the cell in opposite
        # directions; polygon boundary segments lie on faces and are matched by the neighbour cell (same rule)
        directed = {}
        for a, b, c in t:
            for u, v in ((a, b), (b, c), (c, a)):
                directed[(u, v)] = directed.get((u, v), 0) + 1
        assert all(v == 1 for v in directed.values()), case
        # orientation: the normal of every triangle points away from the negative corners it cuts off.  Test with the
        # trilinear-free proxy: signed distance of the negative corners' centroid to the triangle plane is negative for
        # single-loop cases (general cases are covered by the mesh-level volume test in tests/test_mc_triangles.py)
        if len(case_loops(case)) == 1 and 0 < sum(neg) <= 3 and t:
            cneg = CORNER[[c for c in range(8) if neg[c]]].mean(0)
            n_sum = 0.0
            for a, b, c in t:
                nrm = np.cross(mid[b] - mid[a], mid[c] - mid[a])
                n_sum += np.dot(nrm, cneg - mid[a])
            assert n_sum < 0, (case, n_sum)
    return max(len(t) for t in tris)


def main():
    tris = build()
    max_t = check(tris)
    table = np.full((256, max_t * 3), 255, dtype=np.uint8)
    count = np.zeros(256, dtype=np.uint8)
    for case, t in enumerate(tris):
        count[case] = len(t)
        for i, tri in enumerate(t):
            table[case, 3 * i:3 * i + 3] = tri
    np.savez_compressed(os.path.join(ROOT, "oracle", "mc_tri_table.npz"), table=table, count=count)
    with open(os.path.join(ROOT, "sdf-representation_b200", "csrc", "mc_tables.h"), "w") as f:
        f.write("// GENERATED by tools/make_mc_table.py -- do not edit.  Marching-cubes triangle table: for case index c (bit k set iff\n"
                "// corner k is negative, Lorensen-Cline numbering) kMcTriCount[c] triangles, each three edge ids.  Derivation and\n"
                "// the hand-checkable rule: tools/make_mc_table.py; invariants: tests/test_mc_triangles.py.\n#pragma once\n#include <cstdint>\n\n")
        f.write(f"constexpr int kMcMaxTris = {max_t};\n\n")
        f.write("__constant__ uint8_t kMcTriCount[256] = {\n")
        for r in range(0, 256, 32):
            f.write("    " + ", ".join(str(int(v)) for v in count[r:r + 32]) + ",\n")
        f.write("};\n\n")
        f.write(f"__constant__ uint8_t kMcTriTable[256][{max_t * 3}] = {{\n")
        for case in range(256):
            f.write("    {" + ", ".join(f"{int(v):3d}" for v in table[case]) + "},\n")
        f.write("};\n")
    print(f"256 cases, max {max_t} triangles per cell, {int(count.sum())} triangles in the table")


if __name__ == "__main__":
    main()
