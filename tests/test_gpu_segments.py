"""GPU parity of the line-segment query objects: tgq_closest_segment through the C ABI against the CPU oracle
(pinned to the reference's own BvhN with LineSegment3 queries by tests/test_ref_pin.py) -- bit equality of indices, squared distances
and both closest points -- on triangle and point structures, host and device buffers, with distance bounds and the per-query
exclusions of MeshFix's VertexFilter (Tools/MeshFix/MeshFix.cpp:678-697,756-763)."""
import os

import numpy as np
import pytest

from oracle import orc
from thea_b200 import gen

pytestmark = pytest.mark.gpu


def _segments(rng, Q, lo, hi, lens):
    a = rng.uniform(lo, hi, (Q, 3)).astype(np.float32)
    d = rng.standard_normal((Q, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    b = (a + d * rng.choice(lens, (Q, 1))).astype(np.float32)
    return a, b


def _check(got, want, tag):
    gi, gd, ge, gs = got
    wi, wd, we, ws = want
    bad = np.flatnonzero(gi != wi.astype(np.int32))
    assert bad.size == 0, f"{tag}: {bad.size} indices differ, first query {bad[:5]}: {gi[bad[:5]]} vs {wi[bad[:5]]}, d2 {gd[bad[:5]]} vs {wd[bad[:5]]}"
    assert np.array_equal(gd.view(np.uint32), wd.view(np.uint32)), f"{tag}: squared distances differ"
    f = wi >= 0
    assert np.array_equal(ge[f], we[f]) and np.array_equal(gs[f], ws[f]), f"{tag}: closest points differ"


@pytest.mark.parametrize("kind", ["triangles", "points"])
def test_segment_queries_match_oracle(built, kind):
    from thea_b200 import Engine
    rng = np.random.default_rng(41)
    Q = 60000
    e = Engine(0)
    if kind == "triangles":
        elems = gen.torus_mesh(96)                      # 18,432 triangles
        lo, hi = gen.mesh_bounds(elems)
        lo, hi = float(np.min(lo)) * 1.3, float(np.max(hi)) * 1.3
        e.build_triangles(elems)
        ora, n, bound = orc.Oracle(tris9=elems), len(elems), 0.08
        lens = [0.0, 1e-3, 0.02, 0.1, 0.1, 0.6]
    else:
        elems = rng.uniform(-1, 1, (200000, 3)).astype(np.float32)
        elems[70000:70500] = elems[100:600]             # duplicates: exact ties
        lo, hi = -1.1, 1.1
        e.build_points(elems)
        ora, n, bound = orc.Oracle(points=elems), len(elems), 0.02
        lens = [0.0, 1e-3, 0.02, 0.1, 0.1, 1.0]
    a, b = _segments(rng, Q, lo, hi, lens)
    free = ora.closest_segment(a, b, mode=orc.MODE_CANON)
    excl = rng.integers(0, n, (Q, 4))
    excl[::3, 2:] = -1
    excl[::2, 0] = free[0][::2]                         # half of the queries exclude what they would have found
    _check(e.closest_segment(a, b), free, f"{kind} unbounded")
    _check(e.closest_segment(a, b, exclude=excl), ora.closest_segment(a, b, exclude=excl, mode=orc.MODE_CANON), f"{kind} excluded")
    want = ora.closest_segment(a, b, bound, excl[:, :2], mode=orc.MODE_CANON)
    assert 0 < (want[0] >= 0).sum() < Q
    _check(e.closest_segment(a, b, bound, excl[:, :2]), want, f"{kind} bounded + excluded")
    # the stored-direction form: b3 = direction as LineSegment3(a, b) computes it
    _check(e.closest_segment(a, (b - a).astype(np.float32), b_is_direction=True), free, f"{kind} direction form")
    # a small unsorted batch and device-resident inputs
    _check(e.closest_segment(a[:100], b[:100]), tuple(x[:100] for x in free), f"{kind} small batch")
    import torch
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    _check(e.closest_segment(ta, tb), free, f"{kind} device inputs")
    # the error budget: a build with every absolute margin x4 returns the same bits
    wide = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "thea_b200", "libtgq_wide.so")
    if os.path.exists(wide):
        w = Engine(0, lib=wide)
        if kind == "triangles":
            w.build_triangles(elems)
        else:
            w.build_points(elems)
        _check(w.closest_segment(a, b, exclude=excl), ora.closest_segment(a, b, exclude=excl, mode=orc.MODE_CANON), f"{kind} wide margins")
    with pytest.raises(Exception):
        e.closest_segment(a[:4], b[:4], exclude=np.zeros((4, 5), np.int32))       # more than 4 exclusions per query


def test_meshfix_t_junction_queries(built):
    """The query loop of MeshFix's T-junction repair (Tools/MeshFix/MeshFix.cpp:716-763) as one batch: a structure over the boundary
    vertices of an open mesh, one shrunk segment per boundary edge (segFromEdge, :700-713), each query ignoring the vertices of the edge's
    own face (VertexFilter, :678-697), dist_bound = 2 tol.  Every edge that has a foreign vertex on it must find that vertex."""
    from thea_b200 import Engine
    n = 120
    g = np.linspace(0, 1, n + 1, dtype=np.float32)
    X, Y = np.meshgrid(g, g, indexing="ij")
    V = np.stack([X, Y, (0.1 * np.sin(5 * X) * np.cos(3 * Y)).astype(np.float32)], -1).reshape(-1, 3).astype(np.float32)
    vid = lambda i, j: i * (n + 1) + j
    # boundary edges of the grid (with the third vertex of their only face) along the j = 0 side and the i = 0 side
    edges = [(vid(i, 0), vid(i + 1, 0), vid(i + 1, 1)) for i in range(n)] + [(vid(0, j), vid(0, j + 1), vid(1, j + 1)) for j in range(n)]
    edges = np.array(edges, np.int64)
    bverts = np.unique(np.concatenate([edges[:, 0], edges[:, 1]]))
    # foreign boundary vertices (a finer neighbouring patch): one on the middle of every third boundary edge, slightly off it
    rng = np.random.default_rng(43)
    tj = np.arange(0, len(edges), 3)
    mid = (0.5 * (V[edges[tj, 0]] + V[edges[tj, 1]]) + rng.normal(0, 2e-5, (len(tj), 3))).astype(np.float32)
    pts = np.concatenate([V[bverts], mid]).astype(np.float32)
    pos = {int(v): k for k, v in enumerate(bverts)}
    tol = 0.001 * float(np.linalg.norm(V.max(0) - V.min(0)))
    p0, p1 = V[edges[:, 0]], V[edges[:, 1]]
    ln = np.linalg.norm(p1 - p0, axis=1)
    assert (ln >= 2 * tol).all()
    s = (tol / ln).astype(np.float64)[:, None]
    a = ((1 - s.astype(np.float32)) * p0 + s.astype(np.float32) * p1).astype(np.float32)
    b = ((1 - (1 - s).astype(np.float32)) * p0 + (1 - s).astype(np.float32) * p1).astype(np.float32)
    excl = np.array([[pos.get(int(e[0]), -1), pos.get(int(e[1]), -1), pos.get(int(e[2]), -1)] for e in edges], np.int64)
    e = Engine(0)
    e.build_points(pts)
    got = e.closest_segment(a, b, 2 * tol, excl)
    want = orc.Oracle(points=pts).closest_segment(a, b, 2 * tol, excl, mode=orc.MODE_CANON)
    _check(got, want, "t-junctions")
    found = got[0] >= 0
    assert np.array_equal(np.flatnonzero(found), tj)
    assert np.array_equal(got[0][tj], len(bverts) + np.arange(len(tj)))
    from oracle import ref
    if ref.available():
        r = ref.Reference(points=pts).closest_segments(a, b, 2 * tol, excl)
        assert np.array_equal(r[0], want[0]) and np.array_equal(r[1].view(np.uint32), want[1].view(np.uint32))
