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
    # the same loop through the caller (thea_b200/callers.py: segFromEdge + the batch + the acceptance test of MeshFix.cpp:766-771)
    from thea_b200 import callers
    split, ca, cb = callers.t_junction_candidates(e, p0, p1, tol, excl)
    assert np.allclose(ca, a, rtol=0, atol=1e-6) and np.allclose(cb, b, rtol=0, atol=1e-6)      # s = tol / len in double there, in float above
    assert np.array_equal(np.flatnonzero(split >= 0), tj) and np.array_equal(split[tj], len(bverts) + np.arange(len(tj)))
    cw = orc.Oracle(points=pts).closest_segment(ca, cb, 2 * tol, excl, mode=orc.MODE_CANON)
    assert np.array_equal(split[tj], cw[0][tj].astype(np.int32))


@pytest.mark.parametrize("kind", ["triangles", "points"])
def test_triangle_queries_match_oracle(built, kind):
    """tgq_closest_triangle (BvhN::closestPair<MetricL2>(Triangle3, ...)) against the oracle's exhaustive minimum of the reference's
    per-element value -- element.squaredDistance(query) with Moeller's intersection test on triangle structures (Triangle3.hpp:314-392),
    query.closestPoint(element) on point structures -- bit for bit, with and without a distance bound, host and device buffers."""
    from thea_b200 import Engine
    rng = np.random.default_rng(42)
    e = Engine(0)
    if kind == "triangles":
        elems = gen.torus_mesh(96)
        e.build_triangles(elems)
        ora, Q, lo, hi, sizes, bound = orc.Oracle(tris9=elems), 2500, -1.6, 1.6, [0.005, 0.05, 0.3], 0.1
    else:
        elems = rng.uniform(-1, 1, (50000, 3)).astype(np.float32)
        e.build_points(elems)
        ora, Q, lo, hi, sizes, bound = orc.Oracle(points=elems), 3000, -1.1, 1.1, [0.01, 0.1, 0.5], 0.02
    c = rng.uniform(lo, hi, (Q, 1, 3))
    q = (c + rng.normal(0, 1, (Q, 3, 3)) * rng.choice(sizes, (Q, 1, 1))).reshape(Q, 9).astype(np.float32)
    if kind == "triangles":
        q[:300] = (elems[:300] + np.float32(1e-3)).astype(np.float32)        # near copies of elements: intersections and near misses
    free = ora.closest_triangle(q, mode=orc.MODE_BRUTE)
    _check(e.closest_triangle(q), free, f"{kind} triangle queries")
    want = ora.closest_triangle(q, bound, mode=orc.MODE_BRUTE)
    assert 0 < (want[0] >= 0).sum() < Q
    _check(e.closest_triangle(q, bound), want, f"{kind} bounded triangle queries")
    _check(e.closest_triangle(q[:50]), tuple(x[:50] for x in free), f"{kind} small batch")
    import torch
    _check(e.closest_triangle(torch.from_numpy(q).cuda()), free, f"{kind} device inputs")
    wide = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "thea_b200", "libtgq_wide.so")
    if os.path.exists(wide):
        w = Engine(0, lib=wide)
        if kind == "triangles":
            w.build_triangles(elems)
        else:
            w.build_points(elems)
        _check(w.closest_triangle(q), free, f"{kind} wide margins")


def test_mesh_to_mesh_closest_pair(built):
    """tgq_closest_pair with a TRIANGLE structure as the query (BvhN::closestPair(BvhN) between two meshes, BvhN.hpp:1900-2009,
    Test/TestBvh.cpp:380-385): the minimum over query triangles of the canonical closest target triangle, every pair evaluated as
    target.squaredDistance(query); against the reference's own dual-tree traversal where its library travelled (1e-5: the reference swaps
    the roles of some pairs, and the triangle-triangle arithmetic is not symmetric in the last bit)."""
    from thea_b200 import Engine
    mesh = gen.torus_mesh(64)
    ora = orc.Oracle(tris9=mesh)
    n = ora.tri_cache()[0]
    lo, hi = gen.mesh_bounds(mesh)
    diag = float(np.linalg.norm(np.asarray(hi) - np.asarray(lo)))
    sel = np.arange(0, len(mesh), 5)
    v = mesh[sel].reshape(-1, 3, 3)
    cen = v.mean(1, keepdims=True)
    et = Engine(0)
    et.build_triangles(mesh)
    from oracle import ref
    for off in (0.03, 0.004, -0.002):
        other = (cen + 0.8 * (v - cen) + (off * diag) * n[sel][:, None, :] + np.float32([0.0007, 0.0003, -0.0005]) * diag).astype(np.float32).reshape(-1, 9)
        eq = Engine(0)
        eq.build_triangles(other)
        bi, bd, bep, bqp = ora.closest_triangle(other, plain_bound=True, mode=orc.MODE_BRUTE)
        k = np.lexsort((np.arange(len(bd)), bi, bd))[0]
        ti, qi, d2, qp, tp = et.closest_pair(eq)
        assert (ti, qi) == (int(bi[k]), int(k)) and np.float32(d2) == bd[k], (off, ti, qi, d2, int(bi[k]), int(k), float(bd[k]))
        assert np.array_equal(qp, bqp[k]) and np.array_equal(tp, bep[k])
        # bounded: nothing closer than half the minimum
        ti2, qi2, _, _, _ = et.closest_pair(eq, dist_bound=0.5 * float(np.sqrt(np.float64(bd[k]))))
        assert (ti2, qi2) == ((-1, -1) if bd[k] > 0 else (ti, qi))
        if ref.available():
            found, rti, rqi, rd2, _, _ = ref.Reference(tris9=mesh).closest_pair(ref.Reference(tris9=other))
            assert found and abs(rd2 - d2) <= 1e-5 * max(rd2, d2)
    # a point structure against the mesh as the query: the value is symmetric, indices swap roles
    pts = (gen.area_samples(mesh, 3000, seed=3)[0] + np.float32([0.02, 0.0, -0.01])).astype(np.float32)
    ep = Engine(0)
    ep.build_points(pts)
    ti, qi, d2, qp, tp = ep.closest_pair(et)                      # target = points, query = mesh
    pi, pd, pep, pqp = orc.Oracle(points=pts).closest_triangle(mesh, plain_bound=True, mode=orc.MODE_BRUTE)
    k = np.lexsort((np.arange(len(pd)), pi, pd))[0]
    assert (ti, qi) == (int(pi[k]), int(k)) and np.float32(d2) == pd[k]
