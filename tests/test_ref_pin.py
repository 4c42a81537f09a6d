"""Pins the CPU oracle (oracle/thea_oracle.hpp, a restatement) to the reference's OWN code.

oracle/_ref/libthea_ref.so is sidch/Thea's unmodified BvhN / Triangle3 / AxisAlignedBoxN / MetricL2 / Icp3 / ShapeDiameter / PointSet3
sources compiled against the Eigen stand-in of oracle/eigen_shim/ (oracle/Makefile target `ref`; the container has no Eigen).
Every comparison below is BIT EXACT unless a tolerance is stated.  What the pin still assumes is written at the top of
oracle/eigen_shim/Eigen/Core (Eigen's evaluation order of fixed-size reductions and a handful of member semantics).

Skipped when the library is neither present nor buildable (no /root/reference): the GPU box gets the prebuilt file with the snapshot.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import orc, ref
from thea_b200 import gen

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libthea_ref.so not present and /root/reference not available to build it")


def _orc_tri_point_batch(v, p):
    L = orc.lib()
    L.orc_tri_point_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    d2, cp = np.empty(len(v), np.float32), np.empty((len(v), 3), np.float32)
    L.orc_tri_point_batch(v.ctypes.data, p.ctypes.data, len(v), d2.ctypes.data, cp.ctypes.data)
    return d2, cp


def _orc_ray_tri_batch(v, o, d):
    L = orc.lib()
    L.orc_ray_tri_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    t = np.empty(len(v), np.float32)
    L.orc_ray_tri_batch(v.ctypes.data, o.ctypes.data, d.ctypes.data, len(v), t.ctypes.data)
    return t


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _same(a, b):
    """Bit equality of float arrays, NaN payloads included (degenerate triangles produce NaNs in the reference)."""
    return np.array_equal(_bits(a), _bits(b))


def _triangle_cases(rng, n):
    """Random triangles of mixed scale with a share of degenerate ones (Appendix B, Q7): collinear, two or three equal vertices,
    needle and sliver shapes, axis-aligned planes (exercise every primary axis and the first-maximum rule)."""
    v = rng.random((n, 9), dtype=np.float32)
    scale = np.float32(10.0) ** rng.integers(-3, 3, size=(n, 1)).astype(np.float32)
    v = (v * scale).astype(np.float32)
    k = n // 10
    v[0 * k:1 * k, 6:9] = v[0 * k:1 * k, 0:3] + np.float32(2.0) * (v[0 * k:1 * k, 3:6] - v[0 * k:1 * k, 0:3])      # collinear
    v[1 * k:2 * k, 3:6] = v[1 * k:2 * k, 0:3]                                                                      # v1 == v0
    v[2 * k:2 * k + k // 2, 3:6] = v[2 * k:2 * k + k // 2, 0:3]; v[2 * k:2 * k + k // 2, 6:9] = v[2 * k:2 * k + k // 2, 0:3]   # a point
    v[3 * k:4 * k, 6:9] = v[3 * k:4 * k, 3:6] + np.float32(1e-6) * rng.standard_normal((k, 3)).astype(np.float32)   # needles
    v[4 * k:5 * k, [2, 5, 8]] = np.float32(0.25)                                                                    # z = const plane
    v[5 * k:5 * k + k // 2, [0, 3, 6]] = np.float32(-1.5)                                                           # x = const plane
    return np.ascontiguousarray(v, np.float32)


def test_point_triangle_1e7_pairs_bit_exact():
    """MetricL2::closestPoints(Triangle3, point) (MetricL2.hpp:196-216 -> Triangle3.hpp:294-304,1310-1462) vs the oracle."""
    rng = np.random.default_rng(20261018)
    total = 0
    for chunk in range(5):
        n = 2_000_000
        v = _triangle_cases(rng, n)
        p = (rng.random((n, 3), dtype=np.float32) * 3 - 1).astype(np.float32)
        # a share of points ON the triangles (vertices, edge midpoints, interior) and on their planes
        k = n // 8
        p[:k] = v[:k, 0:3]
        p[k:2 * k] = (np.float32(0.5) * (v[k:2 * k, 0:3] + v[k:2 * k, 3:6])).astype(np.float32)
        w = rng.random((k, 3), dtype=np.float32); w /= w.sum(axis=1, keepdims=True)
        p[2 * k:3 * k] = (w[:, 0:1] * v[2 * k:3 * k, 0:3] + w[:, 1:2] * v[2 * k:3 * k, 3:6] + w[:, 2:3] * v[2 * k:3 * k, 6:9]).astype(np.float32)
        rd2, rcp = ref.tri_point_batch(v, p)
        od2, ocp = _orc_tri_point_batch(v, p)
        assert _same(rd2, od2), f"chunk {chunk}: {np.count_nonzero(_bits(rd2) != _bits(od2))} squared distances differ"
        assert _same(rcp, ocp), f"chunk {chunk}: closest points differ"
        total += n
    assert total == 10_000_000


def test_ray_triangle_1e7_pairs_bit_exact():
    """Triangle3::rayIntersectionTime (Triangle3.hpp:536-540,1464-1524) vs the oracle, incl. rays aimed at the triangles."""
    rng = np.random.default_rng(7)
    hits = 0
    for chunk in range(5):
        n = 2_000_000
        v = _triangle_cases(rng, n)
        o = (rng.random((n, 3), dtype=np.float32) * 3 - 1).astype(np.float32)
        d = rng.standard_normal((n, 3)).astype(np.float32)
        k = n // 2
        w = rng.random((k, 3), dtype=np.float32); w /= w.sum(axis=1, keepdims=True)
        tgt = (w[:, 0:1] * v[:k, 0:3] + w[:, 1:2] * v[:k, 3:6] + w[:, 2:3] * v[:k, 6:9]).astype(np.float32)
        d[:k] = (tgt - o[:k]) * rng.random((k, 1), dtype=np.float32) * np.float32(2.0)       # aimed at the triangle, any length
        d[k:k + 1000] = (v[k:k + 1000, 3:6] - v[k:k + 1000, 0:3])                           # parallel to the plane
        rt = ref.ray_tri_batch(v, o, d)
        ot = _orc_ray_tri_batch(v, o, d)
        assert _same(rt, ot), f"chunk {chunk}: {np.count_nonzero(_bits(rt) != _bits(ot))} hit times differ"
        hits += int((rt >= 0).sum())
    assert hits > 1_000_000      # the aimed half hits unless the triangle is degenerate or below the absolute epsilon


def test_fast_sqrt_random_boxes():
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.random(200_000, dtype=np.float32) * np.float32(10.0) ** rng.integers(-12, 12, 200_000).astype(np.float32),
                         np.array([0.0, 1.0, 2.0, 4.0, 1e-30, 1e30, 0.25, 3.0], np.float32)])
    for x in xs[:20000]:
        assert _same(np.float32(ref.fast_sqrt(float(x))), np.float32(orc.fast_sqrt(float(x)))), x
    for seed in (0xFFFF, 1, 0xB2000001):
        assert np.array_equal(ref.random_uniform01(seed, 100_000), orc.thea_random_uniform01(seed, 100_000))
    L = orc.lib()
    L.orc_box_point_sqdist.restype = C.c_float
    L.orc_box_point_sqdist.argtypes = [C.c_void_p] * 3
    for _ in range(5000):
        lo = rng.random(3, dtype=np.float32); hi = lo + rng.random(3, dtype=np.float32)
        p = (rng.random(3, dtype=np.float32) * 3 - 1).astype(np.float32)
        o = (rng.random(3, dtype=np.float32) * 3 - 1).astype(np.float32); d = rng.standard_normal(3).astype(np.float32)
        if rng.random() < 0.2:
            d[rng.integers(0, 3)] = 0.0
        assert _same(np.float32(ref.box_point_sqdist(lo, hi, p)), np.float32(L.orc_box_point_sqdist(lo.ctypes.data, hi.ctypes.data, p.ctypes.data)))
        for mt in (-1.0, 0.7):
            assert _same(np.float32(ref.box_ray_time(lo, hi, o, d, mt)), np.float32(orc.box_ray_time(lo, hi, o, d, mt)))
    for n in list(range(0, 200)) + [1000, 4608, 70688, 1002528, 9999392]:
        assert ref.tree_depth(n, 5) == orc.lib().orc_tree_depth(C.c_int64(n), 5), n


def _testbvh_inputs():
    """Test/TestBvh.cpp:120-126,315-321: 10,000 points then (in a second run of the generator) 1,000 triangles in [0,1]^3."""
    u = ref.random_uniform01(0xFFFF, 30000 + 30000 + 9000)
    return u[:30000].reshape(-1, 3).copy(), u[60000:].reshape(-1, 9).copy()


@pytest.fixture(scope="module")
def torus():
    mesh = gen.torus_mesh(48)
    return mesh, gen.mesh_bounds(mesh), ref.Reference(tris9=mesh), orc.Oracle(tris9=mesh)


def test_closest_element_structure(torus):
    """BvhN::closestPair / closestElement on the kd-tree the reference builds vs the oracle's reference-order mode: same element
    (first-found ties included), same squared distance, same closest point; canonical / brute-force modes: same distance always,
    same element unless the reference's answer is an exact tie."""
    mesh, bounds, R, O = torus
    q = gen.closest_queries(60000, bounds)
    for bound in (-1.0, 0.35):
        ri, rd, rc = R.closest(q, dist_bound=bound)
        oi, od, oc = O.closest(q, dist_bound=bound, mode=orc.MODE_REFERENCE)
        assert np.array_equal(ri, oi)
        found = ri >= 0
        assert _same(rd[found], od[found]) and _same(rc[found], oc[found])
        ci, cd, _ = O.closest(q, dist_bound=bound, mode=orc.MODE_BRUTE)
        assert np.array_equal(ci >= 0, found) and _same(cd[found], rd[found])
        differ = np.flatnonzero(ci != ri)
        # every disagreement is an exact tie: the brute-force winner has the same distance and a LOWER index
        assert np.all(ci[differ] < ri[differ])
    ei, ed, ec = R.closest_element(q[:5000])
    ri, rd, rc = R.closest(q[:5000])
    assert np.array_equal(ei, ri) and np.array_equal(ed, np.sqrt(rd.astype(np.float64))) and _same(ec, rc)
    assert R.num_nodes() == orc.lib().orc_num_nodes(O.h)


def test_testbvh_recipe_structures():
    pts, tris = _testbvh_inputs()
    R, O = ref.Reference(tris9=tris), orc.Oracle(tris9=tris)
    q = np.vstack([np.float32([[0.5, 0.5, 0.5]]), gen.closest_queries(20000, (np.zeros(3, np.float32), np.ones(3, np.float32)))])
    for bound in (-1.0, 0.25):
        ri, rd, rc = R.closest(q, dist_bound=bound)
        oi, od, oc = O.closest(q, dist_bound=bound, mode=orc.MODE_REFERENCE)
        assert np.array_equal(ri, oi) and _same(rd[ri >= 0], od[ri >= 0]) and _same(rc[ri >= 0], oc[ri >= 0])
    RP, OP = ref.Reference(points=pts), orc.Oracle(points=pts)
    for k in (1, 3, 16):
        for bound in (-1.0, 0.25, 0.02):
            ri, rd = RP.knn(q[:5000], k, dist_bound=bound)
            oi, od = OP.knn(q[:5000], k, dist_bound=bound, mode=orc.MODE_REFERENCE)
            assert np.array_equal(ri, oi) and _same(rd, od), (k, bound)
    ri, rd, rc = RP.closest(q, dist_bound=-1.0)
    oi, od, oc = OP.closest(q, dist_bound=-1.0, mode=orc.MODE_REFERENCE)
    assert np.array_equal(ri, oi) and _same(rd, od) and _same(rc, oc)
    # ball range: the same SET of indices (the oracle returns the lowest `max_out` in ascending order)
    for c, r in ((q[0], 0.25), (q[1], 0.05), (q[2], 0.11)):
        ridx, n = RP.ball_range(c, r)
        oidx, ocnt = OP.ball_range(c[None], r, max_out=32)
        assert n == int(ocnt[0]) and len(set(ridx.tolist())) == n
        assert np.array_equal(np.sort(ridx)[:32], oidx[0][:min(32, n)])


def test_ray_structure(torus):
    mesh, bounds, R, O = torus
    o, d = gen.random_rays(60000, bounds)
    ri, rt, rn = R.ray_first(o, d, normals=True)
    oi, ot = O.ray_first(o, d, mode=orc.MODE_REFERENCE)
    assert np.array_equal(ri, oi) and _same(rt, ot)
    assert (ri >= 0).sum() > 5000
    n, _, _, _ = O.tri_cache()
    hit = ri >= 0
    assert _same(rn[hit], n[ri[hit]])
    tmax = np.where(np.arange(len(o)) % 2 == 0, np.float32(0.8), np.float32(-1.0)).astype(np.float32)
    ri, rt = R.ray_first(o, d, tmax=tmax)
    oi, ot = O.ray_first(o, d, tmax=tmax, mode=orc.MODE_REFERENCE)
    assert np.array_equal(ri, oi) and _same(rt, ot)
    ci, ct = O.ray_first(o, d, tmax=tmax, mode=orc.MODE_BRUTE)
    assert np.array_equal(ci >= 0, ri >= 0) and _same(ct, rt) and np.all(ci[ci != ri] < ri[ci != ri])
    t2, h2 = R.ray_time(o[:4000], d[:4000], tmax=tmax[:4000])
    assert _same(t2, rt[:4000]) and np.array_equal(h2.astype(bool), ri[:4000] >= 0)


def test_triangle_cache(torus):
    mesh, _, _, O = torus
    n, d, ax, ar = O.tri_cache()
    for i in range(0, len(mesh), 7):
        rn, _, rax, rar, _ = ref.tri_cache(mesh[i])
        assert _same(rn, n[i]) and rax == ax[i] and _same(np.float32(rar), ar[i])


def test_shape_diameter(torus):
    """ShapeDiameter::compute(position, normal, interior) (ShapeDiameter.hpp:168-258), double statistics, nth_element quirk."""
    mesh, bounds, R, O = torus
    p, n, _ = gen.area_samples(mesh, 1500)
    scale = float(np.linalg.norm(bounds[1] - bounds[0]))
    for interior in (True, False):
        rs = R.sdf(p, n, scale, only_interior=interior)
        os_ = O.sdf(p, n, scale, only_interior=interior, mode=orc.MODE_REFERENCE)[0]
        assert np.array_equal(rs, os_), np.abs(rs - os_).max()
    assert (rs >= 0).sum() > 1000


def test_sample_graph(torus):
    """PointSet3::findSampleNeighbors (PointSet3.hpp:586-681): neighbour indices and fastSqrt separations, with and without the
    reachability ray test against the surface; the average-separation estimate for given seed samples (:363-376)."""
    mesh, bounds, R, O = torus
    p, n, _ = gen.area_samples(mesh, 3000)
    OP = orc.Oracle(points=p, normals=n)
    seeds = (gen.splitmix64(gen.SEED_SAMPLES, 100) % np.uint64(len(p))).astype(np.int64)
    sep = ref.avg_separation(p, seeds)
    assert _same(np.float32(sep), np.float32(OP.avg_separation(seeds)))
    for k in (8, 16):
        r0 = np.float32(np.sqrt(np.float32(k))) * np.float32(sep)
        ri, rs, rc = ref.graph(p, n, k, float(r0))
        oi, os_, oc = OP.ball_graph(k, float(r0), mode=orc.MODE_REFERENCE)
        assert np.array_equal(rc, oc) and np.array_equal(ri, oi) and _same(rs, os_)
    ri, rs, rc = ref.graph(p, n, 8, float(np.float32(np.sqrt(np.float32(8))) * np.float32(sep)), surface=R)
    oi, os_, oc = OP.ball_graph_surface(O, 8, float(np.float32(np.sqrt(np.float32(8))) * np.float32(sep)), mode=orc.MODE_REFERENCE)
    assert np.array_equal(rc, oc) and np.array_equal(ri, oi) and _same(rs, os_)
    # without normals the angle and reachability tests are skipped
    OP2 = orc.Oracle(points=p)
    ri, rs, rc = ref.graph(p, None, 8, float(np.float32(np.sqrt(np.float32(8))) * np.float32(sep)))
    oi, os_, oc = OP2.ball_graph(8, float(np.float32(np.sqrt(np.float32(8))) * np.float32(sep)), mode=orc.MODE_REFERENCE)
    assert np.array_equal(rc, oc) and np.array_equal(ri, oi) and _same(rs, os_)


def test_icp(torus):
    """Icp3<Real>::align (Icp3.hpp:193-464).  Everything but the 3 x 3 SVD is the reference's code; the SVD comes from the Eigen
    stand-in, so the comparison uses the tolerance north_star states for floats (1e-5) instead of bit equality."""
    mesh, bounds, R, O = torus
    p, _, _ = gen.area_samples(mesh, 4000)
    src = gen.apply_pose(p, *gen.rigid_pose())
    rtr, rerr = R.icp(src, min_iters=8, max_iters=8, frac_thresh=1e-4)
    otr, oerr, _ = O.icp(src, min_iters=8, max_iters=8, frac_thresh=1e-4, mode=orc.MODE_REFERENCE)
    assert np.allclose(rtr, otr, rtol=1e-5, atol=1e-5), np.abs(rtr - otr).max()
    assert abs(rerr - oerr) <= 1e-5 * max(1.0, abs(rerr))


def test_icp_accumulation_loops_bit_exact(torus):
    """The two sequential fp32 loops of Icp3 with given correspondences: measureError (Icp3.hpp:418-464) bit for bit; alignOneStep
    (:298-384) to 1e-6 (its rotation goes through the stand-in's SVD on the reference side)."""
    mesh, bounds, R, O = torus
    p, _, _ = gen.area_samples(mesh, 20000)
    src = gen.apply_pose(p, *gen.rigid_pose())
    _, _, dst = O.closest(src, mode=orc.MODE_CANON)
    Rm, t = gen.rigid_pose()
    for tr in (np.hstack([np.eye(3, dtype=np.float32), np.zeros((3, 1), np.float32)]), np.hstack([Rm.astype(np.float32), t.reshape(3, 1).astype(np.float32)])):
        assert _same(np.float32(ref.icp_error(tr, src, dst)), np.float32(orc.icp_error(tr, src, dst)))
    a, b = ref.icp_step(src, dst), orc.icp_step(src, dst)
    assert np.allclose(a, b, rtol=1e-6, atol=1e-6), np.abs(a - b).max()


def test_spin_image():
    """oracle spin image == the reference's SurfaceFeatures::Local::SpinImage on a PointSet3 with the scale set explicitly."""
    mesh = gen.torus_mesh(40)
    p, n, _ = gen.area_samples(mesh, 3000)
    ora = orc.Oracle(points=p, normals=n)
    rng = np.random.default_rng(1)
    q = (p[:300] + rng.normal(scale=0.02, size=(300, 3))).astype(np.float32)
    qn = n[:300].copy(); qn[::7] *= np.float32(3.0)
    scale = np.float32(2.9)
    for nr, nh in ((9, 7), (15, 15)):
        assert np.array_equal(ora.spin_image(q, qn, nr, nh, np.float32(0.75) * scale), ref.spin_image(p, n, scale, q, qn, nr, nh))


def test_sample_graph_shortest_paths():
    """The oracle's restatement of the ShortestPaths consumers == the reference's own code (PointSet3::extractOriginalAdjacencies,
    AverageDistance / LocalDistanceHistogram GEODESIC, FurthestPointSampling's selection around ShortestPaths) on an injected graph.
    Bit-equal: indices, float separations, histogram counts, selections; the average distance to 1e-15 (same sums, one division)."""
    mesh = gen.torus_mesh(40)
    p, n, _ = gen.area_samples(mesh, 3000)
    ora = orc.Oracle(points=p, normals=n)
    seeds = (gen.splitmix64(7, 100) % np.uint64(len(p))).astype(np.int64)
    avg = np.float32(ora.avg_separation(seeds))
    for K in (8, 5):
        gi, gs, gc = ora.ball_graph(K, np.float32(np.sqrt(np.float32(K))) * avg)
        a = orc.graph_contract(gi, gs, 700)
        b = ref.graph_contract(p, 700, gi, gs)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        sd = np.arange(0, 3000, 37)
        scale = np.float32(2.9)
        for md in (np.float32(0.3), np.float32(-1.0)):
            rng_ = md if md > 0 else scale
            s_, c_, h_ = orc.graph_geodesic(gi, gs, sd, md, rng_, 12)
            ravg, rh = ref.geodesic_features(p, gi, gs, scale, p[sd], md, 12)
            mine = np.where(c_ > 0, s_ / np.maximum(c_, 1), 0.0) / np.float64(rng_)
            assert np.allclose(mine, ravg, rtol=1e-15, atol=0) and np.array_equal(h_, rh)
        assert np.array_equal(orc.graph_fps(gi, gs, 80), ref.fps(p, gi, gs, 80))


def test_k_closest_pairs_between_point_structures():
    """BvhN::kClosestPairs with a structure as the query (BvhN.hpp:1137-1146,1900-2009) == the k smallest (d2, target, query) over the
    oracle's brute-force k-NN rows of the query points: what tgq_k_closest_pairs computes on the device."""
    mesh = gen.torus_mesh(30)
    a, _, _ = gen.area_samples(mesh, 3000)
    b, _, _ = gen.area_samples(mesh, 2000, seed=9)
    b = (b + np.float32(0.01)).astype(np.float32)
    R, Q = ref.Reference(points=a), ref.Reference(points=b)
    ora = orc.Oracle(points=a)
    for k, bound in ((12, -1.0), (5, -1.0), (16, 0.0052)):
        ti, qi, d2 = R.k_closest_pairs(Q, k, bound)
        ki, kd = ora.knn(b, k, dist_bound=bound, mode=orc.MODE_BRUTE)
        rows = sorted((float(kd[q, j]), int(ki[q, j]), q) for q in range(len(b)) for j in range(k) if ki[q, j] >= 0)[:k]
        assert [(int(t), int(q)) for t, q in zip(ti, qi)] == [(t, q) for _, t, q in rows]
        assert np.array_equal(d2, np.array([r[0] for r in rows], np.float32))
    assert 0 < len(R.k_closest_pairs(Q, 16, 0.0052)[0]) < 16          # the bound bites


def _segment_cases(rng, n, lo=-1.0, hi=1.0):
    """Segments of mixed length (some of zero length, some built from far apart endpoints) in the box [lo, hi]^3."""
    a = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.standard_normal((n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    b = (a + d * rng.choice([1e-4, 1e-2, 0.3, 2.0], (n, 1))).astype(np.float32)
    b[::97] = a[::97]
    return a, b


def test_line_segment_query_arithmetic_matches_reference():
    """SURVEY 8f rank 4, per-pair: MetricL2::closestPoints(point, LineSegment3) -> LineSegmentN::closestPoint (LineSegmentN.hpp:106-145)
    and MetricL2::closestPoints(Triangle3, LineSegment3) -> TriangleN<3>::squaredDistance(LineSegmentN) (Triangle3.hpp:495-509,559-577,
    LineSegmentN.hpp:262-357), called exactly as BvhN::processNeighborsLeaf calls them (BvhN.hpp:2063), 10^6 pairs each: squared distance
    and both closest points bit for bit, degenerate segments and triangles included."""
    rng = np.random.default_rng(21)
    n = 1_000_000
    a, b = _segment_cases(rng, n)
    p = rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)
    p[::53] = a[::53]
    d_o, s_o = orc.seg_point(a, b, p)
    d_r, s_r = ref.seg_point_batch(a, b, p)
    assert _same(d_o, d_r) and _same(s_o, s_r)
    v = _triangle_cases(rng, n)
    a, b = _segment_cases(rng, n, 0.0, 1.0)
    a = (a * np.float32(10.0) ** rng.integers(-3, 3, size=(n, 1))).astype(np.float32)
    # a share of segments that pierce their triangle (the ray-test branch) and of segments lying in the triangle's plane
    k = n // 4
    c = (v[:k, 0:3] + v[:k, 3:6] + v[:k, 6:9]) / np.float32(3.0)
    nrm = np.cross(v[:k, 3:6] - v[:k, 0:3], v[:k, 6:9] - v[:k, 0:3])
    a[:k] = (c + nrm * np.float32(0.7)).astype(np.float32)
    b[:k] = (c - nrm * rng.uniform(0.2, 1.5, (k, 1)) + rng.standard_normal((k, 3)) * 1e-3 * np.abs(c)).astype(np.float32)
    a[k:k + k // 2] = v[k:k + k // 2, 0:3] * np.float32(1.5) - v[k:k + k // 2, 3:6] * np.float32(0.5)       # in the plane, through v0
    d_o, t_o, s_o = orc.tri_seg(v, a, b)
    d_r, t_r, s_r = ref.tri_seg_batch(v, a, b)
    assert _same(d_o, d_r) and _same(t_o, t_r) and _same(s_o, s_r)
    assert (d_r == 0).sum() > n // 10 and (d_r > 0).sum() > n // 3       # both branches are exercised


@pytest.mark.parametrize("kind", ["triangles", "points"])
def test_line_segment_queries_match_reference_bvh(kind):
    """BvhN::closestPair<MetricL2>(LineSegment3, dist_bound, compat, true) on Thea's own BvhN (a bounded query object: box-box node
    bounds, BvhN.hpp:1794-1803; root rejection and bound widening :1081-1114), with and without a pushed element filter that rejects
    elements by identity (Tools/MeshFix/MeshFix.cpp:678-697,756-763), against the oracle's reference-order traversal: index, squared
    distance and both closest points bit for bit; the canonical answer equals brute force."""
    rng = np.random.default_rng(22)
    Q = 3000
    if kind == "triangles":
        elems = gen.torus_mesh(24)
        lo, hi = gen.mesh_bounds(elems)
        a, b = _segment_cases(rng, Q, float(np.min(lo)) * 1.3, float(np.max(hi)) * 1.3)
        R, ora, n, bound = ref.Reference(tris9=elems), orc.Oracle(tris9=elems), len(elems), 0.15
    else:
        elems = rng.uniform(-1, 1, (20000, 3)).astype(np.float32)
        elems[5000:5200] = elems[4000:4200]                     # duplicates: exact ties
        a, b = _segment_cases(rng, Q, -1.2, 1.2)
        R, ora, n, bound = ref.Reference(points=elems), orc.Oracle(points=elems), len(elems), 0.05
    excl = rng.integers(0, n, (Q, 3))
    excl[::4] = -1
    # every third query excludes what it would otherwise have found (the filter must change the answer)
    first = R.closest_segments(a, b)[0]
    excl[::3, 0] = first[::3]
    for db in (-1.0, bound):
        for ex in (None, excl):
            r = R.closest_segments(a, b, db, ex)
            o = ora.closest_segment(a, b, db, ex, mode=orc.MODE_REFERENCE)
            assert np.array_equal(r[0], o[0]) and _same(r[1], o[1]) and _same(r[2], o[2]) and _same(r[3], o[3])
            c = ora.closest_segment(a, b, db, ex, mode=orc.MODE_CANON)
            bf = ora.closest_segment(a, b, db, ex, mode=orc.MODE_BRUTE)
            assert np.array_equal(c[0], bf[0]) and _same(c[1], bf[1]) and _same(c[1], o[1])
            if ex is not None:
                assert not (r[0][::3] == first[::3]).any() or (first[::3] < 0).any()
    assert 0 < (R.closest_segments(a, b, bound)[0] >= 0).sum() < Q          # the bound bites


def _triangle_pairs(rng, n):
    """Pairs of triangles of mixed size and separation: overlapping, far apart, coplanar, sharing a vertex or an edge (mesh neighbours),
    identical, degenerate."""
    a = rng.uniform(-1, 1, (n, 9))
    size = rng.choice([0.05, 0.3, 1.0], (n, 1))
    a[:, 3:6] = a[:, 0:3] + rng.normal(0, 1, (n, 3)) * size
    a[:, 6:9] = a[:, 0:3] + rng.normal(0, 1, (n, 3)) * size
    b = rng.uniform(-1, 1, (n, 9))
    b[:, 0:3] = a[:, 0:3] + rng.normal(0, 1, (n, 3)) * rng.choice([0.0, 0.02, 0.2, 1.0], (n, 1))
    b[:, 3:6] = b[:, 0:3] + rng.normal(0, 1, (n, 3)) * size
    b[:, 6:9] = b[:, 0:3] + rng.normal(0, 1, (n, 3)) * size
    a, b = a.astype(np.float32), b.astype(np.float32)
    k = n // 10
    a[:k, [2, 5, 8]] = 0.25; b[:k, [2, 5, 8]] = 0.25                                  # coplanar
    b[k:2 * k, 0:3] = a[k:2 * k, 0:3]                                                 # shared vertex
    b[2 * k:3 * k, 0:3] = a[2 * k:3 * k, 3:6]; b[2 * k:3 * k, 3:6] = a[2 * k:3 * k, 0:3]    # shared edge
    b[3 * k:3 * k + k // 4] = a[3 * k:3 * k + k // 4]                                 # identical
    a[4 * k:4 * k + k // 4, 6:9] = a[4 * k:4 * k + k // 4, 3:6]                       # degenerate
    b[5 * k:5 * k + k // 4, 3:6] = b[5 * k:5 * k + k // 4, 0:3]
    return np.ascontiguousarray(a), np.ascontiguousarray(b)


def test_triangle_triangle_distance_matches_reference():
    """SURVEY 8f rank 4, per pair: MetricL2::closestPoints(Triangle3 a, Triangle3 b) = a.squaredDistance(b, &pa, &pb) (Triangle3.hpp:314-392
    with Moeller's tri_tri_intersect_with_isectline, :597-1300) on 10^6 pairs: the oracle's restatement AND the engine's own headers (host
    build, tgq_math.cuh: tri_tri_sqdist) both return the reference's squared distance and closest points bit for bit."""
    import ctypes as C
    import emu_lib
    rng = np.random.default_rng(23)
    n = 1_000_000
    a, b = _triangle_pairs(rng, n)
    r = ref.tri_tri_batch(a, b)
    o = orc.tri_tri(a, b)
    assert _same(r[0], o[0]) and _same(r[1], o[1]) and _same(r[2], o[2])
    L = emu_lib.lib()
    L.emu_tri_tri_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    d2, pa, pb = np.empty(n, np.float32), np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
    L.emu_tri_tri_batch(a.ctypes.data, b.ctypes.data, n, d2.ctypes.data, pa.ctypes.data, pb.ctypes.data)
    assert _same(r[0], d2) and _same(r[1], pa) and _same(r[2], pb)
    assert (r[0] == 0).sum() > n // 5 and (r[0] > 0).sum() > n // 5


def _query_triangles(rng, Q, lo, hi, sizes):
    c = rng.uniform(lo, hi, (Q, 1, 3))
    return (c + rng.normal(0, 1, (Q, 3, 3)) * rng.choice(sizes, (Q, 1, 1))).reshape(Q, 9).astype(np.float32)


@pytest.mark.parametrize("kind", ["triangles", "points"])
def test_triangle_queries_match_reference_bvh(kind):
    """BvhN::closestPair<MetricL2>(Triangle3, dist_bound, compat, true) on Thea's own BvhN against the oracle's reference-order traversal:
    index, squared distance and both closest points bit for bit; canonical == brute force."""
    rng = np.random.default_rng(24)
    Q = 2000
    if kind == "triangles":
        elems = gen.torus_mesh(24)
        R, ora, q, bound = ref.Reference(tris9=elems), orc.Oracle(tris9=elems), _query_triangles(rng, Q, -1.6, 1.6, [0.01, 0.05, 0.3]), 0.2
    else:
        elems = rng.uniform(-1, 1, (20000, 3)).astype(np.float32)
        R, ora, q, bound = ref.Reference(points=elems), orc.Oracle(points=elems), _query_triangles(rng, Q, -1.2, 1.2, [0.01, 0.1, 0.5]), 0.05
    for db in (-1.0, bound):
        r = R.closest_triangles(q, db)
        o = ora.closest_triangle(q, db, mode=orc.MODE_REFERENCE)
        assert np.array_equal(r[0], o[0]) and _same(r[1], o[1]) and _same(r[2], o[2]) and _same(r[3], o[3])
        c = ora.closest_triangle(q, db, mode=orc.MODE_CANON)
        bf = ora.closest_triangle(q, db, mode=orc.MODE_BRUTE)
        assert np.array_equal(c[0], bf[0]) and _same(c[1], bf[1]) and _same(c[1], o[1])
    assert 0 < (r[0] >= 0).sum() < Q


def test_mesh_to_mesh_closest_pair_matches_reference():
    """BvhN::closestPair(BvhN) between two MESHES (BvhN.hpp:1900-2009, Test/TestBvh.cpp:380-385) against the definition the engine
    computes: the minimum over query triangles of the canonical closest target triangle, every pair as target.squaredDistance(query)."""
    mesh = gen.torus_mesh(40)
    ora = orc.Oracle(tris9=mesh)
    n = ora.tri_cache()[0]
    lo, hi = gen.mesh_bounds(mesh)
    diag = float(np.linalg.norm(np.asarray(hi) - np.asarray(lo)))
    sel = np.arange(0, len(mesh), 5)
    v = mesh[sel].reshape(-1, 3, 3)
    cen = v.mean(1, keepdims=True)
    for off in (0.03, 0.004, -0.002):
        other = (cen + 0.8 * (v - cen) + (off * diag) * n[sel][:, None, :] + np.float32([0.0007, 0.0003, -0.0005]) * diag).astype(np.float32).reshape(-1, 9)
        found, ti, qi, d2, _, _ = ref.Reference(tris9=mesh).closest_pair(ref.Reference(tris9=other))
        bi, bd, _, _ = ora.closest_triangle(other, plain_bound=True, mode=orc.MODE_BRUTE)
        k = np.lexsort((np.arange(len(bd)), bi, bd))[0]
        assert found and abs(d2 - float(bd[k])) <= 1e-5 * max(d2, float(bd[k]))       # role swapping in the reference: last-bit differences allowed
        if d2 > 0:
            assert (ti, qi) == (int(bi[k]), int(k))


def test_random_walks_match_reference():
    """SurfaceFeatures::Local::RandomWalks::compute (RandomWalks.cpp:30-117), the reference's own code on an injected sample graph with its
    shared generator Random::common() put into the state of Random(seed), against the oracle: every feature double bit for bit, including
    walks that stop at a sample without neighbours (0 / 0 = NaN where no walk got that far)."""
    mesh = gen.torus_mesh(30)
    p, n, _ = gen.area_samples(mesh, 3000)
    ora = orc.Oracle(points=p, normals=n)
    seeds = (gen.splitmix64(0x5EED, 100) % np.uint64(len(p))).astype(np.int64)
    K = 8
    gi, gs, _ = ora.ball_graph(K, float(np.float32(np.sqrt(np.float32(K))) * np.float32(ora.avg_separation(seeds))))
    gi[17] = -1
    gs[17] = -1
    q = (p[::100] + np.float32(0.01)).astype(np.float32)
    q[3] = p[17]
    for steps, walks, seed in ((5, 50, 1234), (8, 1000, 99)):
        a = orc.Oracle(points=p).random_walks(gi, q, steps, walks, seed)
        b = ref.random_walks(p, gi, gs, q, steps, walks, seed)
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)].view(np.uint64), b[~np.isnan(b)].view(np.uint64))
        assert np.isnan(b).any() and np.isfinite(b).sum() > b.size // 2
