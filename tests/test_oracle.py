"""The CPU oracle against analytic known answers, the reference's own test recipes, and itself (kd-tree vs brute force).

The reference's tests on this path assert nothing (Test/TestBvh.cpp:26-40, Test/TestMetrics.cpp:33-84 only print), so
parity is unpinned; what can be pinned is checked here: closed-form answers for every Voronoi region of a triangle, the
TestMetrics triangle against an independent float64 solution, Thea::Random's MT19937 quirk against numpy's MT19937, and
the equivalence "reference traversal order == canonical == brute force" up to exact ties.
"""
import numpy as np
import pytest

from oracle import orc
from thea_b200 import gen

TRI = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0], np.float32)


def closest_on_triangle_f64(a, b, c, p):
    """Ericson, Real-Time Collision Detection 5.1.5, in float64 (independent of the reference's algorithm)."""
    a, b, c, p = (np.asarray(x, np.float64) for x in (a, b, c, p))
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = ab @ ap, ac @ ap
    if d1 <= 0 and d2 <= 0:
        return a
    bp = p - b
    d3, d4 = ab @ bp, ac @ bp
    if d3 >= 0 and d4 <= d3:
        return b
    vc = d1 * d4 - d3 * d2
    if vc <= 0 and d1 >= 0 and d3 <= 0:
        return a + ab * (d1 / (d1 - d3))
    cp = p - c
    d5, d6 = ab @ cp, ac @ cp
    if d6 >= 0 and d5 <= d6:
        return c
    vb = d5 * d2 - d1 * d6
    if vb <= 0 and d2 >= 0 and d6 <= 0:
        return a + ac * (d2 / (d2 - d6))
    va = d3 * d6 - d5 * d4
    if va <= 0 and (d4 - d3) >= 0 and (d5 - d6) >= 0:
        return b + (c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6)))
    den = 1.0 / (va + vb + vc)
    return a + ab * (vb * den) + ac * (vc * den)


@pytest.mark.parametrize("p, d2, cp", [
    ((0.25, 0.25, 2.0), 4.0, (0.25, 0.25, 0.0)),      # face region
    ((0.25, 0.25, -3.0), 9.0, (0.25, 0.25, 0.0)),     # face region, other side
    ((-1.0, -1.0, 0.0), 2.0, (0.0, 0.0, 0.0)),        # vertex v0
    ((3.0, -1.0, 0.0), 5.0, (1.0, 0.0, 0.0)),         # vertex v1
    ((-1.0, 3.0, 0.0), 5.0, (0.0, 1.0, 0.0)),         # vertex v2
    ((0.5, -1.0, 0.0), 1.0, (0.5, 0.0, 0.0)),         # edge v0-v1
    ((1.0, 1.0, 0.0), 0.5, (0.5, 0.5, 0.0)),          # edge v1-v2
    ((-2.0, 0.5, 0.0), 4.0, (0.0, 0.5, 0.0)),         # edge v2-v0
    ((0.5, -1.0, 1.0), 2.0, (0.5, 0.0, 0.0)),         # edge region off-plane
    ((0.0, 0.0, 0.0), 0.0, (0.0, 0.0, 0.0)),          # on a vertex
])
def test_point_triangle_known_answers(p, d2, cp):
    got_d2, got_cp = orc.tri_point(TRI, p)
    assert got_d2 == pytest.approx(d2, rel=1e-6, abs=1e-12)
    assert np.allclose(got_cp, cp, atol=1e-6)


def test_testmetrics_triangle():
    # Test/TestMetrics.cpp:50-52
    v = np.array([-1, -1, -1, 0.5, 0.5, 0, 0, 0.5, 0.5], np.float32)
    p = np.array([2, 2, 2], np.float32)
    d2, cp = orc.tri_point(v, p)
    ref = closest_on_triangle_f64(v[0:3], v[3:6], v[6:9], p)
    assert np.allclose(cp, ref, atol=1e-6)
    assert d2 == pytest.approx(float(((ref - p.astype(np.float64)) ** 2).sum()), rel=1e-6)


def test_point_triangle_random_vs_float64():
    rng = np.random.default_rng(1)
    for _ in range(3000):
        v = rng.random(9).astype(np.float32)
        p = (rng.random(3) * 3 - 1).astype(np.float32)
        d2, cp = orc.tri_point(v, p)
        ref = closest_on_triangle_f64(v[0:3], v[3:6], v[6:9], p)
        true_d2 = float(((ref - p.astype(np.float64)) ** 2).sum())
        # the reference's algorithm (plane projection + 2-D barycentric test + perimeter) is fp32; slivers lose digits
        assert d2 == pytest.approx(true_d2, rel=2e-3, abs=1e-6)


@pytest.mark.parametrize("o, d, t", [
    ((0.25, 0.25, 1.0), (0, 0, -1), 1.0),        # front face
    ((0.25, 0.25, -2.0), (0, 0, 1), 2.0),        # back face counts too (sign flip, Triangle3.hpp:1479-1487)
    ((0.25, 0.25, 1.0), (0, 0, -0.5), 2.0),      # time is in units of the (unnormalised) direction
    ((0.25, 0.25, 1.0), (0, 0, 1), -1.0),        # behind the origin
    ((0.25, 0.25, 1.0), (1, 0, 0), -1.0),        # parallel
    ((2.0, 2.0, 1.0), (0, 0, -1), -1.0),         # misses the triangle
    ((0.5, 0.0, 1.0), (0, 0, -1), 1.0),          # exactly through an edge: inclusive
    ((0.0, 0.0, 1.0), (0, 0, -1), 1.0),          # exactly through a vertex: inclusive
    ((0.25, 0.25, 0.0), (0, 0, -1), 0.0),        # origin on the triangle: t = 0 is a hit
])
def test_ray_triangle_known_answers(o, d, t):
    assert orc.ray_tri(TRI, o, d) == pytest.approx(t, abs=1e-6)


def test_ray_triangle_absolute_epsilon_quirk():
    # Q5: |D.N| <= FLT_EPSILON on the UNNORMALISED normal is a miss, so a tiny triangle is invisible (Triangle3.hpp:1471-1492)
    tiny = (TRI * np.float32(1e-4)).astype(np.float32)       # |N| = 1e-8 < FLT_EPSILON
    assert orc.ray_tri(tiny, (2e-5, 2e-5, 1.0), (0, 0, -1)) == -1.0
    assert orc.ray_tri(TRI, (0.2, 0.2, 1.0), (0, 0, -1)) == 1.0


def test_fast_sqrt_bit_hack():
    # Math.hpp:194-223 restated independently with numpy scalars
    for x in (4.0, 1e-3, 0.37, 123.0, 1e-12, 0.0):
        xf = np.float32(x)
        i = np.frombuffer(xf.tobytes(), np.int32)[0]
        i = np.int32(0x5f3759df) - (i >> 1)
        y = np.frombuffer(np.int32(i).tobytes(), np.float32)[0]
        y = y * (np.float32(1.5) - (xf * np.float32(0.5) * y * y))
        assert orc.fast_sqrt(x) == float(np.float32(1.0) / y)
    assert orc.fast_sqrt(4.0) == pytest.approx(2.0, rel=2e-3)


def test_box_ray_woo_quirks():
    lo, hi = (0, 0, 0), (1, 1, 1)
    assert orc.box_ray_time(lo, hi, (0.5, 0.5, 0.5), (1, 0, 0)) == 0.0                       # inside => 0 (Q6)
    assert orc.box_ray_time(lo, hi, (-1, 0.5, 0.5), (1, 0, 0)) == pytest.approx(1.0)
    assert orc.box_ray_time(lo, hi, (-1, 0.5, 0.5), (-1, 0, 0)) == -1.0                      # pointing away
    assert orc.box_ray_time(lo, hi, (-1, 0.5, 0.5), (1, 0, 0), max_time=0.5) == -1.0         # early out on max_time
    assert orc.box_ray_time(lo, hi, (-1, 2.0, 0.5), (1, 0, 0)) == -1.0


def test_thea_random_is_mt19937_with_the_g3d_skip():
    # Random.cpp:113-125: the second regeneration loop starts at N-M+1, skipping state[227]
    n = 400
    got = orc.thea_random_uniform01(0xFFFF, n)
    rs = np.random.RandomState(0xFFFF)
    bits = rs.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    std = bits.astype(np.float32) * (np.float32(1.0) / np.float32(0xFFFFFFFF))
    assert np.array_equal(got[:227], std[:227])
    assert got[227] != std[227]
    assert (got >= 0).all() and (got <= 1).all()


def _testbvh_points():
    # Test/TestBvh.cpp:120-126,143-150: 10,000 points, x,y,z drawn in order from Random rng(0xFFFF)
    return orc.thea_random_uniform01(0xFFFF, 30000).reshape(10000, 3)


def test_testbvh_point_recipe():
    pts = _testbvh_points()
    ora = orc.Oracle(points=pts)
    assert orc.lib().orc_max_depth(ora.h) == 11            # ceil(log2(10000 / 5)), Math.cpp:22-36
    q = np.array([[0.5, 0.5, 0.5]], np.float32)
    # closestElement with dist_bound 0.25 (TestBvh.cpp:204-208) and kClosestPairs k = 3 (:221-223)
    for mode in (orc.MODE_REFERENCE, orc.MODE_CANON):
        i, d2, cp = ora.closest(q, dist_bound=0.25, mode=mode)
        bi, bd2, bcp = ora.closest(q, dist_bound=0.25, mode=orc.MODE_BRUTE)
        assert i[0] == bi[0] >= 0 and d2[0] == bd2[0] and np.array_equal(cp, bcp)
        ki, kd = ora.knn(q, 3, mode=mode)
        bki, bkd = ora.knn(q, 3, mode=orc.MODE_BRUTE)
        assert np.array_equal(ki, bki) and np.array_equal(kd, bkd)
    true = ((pts.astype(np.float64) - 0.5) ** 2).sum(axis=1)
    assert i[0] == int(np.argmin(true))


def test_testbvh_triangle_recipe():
    # Test/TestBvh.cpp:315-325: 1,000 random triangles in [0,1]^3, 9 draws each, generator continues after the points
    u = orc.thea_random_uniform01(0xFFFF, 30000 + 9000)[30000:]
    tris = u.reshape(1000, 9)
    ora = orc.Oracle(tris9=tris)
    q = np.array([[0.5, 0.5, 0.5]], np.float32)
    ri, rd, rc = ora.closest(q, mode=orc.MODE_REFERENCE)
    bi, bd, bc = ora.closest(q, mode=orc.MODE_BRUTE)
    assert ri[0] == bi[0] and rd[0] == bd[0] and np.array_equal(rc, bc)
    o = np.zeros((1, 3), np.float32)
    d = orc.thea_random_uniform01(0x1234, 3).reshape(1, 3)
    r = ora.ray_first(o, d, mode=orc.MODE_REFERENCE)
    b = ora.ray_first(o, d, mode=orc.MODE_BRUTE)
    assert r[0][0] == b[0][0] and r[1][0] == b[1][0]


def test_modes_agree_on_mesh():
    mesh = gen.torus_mesh(40)
    ora = orc.Oracle(tris9=mesh)
    b = gen.mesh_bounds(mesh)
    q = gen.closest_queries(6000, b)
    bi, bd, bc = ora.closest(q, mode=orc.MODE_BRUTE)
    ci, cd, cc = ora.closest(q, mode=orc.MODE_CANON)
    ri, rd, rc = ora.closest(q, mode=orc.MODE_REFERENCE)
    assert np.array_equal(bi, ci) and np.array_equal(bd, cd) and np.array_equal(bc, cc)
    # the reference's first-found rule may pick another element among exact ties, never another distance
    assert np.array_equal(rd, bd)
    ties = ri != bi
    assert ties.mean() < 0.6
    assert (ri[ties] > bi[ties]).all()                  # canonical = lowest index of the tie set
    o, d = gen.random_rays(6000, b)
    rb = ora.ray_first(o, d, mode=orc.MODE_BRUTE)
    rc_ = ora.ray_first(o, d, mode=orc.MODE_CANON)
    rr = ora.ray_first(o, d, mode=orc.MODE_REFERENCE)
    assert np.array_equal(rb[0], rc_[0]) and np.array_equal(rb[1], rc_[1])
    assert np.array_equal(rb[1], rr[1])
    assert ((rb[0] >= 0) == (rr[0] >= 0)).all()


def test_point_structure_modes_agree():
    mesh = gen.torus_mesh(32)
    p, n, tri = gen.area_samples(mesh, 3000)
    op, on, ot = orc.Oracle(tris9=mesh).area_sample(3000, gen.SEED_SAMPLES)
    assert np.array_equal(p, op) and np.array_equal(n, on) and np.array_equal(tri, ot)
    ora = orc.Oracle(points=p, normals=n)
    for k in (1, 16):
        b = ora.knn(p, k, exclude_self=True, mode=orc.MODE_BRUTE)
        c = ora.knn(p, k, exclude_self=True, mode=orc.MODE_CANON)
        r = ora.knn(p, k, exclude_self=True, mode=orc.MODE_REFERENCE)
        assert np.array_equal(b[0], c[0]) and np.array_equal(b[1], c[1])
        assert np.array_equal(b[1], r[1])
    seeds = (gen.splitmix64(7, 100) % np.uint64(len(p))).astype(np.int64)
    avg = ora.avg_separation(seeds)
    assert 0 < avg < 0.2
    g0 = ora.ball_graph(16, 4 * avg, mode=orc.MODE_REFERENCE)
    g2 = ora.ball_graph(16, 4 * avg, mode=orc.MODE_BRUTE)
    assert all(np.array_equal(a, b) for a, b in zip(g0, g2))
    assert (g0[2] >= 4).mean() > 0.99


def test_polar_rotation_matches_numpy_svd():
    rng = np.random.default_rng(5)
    for _ in range(50):
        M = rng.standard_normal((3, 3))
        U, _, Vt = np.linalg.svd(M)
        assert np.allclose(orc.polar_rotation(M), U @ Vt, atol=1e-10)


def test_icp_one_step_is_the_reference_formula():
    # Icp3.hpp:298-384: centroids, cov = sum dp dq^T, rot = U V^T (sic -- for cov built this way the least-squares
    # rotation is V U^T; the reference applies its transpose), trans = q_mean - rot p_mean.
    mesh = gen.torus_mesh(48)
    ora = orc.Oracle(tris9=mesh)
    p, _, _ = gen.area_samples(mesh, 2000)
    R, t = gen.rigid_pose(degrees=3.0, t=(0.01, -0.005, 0.008))
    src = gen.apply_pose(p, R, t)
    tr, err, idx = ora.icp(src, min_iters=1, max_iters=1, mode=orc.MODE_CANON)
    assert idx.shape[0] == 2
    _, _, to = ora.closest(src, mode=orc.MODE_CANON)
    P, Q = src.astype(np.float64), to.astype(np.float64)
    pm, qm = P.mean(0), Q.mean(0)
    cov = (P - pm).T @ (Q - qm)
    U, _, Vt = np.linalg.svd(cov)
    rot = U @ Vt
    assert np.allclose(tr[:, :3], rot, atol=2e-5)
    assert np.allclose(tr[:, 3], qm - rot @ pm, atol=2e-5)
    # and that formula rotates the wrong way: the least-squares rotation is its transpose
    kabsch = Vt.T @ U.T
    res_ref = np.linalg.norm((P - pm) @ rot.T - (Q - qm))
    res_ok = np.linalg.norm((P - pm) @ kabsch.T - (Q - qm))
    assert res_ok < res_ref


def test_icp_translation_only_converges_for_a_few_steps():
    mesh = gen.torus_mesh(48)
    ora = orc.Oracle(tris9=mesh)
    p, _, _ = gen.area_samples(mesh, 2000)
    src = (p + np.array([0.01, -0.005, 0.008], np.float32)).astype(np.float32)
    errs = [ora.icp(src, min_iters=k, max_iters=k, mode=orc.MODE_CANON)[1] for k in (1, 2)]
    assert errs[1] < errs[0]


def test_line_segment_known_answers():
    """Hand-derivable answers of the segment metrics (LineSegmentN.hpp:106-145,262-357; Triangle3.hpp:495-509,559-577)."""
    f = np.float32
    # point against the segment (0,0,0)-(2,0,0): interior projection, beyond either end, zero-length segment
    a = np.array([[0, 0, 0]] * 3 + [[1, 1, 1]], f); b = np.array([[2, 0, 0]] * 3 + [[1, 1, 1]], f)
    p = np.array([[1, 3, 0], [-1, 0, 0], [5, 4, 0], [1, 1, 3]], f)
    d2, sp = orc.seg_point(a, b, p)
    assert d2.tolist() == [9.0, 1.0, 25.0, 4.0]
    assert sp.tolist() == [[1, 0, 0], [0, 0, 0], [2, 0, 0], [1, 1, 1]]
    # triangle (0,0,0) (4,0,0) (0,4,0): a piercing segment (distance 0 at the hit point), a parallel segment one unit above an edge
    # region, a segment whose nearest feature is a vertex, and one that stops short of the plane above the interior
    tri = np.array([[0, 0, 0, 4, 0, 0, 0, 4, 0]] * 4, f)
    a = np.array([[1, 1, 1], [1, -2, 1], [6, -1, 0], [1, 1, 3]], f)
    b = np.array([[1, 1, -1], [3, -2, 1], [8, -3, 0], [1, 1, 2]], f)
    d2, tp, sp = orc.tri_seg(tri, a, b)
    assert d2[0] == 0.0 and tp[0].tolist() == [1, 1, 0] and sp[0].tolist() == [1, 1, 0]
    assert d2[1] == 5.0 and tp[1][1] == 0.0 and sp[1][1] == -2.0 and sp[1][2] == 1.0
    assert d2[2] == 5.0 and tp[2].tolist() == [4, 0, 0] and sp[2].tolist() == [6, -1, 0]
    # the reference measures the perimeter only when the ray test misses: a segment ending above the interior gets its distance to
    # the nearest EDGE, not to the face (Triangle3.hpp:495-509 -- reproduced, not fixed)
    assert d2[3] == 5.0 and sp[3].tolist() == [1, 1, 2]
