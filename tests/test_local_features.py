"""SURVEY.md 8f rank 3 / rank 2 on the CPU: the ball-statistics traversal (host build of the engine headers) and the caller
epilogues against the oracle's restatement of AverageDistance / Curvature / LocalPca in the reference's traversal order, and
the PointSet3 text formats.

Tolerances: the per-element terms are the reference's fp32 expressions and the set of elements in a ball is exact, so counts
are compared for equality.  The sums are order dependent: the engine sums in double, the reference in fp32 (curvature
offsets, centroid, covariance) in ITS tree's order -- its own answer moves by fp32 rounding when the tree changes (the
reference-order and element-order oracle modes differ by the same amount).  Stated tolerances: average distance 1e-12
relative (double sums of identical fp32 terms); curvature 2e-5 absolute (values in [-1, 1], n <= a few thousand); centroid
1e-5 of the ball radius + 1e-6 of |c|; covariance 2e-5 of r^2.
"""
import numpy as np
import pytest

import emu_lib
from oracle import orc
from thea_b200 import callers, gen


@pytest.fixture(scope="module")
def cloud():
    mesh = gen.torus_mesh(24)
    p, n, _ = gen.area_samples(mesh, 4000)
    rng = np.random.default_rng(11)
    Q = 600
    sel = rng.integers(0, len(p), Q)
    c = (p[sel] + rng.normal(scale=0.01, size=(Q, 3))).astype(np.float32)
    qn = n[sel].copy()
    rad = rng.uniform(0.02, 0.3, Q).astype(np.float32)
    rad[:4] = [0.0, -1.0, 10.0, 1e-6]
    return p, n, c, qn, rad, emu_lib.Emu(points=p, normals=n), orc.Oracle(points=p, normals=n)


def check_against_oracle(st, c, qn, rad, ref):
    n = st[:, 0].astype(np.int64)
    assert np.array_equal(n, ref["n"])
    assert np.array_equal(st[:, 2].astype(np.int64), ref["num_offsets"])
    avg = np.where(n > 0, st[:, 1] / np.maximum(n, 1), 0.0)
    assert np.allclose(avg, ref["avg_dist"], rtol=1e-12, atol=0)
    curv = callers.projected_curvature(None, c, qn, rad, stats=st)
    assert np.abs(curv - ref["curvature"]).max() <= 2e-5
    evals, evecs, centroid, cov = callers.pca_from_stats(st, c)
    r = np.maximum(rad.astype(np.float64), 0)
    has = n > 0
    tol_c = 1e-5 * r + 1e-6 * np.abs(c).max(axis=1)
    assert (np.abs(centroid - ref["centroid"]).max(axis=1)[has] <= tol_c[has]).all()
    cov6 = np.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], axis=1)
    assert (np.abs(cov6 - ref["cov"]).max(axis=1) <= 2e-5 * r * r + 1e-12).all()
    assert (cov6[n < 2] == 0).all() and (evals[n < 2] == 0).all()
    return evals, evecs, cov


def test_ball_stats_matches_the_reference_functors(cloud):
    p, nrm, c, qn, rad, E, ora = cloud
    st = E.ball_stats(c, rad, qn)
    assert st[1, 0] == 0 and st[2, 0] == len(p)           # negative radius: nothing; huge radius: every sample
    for mode in (orc.MODE_REFERENCE, orc.MODE_BRUTE):
        check_against_oracle(st, c, qn, rad, ora.ball_stats(c, rad, qn, mode=mode))
    # the count equals the range query's count, the moments are the plain sums (float64 cross-check on a few balls)
    _, cnt = E.ball_range(c, rad, 0)
    assert np.array_equal(cnt, st[:, 0].astype(np.int32))
    for i in (2, 10, 50, 333):
        d = (p - c[i]).astype(np.float32)
        inside = (d[:, 0] * d[:, 0] + (d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])) <= rad[i] * rad[i]
        dd = d[inside].astype(np.float64)
        assert np.allclose(st[i, 6:9], dd.sum(axis=0), rtol=1e-12, atol=1e-12)
        assert np.allclose(st[i, [9, 12, 14]], (dd * dd).sum(axis=0), rtol=1e-12)
        assert np.allclose(st[i, 10], (dd[:, 0] * dd[:, 1]).sum(), rtol=1e-10, atol=1e-14)


def test_without_query_normals_the_curvature_sums_stay_zero(cloud):
    p, nrm, c, qn, rad, E, ora = cloud
    st = E.ball_stats(c[:50], rad[:50], None)
    assert (st[:, 2:6] == 0).all() and (st[:, 15] == 0).all()
    st2 = E.ball_stats(c[:50], rad[:50], qn[:50])
    assert np.array_equal(st[:, [0, 1]], st2[:, [0, 1]]) and np.array_equal(st[:, 6:15], st2[:, 6:15])


def test_feature_epilogues(cloud):
    p, nrm, c, qn, rad, E, ora = cloud
    r = np.float32(0.15)
    st = E.ball_stats(c, r, qn)
    ref = ora.ball_stats(c, r, qn)
    ad = callers.average_distance(None, c, r, stats=st)
    assert np.allclose(ad, ref["avg_dist"] / np.float64(r), rtol=1e-12)
    assert ((ad >= 0) & (ad <= 1)).all()
    evals, evecs, cov = check_against_oracle(st, c, qn, np.full(len(c), r, np.float32), ref)
    # eigen pairs reproduce the covariance; on a smooth surface patch the smallest-variance direction is the normal
    rec = np.einsum("qk,qki,qkj->qij", evals, evecs, evecs)
    assert np.abs(rec - cov).max() <= 1e-12
    assert (np.diff(evals, axis=1) <= 1e-18).all()
    assert np.median(np.abs((evecs[:, 2, :] * qn).sum(axis=1))) > 0.95
    # the torus is curved: offsets on the outer side lean inwards, mean projected curvature is positive there
    curv = callers.projected_curvature(None, c, qn, r, stats=st)
    assert np.isfinite(curv).all() and np.abs(curv).max() <= 1.0


def test_visibility_oracle_known_answers():
    """A point just outside a closed surface sees about half of the sphere of directions; one deep inside sees none."""
    mesh = gen.torus_mesh(24)
    ora = orc.Oracle(tris9=mesh)
    rng = np.random.default_rng(2)
    dirs = rng.normal(size=(64, 3))
    dirs = (dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32)
    inside = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], np.float32)        # on the core circle of the tube
    far = np.array([[5.0, 5.0, 5.0]], np.float32)
    lo, hi = gen.mesh_bounds(mesh)
    scale = float(np.linalg.norm(hi - lo))
    for mode in (orc.MODE_REFERENCE, orc.MODE_CANON, orc.MODE_BRUTE):
        vi, ni = ora.visibility(inside, dirs, scale, mode=mode)
        assert (vi == 0).all() and (ni == 0).all()
        vf, nf = ora.visibility(far, dirs, scale, mode=mode)
        assert 0.5 < vf[0] <= 1.0 and nf[0] == round(vf[0] * 64)
    # emulated any-hit rays agree with the oracle on the same rays
    E = emu_lib.Emu(tris9=mesh)
    p, n, _ = gen.area_samples(mesh, 200)
    pts = (p + 0.01 * n).astype(np.float32)
    v0, n0 = ora.visibility(pts, dirs, scale)
    o = (pts[:, None, :] + (dirs * np.float32(np.float32(0.001) * np.float32(scale)))[None, :, :]).astype(np.float32).reshape(-1, 3)
    d = np.broadcast_to(dirs[None, :, :], (len(pts), 64, 3)).reshape(-1, 3)
    _, hit = E.ray(o, np.ascontiguousarray(d), any_hit=True)
    assert np.array_equal((hit.reshape(len(pts), 64) < 0).sum(axis=1), n0)
    assert 0.2 < v0.mean() < 0.8


def test_graph_and_samples_files_round_trip(tmp_path, cloud):
    p, nrm, c, qn, rad, E, ora = cloud
    k = 8
    idx, sep, cnt = E.ball_graph(k, 0.08)
    g = str(tmp_path / "a.graph")
    callers.save_graph(g, k, idx, sep, cnt, write_distances=True)
    lines = open(g).read().split("\n")
    assert lines[0] == str(k) and len(lines) == len(p) + 2 and lines[-1] == ""
    first = lines[1].split()
    assert int(first[0]) == cnt[0] and len(first) == 1 + 2 * cnt[0]
    md, i2, s2, c2 = callers.load_graph(g)
    assert md == k and np.array_equal(i2, idx) and np.array_equal(c2, cnt)
    live = idx >= 0
    assert np.allclose(s2[live], sep[live], rtol=1e-5)                 # 6 significant digits in the text
    # without stored separations they are recomputed as Euclidean distances (PointSet3.cpp:304-311)
    callers.save_graph(g, k, idx, count=cnt)
    md, i3, s3, c3 = callers.load_graph(g, positions=p)
    d = np.linalg.norm(p[idx[live]] - np.repeat(p, k, axis=0).reshape(len(p), k, 3)[live], axis=1)
    assert np.array_equal(np.sort(i3, axis=1), np.sort(idx, axis=1)) and np.allclose(np.sort(s3[live]), np.sort(d.astype(np.float32)), rtol=1e-6)
    s = str(tmp_path / "a.pts")
    callers.save_samples(s, p[:100], nrm[:100])
    p2, n2 = callers.load_samples(s)
    assert np.allclose(p2, p[:100], rtol=1e-5, atol=1e-7) and np.allclose(n2, nrm[:100], rtol=1e-5, atol=1e-7)
    callers.save_samples(s, p[:100])
    p3, n3 = callers.load_samples(s)
    assert n3 is None and len(p3) == 100
    with open(s, "a") as f:
        f.write("1 2 3 0 0 1\n")
    with pytest.raises(ValueError):
        callers.load_samples(s)
    with open(g, "w") as f:
        f.write("4\n1 7\n")
    with pytest.raises(ValueError):
        callers.load_graph(g, positions=p[:1])


def thin_slab_case():
    """Samples on both faces of a thin slab: Euclidean neighbours across the slab exist (0.03 apart) but the surface blocks them."""
    g = 14
    xs = np.linspace(0, 1, g + 1, dtype=np.float32)
    tris = []
    for z, flip in ((0.0, True), (0.03, False)):
        for i in range(g):
            for j in range(g):
                a = [xs[i], xs[j], z]; b = [xs[i + 1], xs[j], z]; c = [xs[i + 1], xs[j + 1], z]; d = [xs[i], xs[j + 1], z]
                tris += ([[a, c, b], [a, d, c]] if flip else [[a, b, c], [a, c, d]])
    mesh = np.asarray(tris, np.float32).reshape(-1, 9)
    rng = np.random.default_rng(3)
    n_s = 700
    xy = rng.uniform(0.02, 0.98, (2 * n_s, 2)).astype(np.float32)
    p = np.concatenate([np.c_[xy[:n_s], np.zeros(n_s, np.float32)], np.c_[xy[n_s:], np.full(n_s, 0.03, np.float32)]]).astype(np.float32)
    nrm = np.concatenate([np.tile([0, 0, -1], (n_s, 1)), np.tile([0, 0, 1], (n_s, 1))]).astype(np.float32)
    return mesh, p, nrm


def test_graph_with_reachability_test_against_the_surface():
    """PointSet3::updateGraph(surface): SURVEY 8f rank 2.  With min_dot = -2 (angle test off) the plain graph links the two faces
    of the slab; the reachability rays remove exactly those links."""
    mesh, p, nrm = thin_slab_case()
    E, ES = emu_lib.Emu(points=p, normals=nrm), emu_lib.Emu(tris9=mesh)
    ora, oS = orc.Oracle(points=p, normals=nrm), orc.Oracle(tris9=mesh)
    k, r0 = 8, 0.06
    for mode in (orc.MODE_REFERENCE, orc.MODE_BRUTE):
        oi, osep, oc = ora.ball_graph_surface(oS, k, r0, mode=mode)
        gi, gs, gc = E.ball_graph_surface(ES, k, r0)
        assert np.array_equal(gi, oi.astype(np.int32)) and np.array_equal(gs, osep) and np.array_equal(gc, oc)
    # default min_dot = -0.5 already separates opposite normals; switch it off to see the rays do the work
    pi, ps, pc = E.ball_graph(k, r0, min_dot=-2.0)
    si, ss, sc = E.ball_graph_surface(ES, k, r0, min_dot=-2.0)
    side = (np.arange(len(p)) >= len(p) // 2)
    cross_plain = (side[pi.clip(0)] != side[:, None]) & (pi >= 0)
    cross_surf = (side[si.clip(0)] != side[:, None]) & (si >= 0)
    assert cross_plain.sum() > 100 and cross_surf.sum() == 0
    # a surface that blocks nothing leaves the plain graph untouched
    far = emu_lib.Emu(tris9=(mesh + np.float32(10.0)))
    fi, fs, fc = E.ball_graph_surface(far, k, r0)
    qi, qs, qc = E.ball_graph(k, r0)
    assert np.array_equal(fi, qi) and np.array_equal(fs, qs) and np.array_equal(fc, qc)


def test_local_distance_histogram_is_bit_exact(cloud):
    """LocalDistanceHistogram (EUCLIDEAN, explicit max_distance): integer bin counts equal the oracle's in both orders."""
    p, nrm, c, qn, rad, E, ora = cloud
    for nb in (1, 7, 32):
        h = E.ball_hist(c, rad, nb)
        for mode in (orc.MODE_REFERENCE, orc.MODE_BRUTE):
            assert np.array_equal(h.astype(np.float64), ora.ball_hist(c, rad, nb, mode=mode))
        assert np.array_equal(h.sum(axis=1), E.ball_stats(c, rad)[:, 0].astype(np.uint32))
    # known answer: distances r * (j + 0.5) / 4 from the centre fall in bin j; the distance r itself goes to the last bin (clamp)
    pts = np.array([[0.125, 0, 0], [0, 0.375, 0], [0, 0, 0.625], [0.875, 0, 0], [0, 1.0, 0], [0, 0, 1.5]], np.float32)
    K = emu_lib.Emu(points=pts)
    assert np.array_equal(K.ball_hist(np.zeros((1, 3), np.float32), 1.0, 4)[0], [1, 1, 1, 2])


def test_spin_image_is_bit_exact(cloud):
    """SpinImage::compute (SpinImage.cpp:47-75): integer counts over all samples; unnormalised and degenerate query normals
    (stableNormalized returns a zero vector unchanged), queries far outside the limits (everything clamps into the border bins)."""
    p, n, c, qn, rad, E, ora = cloud
    qn = qn.copy(); qn[::5] *= np.float32(7.5); qn[3] = 0
    c = c.copy(); c[7] += np.float32(50.0)
    ext = np.float32(0.75) * np.float32(2.4)
    for nr, nh in ((8, 8), (5, 12), (1, 1)):
        img = E.spin_image(c, qn, nr, nh, ext)
        assert np.array_equal(img.astype(np.float64), ora.spin_image(c, qn, nr, nh, ext))
        assert (img.reshape(len(c), -1).sum(axis=1) == len(p)).all()


def _walk_case():
    mesh = gen.torus_mesh(30)
    p, n, _ = gen.area_samples(mesh, 3000)
    ora = orc.Oracle(points=p, normals=n)
    seeds = (gen.splitmix64(0x5EED, 100) % np.uint64(len(p))).astype(np.int64)
    K = 8
    gi, gs, _ = ora.ball_graph(K, float(np.float32(np.sqrt(np.float32(K))) * np.float32(ora.avg_separation(seeds))))
    iso = gi.copy()
    iso[17] = -1                                   # an isolated sample: walks that reach it stop and shift the stream
    q = (p[::100] + np.float32(0.01)).astype(np.float32)
    q[3] = p[17]
    return p, gi, gs, iso, q


class _SeedLookup:
    """closest() of an Engine, answered by the oracle (CPU suite)."""

    def __init__(self, pts):
        self.o = orc.Oracle(points=pts)

    def closest(self, q):
        return self.o.closest(q, mode=orc.MODE_CANON)


def _same_or_both_nan(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)].view(np.uint64), b[~np.isnan(b)].view(np.uint64))


def test_thea_random_restatement_in_the_callers():
    """callers.TheaRandom (the vectorised MT19937 with the reference's loop bounds) against the oracle's scalar restatement, which
    tests/test_oracle.py pins to the reference's Random; integer(lo, hi) against its definition."""
    ur = orc.thea_random_uniform01(4321, 5000)
    b = callers.TheaRandom(4321).bits(5000)
    assert np.array_equal(b.astype(np.float32) * (np.float32(1.0) / np.float32(0xFFFFFFFF)), ur)
    r = callers.TheaRandom(4321)
    assert [r.bits() for _ in range(700)] == b[:700].tolist()          # scalar draws cross a regeneration boundary like bulk draws
    v = callers.TheaRandom.to_integer(np.array([0, 1, 0x7FFFFFFF, 0xFFFFFFFE, 0xFFFFFFFF], np.uint32), 0, 7)
    assert v.tolist() == [0, 0, 3, 7, 7]                                # bits = 0xFFFFFFFF would give 8: clamped (Random.cpp:152-156)


def test_random_walks_caller_matches_the_oracle(tmp_path):
    """SurfaceFeatures::Local::RandomWalks (RandomWalks.cpp:30-117): the Python caller (both of its paths) and the C++ caller of
    LocalFeatures.hpp (run on the host with given seed samples) against the oracle, which tests/test_ref_pin.py pins to the reference's
    own RandomWalks on an injected Random state.  Bit equality of every double."""
    import os
    import struct
    import subprocess
    p, gi, gs, iso, q = _walk_case()
    eng = _SeedLookup(p)
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "random_walks_host_check")
    for graph in (gi, iso):
        for steps, walks, seed in ((5, 50, 1234), (6, 300, 7)):
            want = orc.Oracle(points=p).random_walks(graph, q, steps, walks, seed)
            got = callers.random_walks(eng, graph, p, q, steps, walks, callers.TheaRandom(seed))
            assert _same_or_both_nan(got, want)
            if os.path.exists(exe):
                inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
                with open(inp, "wb") as f:
                    f.write(struct.pack("<6q", len(p), graph.shape[1], len(q), steps, walks, seed))
                    f.write(np.ascontiguousarray(p, np.float32).tobytes())
                    f.write(np.ascontiguousarray(graph, np.int32).tobytes())
                    f.write(np.ascontiguousarray(q, np.float32).tobytes())
                    f.write(eng.closest(q)[0].astype(np.int32).tobytes())
                assert subprocess.run([exe, str(inp), str(out)]).returncode == 0
                cpp = np.fromfile(out, np.float64).reshape(want.shape)
                assert _same_or_both_nan(cpp, want)
    assert np.isnan(want).any()                     # the isolated sample was reached
