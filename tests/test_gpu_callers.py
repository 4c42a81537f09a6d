"""GPU parity of the callers (SURVEY.md 8a A13 and configs 2-4 at sizes the oracle finishes in seconds): ICP, shape
diameter, area sampling, proximity graph, refit, indexed build -- all through the C ABI against the CPU oracle."""
import numpy as np
import pytest

from oracle import orc
from thea_b200 import callers, gen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(built):
    from thea_b200 import Engine
    mesh = gen.torus_mesh(96)          # 18,432 triangles
    e = Engine(0)
    e.build_triangles(mesh)
    return e, mesh, orc.Oracle(tris9=mesh)


def test_area_sampling_is_the_reference_sampler(setup):
    e, mesh, ora = setup
    p, n, t = e.area_sample(20000, gen.SEED_SAMPLES)
    op, on, ot = ora.area_sample(20000, gen.SEED_SAMPLES)
    assert np.array_equal(t, ot.astype(np.int32)) and np.array_equal(p, op) and np.array_equal(n, on)
    gp, gn, gt = gen.area_samples(mesh, 20000)
    assert np.array_equal(p, gp) and np.array_equal(t, gt.astype(np.int32))


def test_icp_cfg3_shape(setup):
    """cfg3 at reduced size: per-batch correspondences bit-exact, final transform within 1e-5 (BASELINE north_star)."""
    e, mesh, ora = setup
    p, _, _ = gen.area_samples(mesh, 4000)
    R, t = gen.rigid_pose()
    src = gen.apply_pose(p, R, t)
    tr, err, corr = callers.icp_align(e, src, min_iterations=8, max_iterations=8)
    otr, oerr, ocorr = ora.icp(src, min_iters=8, max_iters=8, mode=orc.MODE_BRUTE)
    assert corr.shape == ocorr.shape == (9, 4000)
    for b in range(corr.shape[0]):
        same = corr[b] == ocorr[b].astype(np.int32)
        # the two sides accumulate the 16 sums differently (fixed point vs sequential fp32), so from batch 2 on the source
        # positions differ in the last bit; closest points that lie on an edge shared by two triangles (a few % of an
        # off-surface sample set) are exact ties that such noise flips, everything else must agree
        assert same.mean() > 0.93, f"batch {b}: {int((~same).sum())} correspondences differ"
    assert np.array_equal(corr[0], ocorr[0].astype(np.int32))          # identical inputs -> identical indices
    assert np.array_equal(corr[1], ocorr[1].astype(np.int32))          # (batch 1 queries the same, not yet moved, positions)
    assert np.allclose(tr, otr, rtol=1e-5, atol=1e-5)
    assert abs(err - oerr) <= 1e-4 * max(abs(oerr), 1e-12) + 1e-9
    # the in-library loop (tgq_icp_align) accumulates in the reference's order by default: every batch and the transform are equal
    tr3, err3, corr3 = e.icp_align(src, min_iterations=8, max_iterations=8)
    assert np.array_equal(corr3, ocorr.astype(np.int32)) and np.array_equal(tr3, otr) and err3 == oerr
    # convergence-controlled run (Icp3 defaults) stops after the same number of batches
    tr2, err2, corr2 = callers.icp_align(e, src)
    otr2, oerr2, ocorr2 = ora.icp(src, mode=orc.MODE_BRUTE)
    assert corr2.shape[0] == ocorr2.shape[0]
    assert np.allclose(tr2, otr2, rtol=1e-5, atol=1e-5)


def test_icp_sums_are_launch_shape_independent(setup):
    e, mesh, _ = setup
    p, _, _ = gen.area_samples(mesh, 10001)
    q = p[::-1].copy()
    e.set_option(6, 0)                 # TGQ_OPT_ICP_REFERENCE_ORDER off: the order-independent fixed-point sums
    try:
        _check_fixed_point_sums(e, p, q)
    finally:
        e.set_option(6, 1)
    # default (reference order): the reference's sequential fp32 loops, bit for bit
    tr = np.hstack([gen.rigid_pose()[0].astype(np.float32), gen.rigid_pose()[1].reshape(3, 1).astype(np.float32)])
    s = e.icp_sums(p, q, tr)
    assert np.float32(s[15]) == np.float32(orc.icp_error(tr, p, q))
    pm = np.zeros(3, np.float32)
    for x in p:
        pm = pm + x
    assert np.array_equal(s[0:3].astype(np.float32), pm / np.float32(len(p)))


def _check_fixed_point_sums(e, p, q):
    a = e.icp_sums(p, q)
    # the same pairs in another order give bit-identical sums (fixed-point accumulation)
    perm = np.random.default_rng(3).permutation(len(p))
    b = e.icp_sums(p[perm], q[perm])
    assert np.array_equal(a, b)
    pm = p.astype(np.float64).mean(axis=0)
    assert np.allclose(a[0:3], pm, rtol=0, atol=1e-6)


def test_shape_diameter_cfg4_shape(setup):
    e, mesh, ora = setup
    p, n, _ = gen.area_samples(mesh, 1500)
    lo, hi = gen.mesh_bounds(mesh)
    scale = float(np.linalg.norm((hi - lo).astype(np.float32)))
    sdf, rt, ri = e.sdf(p, n, scale, True)
    osdf, ro, rd, ohi, oht = ora.sdf(p, n, scale, True, mode=orc.MODE_BRUTE)
    assert np.array_equal(ri.reshape(-1), ohi.astype(np.int32))
    assert np.array_equal(rt.reshape(-1), oht)
    assert np.array_equal(sdf, osdf)
    # interior hits exist on a closed, outward-oriented torus; the retry pass only touches failed samples
    assert (sdf >= 0).mean() > 0.95
    full = callers.shape_diameter(e, p, n, scale)
    ok = sdf >= 0
    assert np.array_equal(full[ok], sdf[ok])
    if (~ok).any():
        again, _, _, _, _ = ora.sdf(p[~ok], n[~ok], scale, False, mode=orc.MODE_BRUTE)
        assert np.array_equal(full[~ok], again)


def test_sample_graph_cfg2_shape(built):
    from thea_b200 import Engine
    mesh = gen.torus_mesh(64)
    p, n, _ = gen.area_samples(mesh, 8000)
    e = Engine(0)
    e.build_points(p, n)
    ora = orc.Oracle(points=p, normals=n)
    seeds = (gen.splitmix64(0x5EED, 100) % np.uint64(len(p))).astype(np.int64)
    idx, sep, cnt, avg = callers.sample_graph(e, 16, seeds)
    oavg = ora.avg_separation(seeds)
    assert avg == oavg
    radius = float(np.float32(np.sqrt(np.float32(16))) * np.float32(oavg))
    oi, os_, oc = ora.ball_graph(16, radius, mode=orc.MODE_BRUTE)
    assert np.array_equal(cnt, oc) and np.array_equal(idx, oi.astype(np.int32)) and np.array_equal(sep, os_)


def test_refit_equals_rebuild(setup):
    from thea_b200 import Engine
    _, mesh, _ = setup
    e = Engine(0)
    e.build_triangles(mesh)
    moved = (mesh.reshape(-1, 3) * np.float32(1.1) + np.array([0.05, -0.02, 0.01], np.float32)).reshape(-1, 9).astype(np.float32)
    e.refit_triangles(moved)
    fresh = Engine(0)
    fresh.build_triangles(moved)
    ora = orc.Oracle(tris9=moved)
    b = gen.mesh_bounds(moved)
    q = gen.closest_queries(20000, b, seed=11)
    o, d = gen.random_rays(20000, b, seed=12)
    for eng in (e, fresh):
        idx, d2, cp = eng.closest(q)
        oi, od, oc = ora.closest(q, mode=orc.MODE_BRUTE)
        assert np.array_equal(idx, oi.astype(np.int32)) and np.array_equal(d2, od) and np.array_equal(cp, oc)
        t, ri, _ = eng.ray_first(o, d)
        ori, ort = ora.ray_first(o, d, mode=orc.MODE_BRUTE)
        assert np.array_equal(ri, ori.astype(np.int32)) and np.array_equal(t, ort)
    assert np.array_equal(e.triangle_cache()[0], fresh.triangle_cache()[0])


def test_indexed_build_equals_soup(built):
    from thea_b200 import Engine
    verts, tris = gen.torus_vertices(40), gen.torus_indices(40)
    soup = gen.torus_mesh(40)
    a, b = Engine(0), Engine(0)
    a.build_triangles_indexed(verts, tris)
    b.build_triangles(soup)
    q = gen.closest_queries(5000, gen.mesh_bounds(soup))
    ra, rb = a.closest(q), b.closest(q)
    assert all(np.array_equal(x, y) for x, y in zip(ra, rb))


def test_device_pointers_and_host_pointers_agree(setup):
    torch = pytest.importorskip("torch")
    e, mesh, _ = setup
    q = gen.closest_queries(300000, gen.mesh_bounds(mesh), seed=5)
    e.set_option(4, 65536)                      # several staged chunks on the host path
    hi, hd, hc = e.closest(q)
    e.set_option(4, 8388608)
    dq = torch.from_numpy(q).cuda()
    di, dd, dc = e.closest(dq)
    assert np.array_equal(hi, di.cpu().numpy()) and np.array_equal(hd, dd.cpu().numpy()) and np.array_equal(hc, dc.cpu().numpy())


def test_filtered_nearest_neighbours_register_shape(built):
    """SURVEY 8f rank 1 (Register.cpp findNeighbors / updateOffsets): label- and normal-filtered NN and 3*MAX_NBRS k-NN."""
    from thea_b200 import Engine
    mesh = gen.torus_mesh(64)
    p, n, _ = gen.area_samples(mesh, 30000)
    lab = (gen.splitmix64(5, len(p)) % np.uint64(4)).astype(np.int32)
    e = Engine(0)
    e.build_points(p, n)
    e.set_point_labels(lab)
    ora = orc.Oracle(points=p, normals=n)
    ora.set_labels(lab)
    rng = np.random.default_rng(2)
    q = (p[:8000] + rng.normal(scale=0.02, size=(8000, 3))).astype(np.float32)
    qn, ql = n[:8000], lab[:8000]
    for kw in ({"qn3": qn, "qlabel": ql}, {"qn3": qn}, {"qlabel": ql}, {"qn3": qn, "qlabel": ql, "min_dot": 0.5}):
        okw = {("qn" if k == "qn3" else k): v for k, v in kw.items()}
        gi, gd, gc = e.closest_filtered(q, **kw)
        oi, od, oc = ora.closest_filtered(q, mode=orc.MODE_CANON, **okw)
        assert np.array_equal(gi, oi.astype(np.int32)) and np.array_equal(gd, od) and np.array_equal(gc, oc)
        ki, kd = e.knn_filtered(q[:3000], 24, **{k: (v[:3000] if hasattr(v, "__len__") else v) for k, v in kw.items()})
        oki, okd = ora.knn_filtered(q[:3000], 24, mode=orc.MODE_CANON, **{k: (v[:3000] if hasattr(v, "__len__") else v) for k, v in okw.items()})
        assert np.array_equal(ki, oki.astype(np.int32)) and np.array_equal(kd, okd)
    # misuse is an error, not a silent unfiltered answer
    from thea_b200 import TgqError
    bare = Engine(0)
    bare.build_points(p)
    with pytest.raises(TgqError):
        bare.closest_filtered(q[:10], qn3=qn[:10])
    with pytest.raises(TgqError):
        bare.closest_filtered(q[:10], qlabel=ql[:10])


def test_device_resident_icp_equals_the_host_driven_loop(setup):
    """tgq_icp_align (whole loop inside the library) against callers.icp_align (same steps driven from Python) and the oracle."""
    e, mesh, ora = setup
    p, _, _ = gen.area_samples(mesh, 5000)
    R, t = gen.rigid_pose()
    src = gen.apply_pose(p, R, t)
    for kw in ({}, {"min_iterations": 6, "max_iterations": 6}):
        tr_a, err_a, corr_a = e.icp_align(src, **kw)
        tr_b, err_b, corr_b = callers.icp_align(e, src, **kw)
        assert corr_a.shape == corr_b.shape and np.array_equal(corr_a[:2], corr_b[:2])
        assert np.allclose(tr_a, tr_b, rtol=1e-6, atol=1e-6) and abs(err_a - err_b) <= 1e-5 * abs(err_b) + 1e-9
        otr, oerr, ocorr = ora.icp(src, min_iters=kw.get("min_iterations", 3), max_iters=kw.get("max_iterations", 10), mode=orc.MODE_BRUTE)
        assert corr_a.shape[0] == ocorr.shape[0] and np.array_equal(corr_a[0], ocorr[0].astype(np.int32))
        assert np.allclose(tr_a, otr, rtol=1e-5, atol=1e-5)
    # device pointers in, no correspondences out
    torch = pytest.importorskip("torch")
    tr_c, err_c, nb = e.icp_align(torch.from_numpy(src).cuda(), want_correspondences=False)
    assert nb == corr_a.shape[0] or nb > 0
    assert np.allclose(tr_c, e.icp_align(src)[0], rtol=0, atol=0)


def test_ball_range_queries(built):
    """rangeQueryIndices with a Ball3 per query (SURVEY 8f rank 3 primitive): counts and lowest indices, bit-exact membership."""
    from thea_b200 import Engine
    mesh = gen.torus_mesh(64)
    p, n, _ = gen.area_samples(mesh, 20000)
    e = Engine(0)
    e.build_points(p)
    ora = orc.Oracle(points=p)
    rng = np.random.default_rng(4)
    c = (p[:6000] + rng.normal(scale=0.01, size=(6000, 3))).astype(np.float32)
    rad = rng.uniform(0.0, 0.08, 6000).astype(np.float32)
    rad[:4] = [0.0, -1.0, 10.0, 1e-6]
    for max_out, excl in ((16, False), (64, True), (0, False)):
        gi, gc = e.ball_range(c, rad, max_out, excl)
        oi, oc = ora.ball_range(c, rad, max(max_out, 1), excl)
        assert np.array_equal(gc, oc)
        if max_out:
            assert np.array_equal(gi, oi.astype(np.int32))
    assert gc[1] == 0 and gc[2] == len(p)


def test_local_features_ball_stats(built):
    """AverageDistance / Curvature / LocalPca as one ball-statistics pass, against the oracle's restatement of
    the reference functors in the reference's traversal order (tolerances stated in tests/test_local_features.py)."""
    import torch
    from thea_b200 import Engine
    from test_local_features import check_against_oracle
    mesh = gen.torus_mesh(64)
    p, n, _ = gen.area_samples(mesh, 30000)
    e = Engine(0)
    e.build_points(p, n)
    ora = orc.Oracle(points=p, normals=n)
    rng = np.random.default_rng(5)
    Q = 5000
    sel = rng.integers(0, len(p), Q)
    c = (p[sel] + rng.normal(scale=0.005, size=(Q, 3))).astype(np.float32)
    qn = n[sel].copy()
    rad = rng.uniform(0.01, 0.15, Q).astype(np.float32)
    rad[:4] = [0.0, -1.0, 10.0, 1e-6]
    st = e.ball_stats(c, rad, qn)
    assert st[1, 0] == 0 and st[2, 0] == len(p)
    check_against_oracle(st, c, qn, rad, ora.ball_stats(c, rad, qn, mode=orc.MODE_REFERENCE))
    _, cnt = e.ball_range(c, rad, 0)
    assert np.array_equal(cnt, st[:, 0].astype(np.int32))
    # device pointers give the same bits as host pointers (one lane per ball, fixed traversal order: deterministic)
    dc, dn, dr = (torch.from_numpy(a).cuda() for a in (c, qn, rad))
    dout = torch.empty((Q, 16), dtype=torch.float64, device="cuda")
    e.ball_stats(dc, dr, dn, out=dout)
    assert np.array_equal(dout.cpu().numpy(), st)
    assert np.array_equal(e.ball_stats(c, rad, qn), st)
    # epilogues
    r = np.float32(0.1)
    ad = callers.average_distance(e, c, r)
    assert np.allclose(ad, ora.ball_stats(c, r)["avg_dist"] / np.float64(r), rtol=1e-12)
    evals, evecs = callers.local_pca(e, c, r)
    assert np.median(np.abs((evecs[:, 2, :] * qn).sum(axis=1))) > 0.95
    cu = callers.projected_curvature(e, c, qn, r)
    assert np.abs(cu - ora.ball_stats(c, r, qn)["curvature"]).max() <= 2e-5
    # LocalDistanceHistogram: integer counts, bit exact
    for nb in (5, 64):
        assert np.array_equal(e.ball_histogram(c, rad, nb).astype(np.float64), ora.ball_hist(c, rad, nb))
    # a structure without normals refuses query normals loudly
    e2 = Engine(0)
    e2.build_points(p)
    with pytest.raises(Exception):
        e2.ball_stats(c[:4], rad[:4], qn[:4])
    assert (e2.ball_stats(c[:64], rad[:64])[:, 2:6] == 0).all()


def test_visibility(setup):
    """Visibility::compute with fixed directions: escaped-ray counts equal to the oracle's (hit / miss flags are bit-exact)."""
    e, mesh, ora = setup
    rng = np.random.default_rng(8)
    D = 48
    dirs = rng.normal(size=(D, 3))
    dirs = (dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32)
    p, n, _ = gen.area_samples(mesh, 3000)
    pts = np.concatenate([(p + 0.02 * n).astype(np.float32), np.array([[1, 0, 0], [5, 5, 5]], np.float32)])
    lo, hi = gen.mesh_bounds(mesh)
    ext = (hi - lo).astype(np.float32)
    scale = float(np.sqrt(np.float32(ext[0] * ext[0]) + (np.float32(ext[1] * ext[1]) + np.float32(ext[2] * ext[2]))))
    vis, nesc = e.visibility(pts, dirs, scale)
    ovis, onesc = ora.visibility(pts, dirs, scale, mode=orc.MODE_CANON)
    assert np.array_equal(nesc, onesc) and np.array_equal(vis, ovis)
    assert vis[-2] == 0.0 and vis[-1] > 0.5
    assert np.array_equal(callers.visibility(e, pts, dirs), vis)
    # per-sample directions
    pd = rng.normal(size=(200, 7, 3))
    pd = (pd / np.linalg.norm(pd, axis=2, keepdims=True)).astype(np.float32)
    v2, n2 = e.visibility(pts[:200], pd, scale)
    o2, on2 = ora.visibility(pts[:200], pd, scale, mode=orc.MODE_CANON)
    assert np.array_equal(n2, on2) and np.array_equal(v2, o2)


def test_sample_graph_with_reachability(setup):
    """PointSet3::updateGraph(surface) on the GPU: bit-equal to the oracle's reference-order restatement."""
    from thea_b200 import Engine
    from test_local_features import thin_slab_case
    mesh, p, nrm = thin_slab_case()
    pe, se = Engine(0), Engine(0)
    pe.build_points(p, nrm)
    se.build_triangles(mesh)
    ora, oS = orc.Oracle(points=p, normals=nrm), orc.Oracle(tris9=mesh)
    gi, gs, gc = pe.ball_graph_surface(se, 8, 0.06)
    oi, osep, oc = ora.ball_graph_surface(oS, 8, 0.06)
    assert np.array_equal(gi, oi.astype(np.int32)) and np.array_equal(gs, osep) and np.array_equal(gc, oc)
    si, _, _ = pe.ball_graph_surface(se, 8, 0.06, min_dot=-2.0)
    side = (np.arange(len(p)) >= len(p) // 2)
    assert ((side[si.clip(0)] != side[:, None]) & (si >= 0)).sum() == 0
    pi, _, _ = pe.ball_graph(8, 0.06, min_dot=-2.0)
    assert ((side[pi.clip(0)] != side[:, None]) & (pi >= 0)).sum() > 100
    # the torus: samples of a closed smooth surface, graph through the caller with and without the surface
    e, tmesh, tora = setup
    sp, sn, _ = e.area_sample(6000, gen.SEED_SAMPLES)
    te = Engine(0)
    te.build_points(sp, sn)
    seeds = (gen.splitmix64(0x5EED, 100) % np.uint64(len(sp))).astype(np.int64)
    ti, ts, tc, avg = callers.sample_graph(te, 8, seeds, surface=e)
    po = orc.Oracle(points=sp, normals=sn)
    r0 = float(np.float32(np.sqrt(np.float32(8))) * np.float32(avg))
    oi, osep, oc = po.ball_graph_surface(tora, 8, r0)
    assert np.array_equal(ti, oi.astype(np.int32)) and np.array_equal(ts, osep) and np.array_equal(tc, oc)
    # errors: no normals / wrong surface kind
    nn = Engine(0)
    nn.build_points(p)
    with pytest.raises(Exception):
        nn.ball_graph_surface(se, 8, 0.06)
    with pytest.raises(Exception):
        pe.ball_graph_surface(pe, 8, 0.06)


def test_set_transform(setup):
    """BvhN::setTransform / clearTransform: queries see the transformed elements, as the reference's
    MetricL2 on Transformer::transform(element, transform) does; clearing restores the original answers bit for bit."""
    from thea_b200 import Engine
    e0, mesh, ora0 = setup
    tr = np.array([[0.9, 0.1, 0.0, 0.3], [-0.1, 1.1, 0.05, -0.2], [0.02, 0.0, 0.8, 0.1]], np.float32)   # Register aligns boxes with scale + shift
    b = gen.mesh_bounds(mesh)
    q = gen.closest_queries(20000, b)
    o, d = gen.random_rays(20000, b)
    e = Engine(0)
    e.build_triangles(mesh)
    i0, d0, c0 = e.closest(q)
    t0, r0, _ = e.ray_first(o, d)
    moved = e.transform_points(mesh.reshape(-1, 3), tr).reshape(-1, 9)
    ora = orc.Oracle(tris9=moved)
    e.set_transform(tr)
    i1, d1, c1 = e.closest(q)
    oi, od, oc = ora.closest(q, mode=orc.MODE_CANON)
    assert np.array_equal(i1, oi.astype(np.int32)) and np.array_equal(d1, od) and np.array_equal(c1, oc)
    t1, r1, _ = e.ray_first(o, d)
    ori, ort = ora.ray_first(o, d, mode=orc.MODE_CANON)
    assert np.array_equal(r1, ori.astype(np.int32)) and np.array_equal(t1, ort)
    _, _, lo, hi = e.info()
    assert np.array_equal(lo, moved.reshape(-1, 3).min(axis=0)) and np.array_equal(hi, moved.reshape(-1, 3).max(axis=0))
    e.set_transform(tr)                                  # setting it again starts from the untransformed elements, not from the moved ones
    assert np.array_equal(e.closest(q)[1], od)
    e.set_transform(None)
    i2, d2, c2 = e.closest(q)
    assert np.array_equal(i2, i0) and np.array_equal(d2, d0) and np.array_equal(c2, c0)
    assert np.array_equal(e.ray_first(o, d)[0], t0)
    # point structure with normals and labels: the filters keep working on the elements' own normals / labels
    p, n, _ = gen.area_samples(mesh, 8000)
    lab = (np.arange(len(p)) % 3).astype(np.int32)
    pe = Engine(0)
    pe.build_points(p, n)
    pe.set_point_labels(lab)
    pe.set_transform(tr)
    pm = pe.transform_points(p, tr)
    po = orc.Oracle(points=pm, normals=n)
    po.set_labels(lab)
    fi, fd, _ = pe.closest_filtered(q[:5000], n[:5000], lab[:5000])
    ofi, ofd, _ = po.closest_filtered(q[:5000], n[:5000], lab[:5000], mode=orc.MODE_CANON)
    assert np.array_equal(fi, ofi.astype(np.int32)) and np.array_equal(fd, ofd)
    ki, kd = pe.knn(q[:3000], 5)
    oki, okd = po.knn(q[:3000], 5, mode=orc.MODE_CANON)
    assert np.array_equal(ki, oki.astype(np.int32)) and np.array_equal(kd, okd)
    pe.set_transform(None)
    po0 = orc.Oracle(points=p, normals=n)
    po0.set_labels(lab)
    fi, fd, _ = pe.closest_filtered(q[:5000], n[:5000], lab[:5000])
    ofi, ofd, _ = po0.closest_filtered(q[:5000], n[:5000], lab[:5000], mode=orc.MODE_CANON)
    assert np.array_equal(fi, ofi.astype(np.int32)) and np.array_equal(fd, ofd)
    # a refit forgets the transform (new elements)
    e.set_transform(tr)
    e.refit_triangles(mesh)
    assert np.array_equal(e.closest(q)[1], d0)


def test_structure_to_structure_closest_pair(built):
    """tgq_closest_pair (BvhN::closestPair with a structure as the query, BvhN.hpp:1081-1114,1900-2009): the canonical minimum over
    all (target element, query element) pairs, against the fp32 definition; against the reference's own code where its library
    travelled."""
    from thea_b200 import Engine
    mesh = gen.torus_mesh(40)
    a, _, _ = gen.area_samples(mesh, 4000, seed=1)
    b = (gen.area_samples(mesh, 3000, seed=2)[0] + np.float32([0.011, -0.006, 0.009])).astype(np.float32)
    ea, eb, et = Engine(0), Engine(0), Engine(0)
    ea.build_points(a)
    eb.build_points(b)
    et.build_triangles(mesh)
    d = (a[:, None, :] - b[None, :, :]).astype(np.float32)
    xx, yy, zz = d[..., 0] * d[..., 0], d[..., 1] * d[..., 1], d[..., 2] * d[..., 2]
    m = xx + (yy + zz)                                   # MetricL2 point-point in Eigen's order
    for bound in (-1.0, 0.05, float(np.sqrt(np.float64(m.min()))) * 0.999):
        ti, qi, d2, qp, tp = ea.closest_pair(eb, dist_bound=bound)
        ok = m < np.float32(bound) ** 2 if bound >= 0 else np.ones_like(m, bool)
        ok = (m.astype(np.float64) < bound * bound) if bound >= 0 else ok
        if not ok.any():
            assert ti == -1 and qi == -1
            continue
        best = np.where(ok, m, np.float32(np.inf)).min()
        cand = np.argwhere(ok & (m == best))
        want = cand[np.lexsort((cand[:, 1], cand[:, 0]))][0]          # lowest target index, then lowest query index
        assert (ti, qi) == (int(want[0]), int(want[1])) and np.float32(d2) == best
        assert np.array_equal(qp, b[qi]) and np.array_equal(tp, a[ti])
    # triangles as the target: per query point the canonical closest triangle, then the minimum
    oi, od, oc = orc.Oracle(tris9=mesh).closest(b, mode=orc.MODE_BRUTE)
    ti, qi, d2, qp, tp = et.closest_pair(eb)
    best = od.min()
    cand = np.flatnonzero(od == best)
    k = cand[np.lexsort((cand, oi[cand]))][0]
    assert (ti, qi) == (int(oi[k]), int(k)) and np.float32(d2) == best and np.array_equal(tp, oc[k])
    from oracle import ref
    if ref.available():
        found, rti, rqi, rd2, ra, rb = ref.Reference(points=a).closest_pair(ref.Reference(points=b))
        ti, qi, d2, _, _ = ea.closest_pair(eb)
        assert found and np.float32(rd2) == np.float32(d2)


def test_spin_images(built):
    """tgq_spin_image vs the oracle (SpinImage.cpp:47-75): exact counts, host and device buffers, bin counts that change the block shape."""
    import torch
    from thea_b200 import Engine
    mesh = gen.torus_mesh(64)
    p, n, _ = gen.area_samples(mesh, 20_000)
    e = Engine(0)
    e.build_points(p, n)
    ora = orc.Oracle(points=p, normals=n)
    rng = np.random.default_rng(3)
    q = (p[:700] + rng.normal(scale=0.02, size=(700, 3))).astype(np.float32)
    qn = n[:700].copy(); qn[::5] *= np.float32(4.0); qn[3] = 0
    ext = np.float32(0.75) * np.float32(2.9)
    for nr, nh in ((8, 8), (15, 15), (32, 32), (1, 3)):
        img = e.spin_image(q, qn, nr, nh, ext)
        assert np.array_equal(img.astype(np.float64), ora.spin_image(q, qn, nr, nh, ext))
    dimg = torch.zeros((700, 8, 8), dtype=torch.int32, device="cuda")
    dq, dqn = torch.from_numpy(q).cuda(), torch.from_numpy(qn).cuda()
    e._ck(e._L.tgq_spin_image(e._h, dq.data_ptr(), dqn.data_ptr(), 700, 8, 8, float(ext), dimg.data_ptr()))
    torch.cuda.synchronize()
    assert np.array_equal(dimg.cpu().numpy().astype(np.float64), ora.spin_image(q, qn, 8, 8, ext))


def test_sample_graph_shortest_paths(built):
    """tgq_graph_contract / tgq_graph_geodesic / tgq_graph_fps vs the oracle's Dijkstra (ShortestPaths.hpp:166-353 restated; pinned to the
    reference's own code in tests/test_ref_pin.py).  Contraction rows, geodesic counts and histograms and the selections are exact; the
    geodesic distance sums are compared to 1e-12 relative (the engine adds the same doubles in another order)."""
    from thea_b200 import Engine
    mesh = gen.torus_mesh(64)
    p, n, _ = gen.area_samples(mesh, 12_000)
    e = Engine(0)
    e.build_points(p, n)
    ora = orc.Oracle(points=p, normals=n)
    seeds = (gen.splitmix64(7, 100) % np.uint64(len(p))).astype(np.int64)
    avg = np.float32(e.avg_separation(seeds))
    for K in (8, 16, 5):
        gi, gs, gc = e.ball_graph(K, np.float32(np.sqrt(np.float32(K))) * avg)
        # contraction: the first 3000 samples are the originals, the rest the dense oversampling
        ci, cs, cc = e.graph_contract(gi, gs, 3000)
        oi, os_, oc = orc.graph_contract(gi, gs, 3000)
        assert np.array_equal(ci, oi.astype(np.int32)) and np.array_equal(cs, os_) and np.array_equal(cc, oc)
        sd = np.arange(0, 12_000, 23)
        for md, rng_ in ((np.float32(0.25), np.float32(0.25)), (np.float32(-1.0), np.float32(2.9))):
            s_, c_, h_ = e.graph_geodesic(gi, gs, sd, md, rng_, 16)
            rs, rc, rh = orc.graph_geodesic(gi, gs, sd, md, rng_, 16)
            assert np.array_equal(c_, rc.astype(np.int32)) and np.array_equal(h_.astype(np.float64), rh)
            assert np.allclose(s_, rs, rtol=1e-12, atol=0)
        assert np.array_equal(e.graph_fps(gi, gs, 300), orc.graph_fps(gi, gs, 300).astype(np.int32))
    # tiny graphs (fewer samples than 2 * max_degree: the contraction's pick area is larger than a queue row)
    tiny = p[:9].copy()
    e3 = Engine(0); e3.build_points(tiny)
    gi, gs, gc = e3.ball_graph(16, np.float32(10.0))
    for got, want in zip(e3.graph_contract(gi, gs, 5), orc.graph_contract(gi, gs, 5)):
        assert np.array_equal(got, want.astype(got.dtype))
    assert np.array_equal(e3.graph_fps(gi, gs, 9), orc.graph_fps(gi, gs, 9).astype(np.int32))
    s_, c_, _ = e3.graph_geodesic(gi, gs, np.arange(9), np.float32(-1.0), np.float32(1.0), 0)
    rs, rc, _ = orc.graph_geodesic(gi, gs, np.arange(9), np.float32(-1.0), np.float32(1.0), 0)
    assert np.array_equal(c_, rc.astype(np.int32)) and np.allclose(s_, rs, rtol=1e-12, atol=0)
    # a graph with unreachable parts: two far clusters with short neighbour lists
    q = np.concatenate([p[:500], p[500:1000] + np.float32(10.0)])
    e2 = Engine(0); e2.build_points(q)
    gi, gs, gc = e2.ball_graph(4, np.float32(2.0) * avg, rounds=1)
    assert np.array_equal(e2.graph_fps(gi, gs, 40), orc.graph_fps(gi, gs, 40).astype(np.int32))
    ci, cs, cc = e2.graph_contract(gi, gs, 600)
    oi, os_, oc = orc.graph_contract(gi, gs, 600)
    assert np.array_equal(ci, oi.astype(np.int32)) and np.array_equal(cs, os_) and np.array_equal(cc, oc)


def test_python_callers_for_the_graph_and_spin_features(built):
    """thea_b200.callers twins of the new adaptor members: oversampled graph, geodesic features, spin images with inferred normals,
    furthest point subsampling, sampleEvenlyBySeparation -- against the oracle on the same inputs."""
    from thea_b200 import Engine, callers
    mesh = gen.torus_mesh(48)
    p, n, _ = gen.area_samples(mesh, 6000)
    e = Engine(0)
    e.build_points(p, n)
    ora = orc.Oracle(points=p, normals=n)
    seeds = (gen.splitmix64(7, 100) % np.uint64(len(p))).astype(np.int64)
    K = 6
    ci, cs, cc, avg = callers.sample_graph_oversampled(e, 1500, K, seeds)
    gi, gs, _ = ora.ball_graph(K, np.float32(np.sqrt(np.float32(K))) * np.float32(avg), mode=orc.MODE_BRUTE)
    oi, os_, oc = orc.graph_contract(gi, gs, 1500)
    assert np.array_equal(ci, oi.astype(np.int32)) and np.array_equal(cs, os_) and np.array_equal(cc, oc)
    rng = np.random.default_rng(5)
    q = (p[:400] + rng.normal(scale=0.01, size=(400, 3))).astype(np.float32)
    seed_idx = ora.closest(q, mode=orc.MODE_BRUTE)[0]
    scale, md = np.float32(2.9), np.float32(0.3)
    rs, rc, rh = orc.graph_geodesic(gi, gs, seed_idx, md, md, 10)
    ga = callers.geodesic_average_distance(e, gi.astype(np.int32), gs, q, md, scale)
    assert np.allclose(ga, np.where(rc > 0, rs / np.maximum(rc, 1), 0.0) / np.float64(md), rtol=1e-12, atol=0)
    assert np.array_equal(callers.geodesic_distance_histogram(e, gi.astype(np.int32), gs, q, 10, md, scale), rh)
    img = callers.spin_images(e, q, None, 6, 6, scale, sample_normals=n)
    assert np.array_equal(img, ora.spin_image(q, n[seed_idx], 6, 6, np.float32(0.75) * scale))
    sel = callers.furthest_point_subsample(p, 200, seeds)
    g8i, g8s, _ = orc.Oracle(points=p).ball_graph(8, np.float32(np.sqrt(np.float32(8))) * np.float32(avg), mode=orc.MODE_BRUTE)
    assert np.array_equal(sel, orc.graph_fps(g8i, g8s, 200).astype(np.int32))
    te = Engine(0)
    te.build_triangles(mesh)
    out = callers.sample_evenly_by_separation(te, 300, 99, np.arange(100, dtype=np.int64))
    assert out is not None and out[0].shape == (300, 3) and len(np.unique(out[0], axis=0)) == 300


def test_k_closest_pairs_between_point_structures(built):
    """tgq_k_closest_pairs vs the k smallest (d2, target, query) over the oracle's brute-force k-NN rows (pinned to the reference's own
    BvhN::kClosestPairs(BvhN) in tests/test_ref_pin.py)."""
    from thea_b200 import Engine
    mesh = gen.torus_mesh(30)
    a, _, _ = gen.area_samples(mesh, 3000)
    b, _, _ = gen.area_samples(mesh, 2000, seed=9)
    b = (b + np.float32(0.01)).astype(np.float32)
    ea, eb = Engine(0), Engine(0)
    ea.build_points(a); eb.build_points(b)
    ora = orc.Oracle(points=a)
    for k, bound in ((12, -1.0), (1, -1.0), (64, -1.0), (16, 0.0052)):
        ti, qi, d2 = ea.k_closest_pairs(eb, k, bound)
        ki, kd = ora.knn(b, k, dist_bound=bound, mode=orc.MODE_BRUTE)
        rows = sorted((float(kd[q, j]), int(ki[q, j]), q) for q in range(len(b)) for j in range(k) if ki[q, j] >= 0)[:k]
        assert [(int(t), int(q)) for t, q in zip(ti, qi)] == [(t, q) for _, t, q in rows]
        assert np.array_equal(d2, np.array([r[0] for r in rows], np.float32))


def test_random_walks_on_the_sample_graph(built):
    """SurfaceFeatures::Local::RandomWalks (RandomWalks.cpp:30-117) through callers.random_walks: seed lookup (tgq_closest) and the sample
    graph (tgq_ball_graph) on the device, the walks on the host over a TheaRandom; against the oracle, which is pinned to the reference's
    own RandomWalks on an injected generator state (tests/test_ref_pin.py).  Every double bit for bit."""
    from thea_b200 import Engine
    mesh = gen.torus_mesh(30)
    p, n, _ = gen.area_samples(mesh, 3000)
    pe = Engine(0)
    pe.build_points(p, n)
    seeds = (gen.splitmix64(0x5EED, 100) % np.uint64(len(p))).astype(np.int64)
    idx, sep, cnt, _ = callers.sample_graph(pe, 8, seeds)
    q = (p[::100] + np.float32(0.01)).astype(np.float32)
    got = callers.random_walks(pe, idx, p, q, 6, 200, callers.TheaRandom(5))
    want = orc.Oracle(points=p).random_walks(idx.astype(np.int64), q, 6, 200, 5)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~np.isnan(got)].view(np.uint64), want[~np.isnan(want)].view(np.uint64))
    assert np.isfinite(want).all() and np.abs(want).max() < 1.0
