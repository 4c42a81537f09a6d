"""GPU parity: the CUDA engine through its C ABI (ctypes) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): element indices and hit flags bit-exact (ties -> lowest index); distances and hit
times within 1e-5 relative -- the parity build (-fmad=false) actually reproduces them bit for bit, and that is what is
asserted here.
"""
import numpy as np
import pytest

from oracle import orc
from thea_b200 import gen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng_mesh(built):
    from thea_b200 import Engine
    mesh = gen.torus_mesh(64)          # 8,192 triangles
    e = Engine(0)
    e.build_triangles(mesh)
    return e, mesh, orc.Oracle(tris9=mesh)


def test_triangle_cache_bit_exact(eng_mesh):
    e, mesh, ora = eng_mesh
    n, d, ax, ar = e.triangle_cache()
    on, od, oax, oar = ora.tri_cache()
    assert np.array_equal(n, on) and np.array_equal(d, od) and np.array_equal(ax, oax) and np.array_equal(ar, oar)


@pytest.mark.parametrize("sort", [0, 1])
def test_closest_vs_brute_force(eng_mesh, sort):
    e, mesh, ora = eng_mesh
    e.set_option(1, sort)
    q = gen.closest_queries(30000, gen.mesh_bounds(mesh))
    idx, d2, cp = e.closest(q)
    oi, od, oc = ora.closest(q, mode=orc.MODE_BRUTE)
    e.set_option(1, 2)
    assert np.array_equal(idx, oi.astype(np.int32))
    assert np.array_equal(d2, od)
    assert np.array_equal(cp, oc)


def test_closest_dist_bound_is_strict(eng_mesh):
    e, mesh, ora = eng_mesh
    q = gen.closest_queries(5000, gen.mesh_bounds(mesh), seed=77)
    _, d2, _ = e.closest(q)
    # bound exactly at the true distance of each query must reject it (BvhN.hpp:448,462), a hair above must accept
    for i in range(0, 5000, 250):
        d = float(np.sqrt(np.float64(d2[i])))
        bi, _, _ = e.closest(q[i:i + 1], dist_bound=d)
        oi, _, _ = ora.closest(q[i:i + 1], dist_bound=d, mode=orc.MODE_BRUTE)
        assert bi[0] == oi[0]
        bi2, _, _ = e.closest(q[i:i + 1], dist_bound=d * (1 + 1e-6) + 1e-12)
        oi2, _, _ = ora.closest(q[i:i + 1], dist_bound=d * (1 + 1e-6) + 1e-12, mode=orc.MODE_BRUTE)
        assert bi2[0] == oi2[0] and bi2[0] >= 0
    idx, d2b, _ = e.closest(q, dist_bound=0.05)
    oi, od, _ = ora.closest(q, dist_bound=0.05, mode=orc.MODE_BRUTE)
    assert np.array_equal(idx, oi.astype(np.int32)) and np.array_equal(d2b, od)
    assert (idx < 0).any() and (idx >= 0).any()


def test_ray_first_vs_brute_force(eng_mesh):
    e, mesh, ora = eng_mesh
    o, d = gen.random_rays(30000, gen.mesh_bounds(mesh))
    t, idx, nrm = e.ray_first(o, d, want_normal=True)
    oi, ot = ora.ray_first(o, d, mode=orc.MODE_BRUTE)
    assert np.array_equal(idx, oi.astype(np.int32))
    assert np.array_equal(t, ot)
    on = ora.tri_cache()[0]
    hit = idx >= 0
    assert hit.any() and (~hit).any()
    assert np.array_equal(nrm[hit], on[idx[hit]])
    assert np.array_equal(e.ray_any(o, d), hit.astype(np.uint8))


def test_ray_max_time_is_strict(eng_mesh):
    e, mesh, ora = eng_mesh
    o, d = gen.random_rays(8000, gen.mesh_bounds(mesh), seed=5)
    t, idx, _ = e.ray_first(o, d)
    tm = np.where(idx >= 0, t, np.float32(1.0)).astype(np.float32)   # bound exactly at the first hit: must be rejected
    tm[::2] = -1.0
    t2, idx2, _ = e.ray_first(o, d, tmax=tm)
    oi, ot = ora.ray_first(o, d, tmax=tm, mode=orc.MODE_BRUTE)
    assert np.array_equal(idx2, oi.astype(np.int32)) and np.array_equal(t2, ot)


@pytest.mark.parametrize("set_buffer", [1, 0], ids=["candidate-buffer", "sorted-set"])
def test_points_knn_closest_graph(built, set_buffer):
    from thea_b200 import Engine, engine as _eng
    mesh = gen.torus_mesh(48)
    p, n, _ = gen.area_samples(mesh, 6000)
    e = Engine(0)
    e.set_option(_eng.OPT_SET_BUFFER, set_buffer)       # 1 (default): k_knn_buf / k_ball_graph_buf; 0: k_knn / k_ball_graph
    e.build_points(p, n)
    ora = orc.Oracle(points=p, normals=n)
    for k in (1, 3, 16, 24, 33, 64):
        ki, kd = e.knn(p, k, exclude_self=True)
        oi, od = ora.knn(p, k, exclude_self=True, mode=orc.MODE_BRUTE)
        assert np.array_equal(ki, oi.astype(np.int32)) and np.array_equal(kd, od)
    q = gen.closest_queries(4000, gen.mesh_bounds(mesh))
    ki, kd = e.knn(q, 8, dist_bound=0.2)
    oi, od = ora.knn(q, 8, dist_bound=0.2, mode=orc.MODE_BRUTE)
    assert np.array_equal(ki, oi.astype(np.int32)) and np.array_equal(kd, od)
    ci, cd, cc = e.closest(q)
    oci, ocd, occ = ora.closest(q, mode=orc.MODE_BRUTE)
    assert np.array_equal(ci, oci.astype(np.int32)) and np.array_equal(cd, ocd) and np.array_equal(cc, occ)
    seeds = (gen.splitmix64(7, 100) % np.uint64(len(p))).astype(np.int64)
    avg = e.avg_separation(seeds)
    assert avg == ora.avg_separation(seeds)
    gi, gs, gc = e.ball_graph(16, np.float32(4.0) * np.float32(avg))
    oi, os_, oc = ora.ball_graph(16, np.float32(4.0) * np.float32(avg), mode=orc.MODE_BRUTE)
    assert np.array_equal(gi, oi.astype(np.int32)) and np.array_equal(gs, os_) and np.array_equal(gc, oc)
    # a tiny radius forces the doubling rounds
    gi, gs, gc = e.ball_graph(16, np.float32(0.3) * np.float32(avg))
    oi, os_, oc = ora.ball_graph(16, np.float32(0.3) * np.float32(avg), mode=orc.MODE_BRUTE)
    assert np.array_equal(gi, oi.astype(np.int32)) and np.array_equal(gs, os_) and np.array_equal(gc, oc)


def test_edge_cases(built):
    from thea_b200 import Engine, TgqError
    e = Engine(0)
    with pytest.raises(TgqError):
        e.closest(np.zeros((1, 3), np.float32))          # nothing built
    e.build_triangles(np.zeros((0, 9), np.float32))      # empty structure: -1, never an error (BvhN.hpp:1086-1087)
    idx, d2, _ = e.closest(np.zeros((5, 3), np.float32))
    assert (idx == -1).all() and (d2 == -1).all()
    t, ridx, _ = e.ray_first(np.zeros((5, 3), np.float32), np.ones((5, 3), np.float32))
    assert (ridx == -1).all() and (t == -1).all()
    one = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    e.build_triangles(one)
    idx, d2, cp = e.closest(np.array([[0.25, 0.25, 2.0], [5, 5, 0]], np.float32))
    assert idx.tolist() == [0, 0] and d2[0] == 4.0
    assert np.allclose(cp[0], [0.25, 0.25, 0.0])
    idx, _, _ = e.closest(np.zeros((0, 3), np.float32))
    assert idx.shape == (0,)
    # ragged sizes around the leaf size and block size
    rng = np.random.default_rng(3)
    for T in (2, 3, 4, 5, 7, 129):
        tris = rng.random((T, 9)).astype(np.float32)
        e.build_triangles(tris)
        ora = orc.Oracle(tris9=tris)
        q = rng.random((257, 3)).astype(np.float32) * 2 - 0.5
        idx, d2, cp = e.closest(q)
        oi, od, oc = ora.closest(q, mode=orc.MODE_BRUTE)
        assert np.array_equal(idx, oi.astype(np.int32)) and np.array_equal(d2, od) and np.array_equal(cp, oc)
        o = rng.random((257, 3)).astype(np.float32) * 2 - 0.5
        d = rng.standard_normal((257, 3)).astype(np.float32)
        t, ridx, _ = e.ray_first(o, d)
        ri, rt = ora.ray_first(o, d, mode=orc.MODE_BRUTE)
        assert np.array_equal(ridx, ri.astype(np.int32)) and np.array_equal(t, rt)


def test_duplicate_elements_tie_to_lowest_index(built):
    from thea_b200 import Engine
    rng = np.random.default_rng(11)
    base = rng.random((50, 9)).astype(np.float32)
    tris = np.concatenate([base, base, base])[rng.permutation(150)]
    e = Engine(0)
    e.build_triangles(tris)
    ora = orc.Oracle(tris9=tris)
    q = rng.random((2000, 3)).astype(np.float32)
    idx, d2, _ = e.closest(q)
    oi, od, _ = ora.closest(q, mode=orc.MODE_BRUTE)
    assert np.array_equal(idx, oi.astype(np.int32)) and np.array_equal(d2, od)
    o = rng.random((2000, 3)).astype(np.float32)
    d = rng.standard_normal((2000, 3)).astype(np.float32)
    t, ridx, _ = e.ray_first(o, d)
    ri, rt = ora.ray_first(o, d, mode=orc.MODE_BRUTE)
    assert np.array_equal(ridx, ri.astype(np.int32)) and np.array_equal(t, rt)
    pts = np.repeat(rng.random((40, 3)).astype(np.float32), 5, axis=0)
    pe = Engine(0)
    pe.build_points(pts)
    po = orc.Oracle(points=pts)
    ki, kd = pe.knn(pts, 7, exclude_self=True)
    oki, okd = po.knn(pts, 7, exclude_self=True, mode=orc.MODE_BRUTE)
    assert np.array_equal(ki, oki.astype(np.int32)) and np.array_equal(kd, okd)


def test_degenerate_triangles_match_the_reference_quirk(built):
    """Zero-area triangles: same answers as the oracle, including queries that sit exactly on them."""
    from thea_b200 import Engine
    mesh = gen.torus_mesh(16).copy()
    rng = np.random.default_rng(9)
    pick = rng.choice(len(mesh), 60, replace=False)
    for j, t in enumerate(pick):
        v = mesh[t].reshape(3, 3)
        if j % 3 == 0:
            v[1] = v[0]
        elif j % 3 == 1:
            v[2] = v[0] + np.float32(0.5) * (v[1] - v[0])
        else:
            v[1] = v[0]; v[2] = v[0]
    e = Engine(0)
    e.build_triangles(mesh)
    ora = orc.Oracle(tris9=mesh)
    q = gen.closest_queries(20000, gen.mesh_bounds(mesh), seed=77)
    q[:60] = mesh[pick].reshape(-1, 3, 3)[:, 0]
    gi, gd, gc = e.closest(q)
    oi, od, oc = ora.closest(q, mode=orc.MODE_BRUTE)
    assert np.array_equal(gi, oi.astype(np.int32)) and np.array_equal(gd, od) and np.array_equal(gc, oc)
    o, d = gen.random_rays(20000, gen.mesh_bounds(mesh), seed=78)
    t, ri, _ = e.ray_first(o, d)
    ori, ort = ora.ray_first(o, d, mode=orc.MODE_BRUTE)
    assert np.array_equal(ri, ori.astype(np.int32)) and np.array_equal(t, ort)


def test_shared_memory_limit_survives_smaller_launches(built):
    """The dynamic shared-memory limit of a kernel is process-wide state: a large k, then a small k, then the large k again, on
    one handle and across two handles (and a second handle created in between), must all launch."""
    from thea_b200 import Engine
    mesh = gen.torus_mesh(32)
    p, n, _ = gen.area_samples(mesh, 3000)
    ora = orc.Oracle(points=p, normals=n)
    ref48 = ora.knn(p[:500], 48, mode=orc.MODE_BRUTE)
    ref8 = ora.knn(p[:500], 8, mode=orc.MODE_BRUTE)
    a = Engine(0)
    a.build_points(p, n)
    for k, want in ((48, ref48), (8, ref8), (48, ref48), (64, None), (8, ref8), (48, ref48)):
        ki, kd = a.knn(p[:500], k)
        if want is not None:
            assert np.array_equal(ki, want[0].astype(np.int32)) and np.array_equal(kd, want[1])
    b = Engine(0)                       # creating a second handle must not lower the first one's limits
    b.build_points(p, n)
    ki, kd = b.knn(p[:500], 8)
    assert np.array_equal(ki, ref8[0].astype(np.int32))
    ki, kd = a.knn(p[:500], 48)
    assert np.array_equal(ki, ref48[0].astype(np.int32)) and np.array_equal(kd, ref48[1])
    gi, _, _ = a.ball_graph(48, 0.2)
    gi2, _, _ = b.ball_graph(4, 0.2)
    gi3, _, _ = a.ball_graph(48, 0.2)
    assert np.array_equal(gi, gi3)


def test_device_pointer_arguments_and_dtype_checks(built):
    torch = pytest.importorskip("torch")
    from thea_b200 import Engine, TgqError
    mesh = gen.torus_mesh(32)
    p, n, _ = gen.area_samples(mesh, 3000)
    e = Engine(0)
    e.build_points(p, n)
    seeds = (gen.splitmix64(7, 100) % np.uint64(len(p))).astype(np.int32)
    want = e.avg_separation(seeds.astype(np.int64))
    import ctypes as C
    out = C.c_float()
    ds = torch.from_numpy(seeds).cuda()
    rc = e._L.tgq_avg_separation(e._h, C.c_void_p(ds.data_ptr()), 100, C.byref(out))      # seeds as a DEVICE pointer
    assert rc == 0 and out.value == want
    with pytest.raises(TypeError):
        e.closest(torch.zeros((4, 3), dtype=torch.float64, device="cuda"))
    # torch tensors produced on the current stream are consumed in order without an explicit set_stream
    q = torch.from_numpy(gen.closest_queries(200000, gen.mesh_bounds(mesh))).cuda()
    q2 = q * 1.0 + 0.0
    i1, d1, _ = e.closest(q2)
    i0, d0, _ = e.closest(q.cpu().numpy())
    torch.cuda.synchronize()
    assert np.array_equal(i1.cpu().numpy(), i0) and np.array_equal(d1.cpu().numpy(), d0)
    t = Engine(0)
    verts, tris = gen.torus_vertices(12), gen.torus_indices(12).copy()
    tris[5, 1] = len(verts) + 3
    with pytest.raises(TgqError):
        t.build_triangles_indexed(verts, tris)
    flat = gen.torus_mesh(12)
    t.build_triangles(flat)
    t.refit_triangles(flat.reshape(-1))          # flat arrays are accepted by refit as by build


def test_engine_against_the_reference_library_directly(built):
    """Where oracle/_ref/libthea_ref.so travelled with the snapshot: the engine against the REFERENCE'S OWN CODE (no restatement in
    between): squared distances, closest points and hit times bit-equal on every query, indices equal wherever the reference's
    first-found answer is the lowest-index one (all differences are exact ties, checked)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libthea_ref.so not present")
    from thea_b200 import Engine
    mesh = gen.torus_mesh(gen.TORUS_70K)
    b = gen.mesh_bounds(mesh)
    R = ref.Reference(tris9=mesh)
    e = Engine(0)
    e.build_triangles(mesh)
    q = gen.closest_queries(300_000, b)
    for bound in (-1.0, 0.3):
        gi, gd, gc = e.closest(q, dist_bound=bound)
        ri, rd, rc = R.closest(q, dist_bound=bound)
        assert np.array_equal(gi >= 0, ri >= 0)
        f = ri >= 0
        assert np.array_equal(gd[f], rd[f])
        same = gi == ri
        assert np.array_equal(gc[same & f], rc[same & f])
        assert np.all(gi[~same] < ri[~same])                    # ties: the engine returns the lowest index, the reference the first found
        assert (~same).mean() < 0.5
    o, d = gen.random_rays(300_000, b)
    gt, gri, _ = e.ray_first(o, d)
    rri, rt = R.ray_first(o, d)
    assert np.array_equal(gt, rt) and np.array_equal(gri >= 0, rri >= 0) and np.all(gri[gri != rri] < rri[gri != rri])
    p, n, _ = gen.area_samples(mesh, 100_000)
    pe = Engine(0)
    pe.build_points(p, n)
    RP = ref.Reference(points=p)
    ki, kd = pe.knn(q[:50_000], 16, dist_bound=0.2)
    rki, rkd = RP.knn(q[:50_000], 16, dist_bound=0.2)
    assert np.array_equal(kd, rkd)
    assert (ki != rki).any(axis=1).mean() < 0.01


def test_four_lane_wide_node_kernel_matches_the_default(built):
    """TGQ_OPT_WIDE = 1 (k_closest_wide: four lanes per query on 4-wide nodes; slower than the default, kept as an option with its
    ncu evidence in profiles/r02_b_*): identical bits to the default kernel and to the oracle."""
    from thea_b200 import Engine, engine as eng_mod
    mesh = gen.torus_mesh(gen.TORUS_70K)
    b = gen.mesh_bounds(mesh)
    q = gen.closest_queries(400_000, b)
    d = Engine(0)
    d.build_triangles(mesh)
    w = Engine(0)
    w.set_option(eng_mod.OPT_WIDE, 1)
    w.build_triangles(mesh)
    ora = orc.Oracle(tris9=mesh)
    for bound in (-1.0, 0.25):
        di, dd, dc = d.closest(q, dist_bound=bound)
        wi, wd, wc = w.closest(q, dist_bound=bound)
        assert np.array_equal(di, wi) and np.array_equal(dd, wd) and np.array_equal(dc, wc)
    oi, od, oc = ora.closest(q[:100_000], mode=orc.MODE_CANON)
    wi, wd, wc = w.closest(q[:100_000])
    assert np.array_equal(wi, oi.astype(np.int32)) and np.array_equal(wd, od) and np.array_equal(wc, oc)
    for k in (1, 2, 3, 5):
        small = mesh[:k]
        s = Engine(0)
        s.set_option(eng_mod.OPT_WIDE, 1)
        s.build_triangles(small)
        si, sd, _ = s.closest(q[:2000])
        bi, bd, _ = orc.Oracle(tris9=small).closest(q[:2000], mode=orc.MODE_BRUTE)
        assert np.array_equal(si, bi.astype(np.int32)) and np.array_equal(sd, bd)
