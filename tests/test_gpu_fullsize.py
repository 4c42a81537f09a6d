"""GPU parity at the FULL sizes of BASELINE.json's configs 2-5, every result compared with the CPU oracle
(canonical kd-tree mode, itself pinned to the reference's code by tests/test_ref_pin.py) -- bit equality of indices, hit flags,
squared distances, closest points, hit times and SDF values; the one tolerance is the final ICP transform when the engine
accumulates in fixed point (1e-5, the tolerance north_star states).  Exact-tie counts (queries where the reference's own
first-found answer differs from the canonical lowest-index answer) are printed per config.

The oracle runs on all host cores; the whole file takes a few minutes on the GPU box.
"""
import os
import time

import numpy as np
import pytest

from oracle import orc
from thea_b200 import gen

pytestmark = pytest.mark.gpu


def _note(msg):
    print(f"[fullsize] {msg}", flush=True)


@pytest.fixture(scope="module")
def torus1m(built):
    from thea_b200 import Engine
    mesh = gen.torus_mesh(gen.TORUS_1M)
    assert len(mesh) == 1_002_528
    e = Engine(0)
    e.build_triangles(mesh)
    return e, mesh, orc.Oracle(tris9=mesh), gen.mesh_bounds(mesh)


def test_cfg2_knn_and_ball_graph_1m_samples(torus1m):
    """cfg2: 1,000,000 area samples of the 1M-triangle torus; (i) k = 16 nearest neighbours of every sample (exclude self), (ii) the
    PointSet3 proximity graph (max_degree 16, radius sqrt(16) * avg_sep, min_dot -0.5, three rounds) -- all 1M rows checked."""
    from thea_b200 import Engine, callers
    e, mesh, _, _ = torus1m
    p, n, _ = e.area_sample(1_000_000, gen.SEED_SAMPLES)
    pe = Engine(0)
    pe.build_points(p, n)
    ora = orc.Oracle(points=p, normals=n)
    t0 = time.time()
    ki, kd = pe.knn(p, 16, exclude_self=True)
    t1 = time.time()
    oi, od = ora.knn(p, 16, exclude_self=True, mode=orc.MODE_CANON)
    t2 = time.time()
    assert np.array_equal(ki, oi.astype(np.int32)), f"{int((ki != oi).any(axis=1).sum())} k-NN rows differ"
    assert np.array_equal(kd, od)
    ri, _ = ora.knn(p[:200_000], 16, exclude_self=True, mode=orc.MODE_REFERENCE)
    _note(f"cfg2 k-NN: 1M x 16 equal; engine {t1 - t0:.2f}s incl. copies, oracle {t2 - t1:.1f}s; reference-order rows that differ from canonical "
          f"(exact ties) in the first 200k: {int((ri != oi[:200_000]).any(axis=1).sum())}")
    seeds = (gen.splitmix64(gen.SEED_SAMPLES, 100) % np.uint64(len(p))).astype(np.int64)
    idx, sep, cnt, avg = callers.sample_graph(pe, 16, seeds)
    oavg = ora.avg_separation(seeds)
    assert avg == oavg
    radius = float(np.float32(np.sqrt(np.float32(16))) * np.float32(oavg))
    gi, gs, gc = ora.ball_graph(16, radius, mode=orc.MODE_CANON)
    assert np.array_equal(cnt, gc) and np.array_equal(idx, gi.astype(np.int32)) and np.array_equal(sep, gs)
    _note(f"cfg2 ball graph: 1M x 16 equal, mean degree {cnt.mean():.2f}")


def test_cfg3_icp_100k_x_51_batches(torus1m):
    """cfg3: 100,000-point source against the 1M-triangle target, min = max = 50 iterations -> 51 correspondence batches.
    (a) decoupled: the oracle's own iterate of every batch is fed to tgq_closest and all 100k indices, distances and closest points
        must be equal (exactness on identical inputs, all 51 batches);
    (b) end to end (default: sums in the reference's order): iterates are the reference's bit for bit, so the engine's own
        per-iteration correspondences equal the oracle's for all 51 batches and the final transform is identical;
    (c) end to end with the opt-out fixed-point sums: the same optimum, not the same bits."""
    from thea_b200 import engine as eng_mod
    e, mesh, ora, _ = torus1m
    p, _, _ = e.area_sample(100_000, gen.SEED_ICP)
    src = gen.apply_pose(p, *gen.rigid_pose())
    t0 = time.time()
    otr, oerr, oidx, opos = ora.icp_iterates(src, min_iters=50, max_iters=50, mode=orc.MODE_CANON)
    _note(f"cfg3 oracle: {oidx.shape[0]} batches in {time.time() - t0:.1f}s")
    assert oidx.shape == (51, 100_000)
    ties = 0
    for b in range(51):
        gi, gd, gc = e.closest(opos[b])
        assert np.array_equal(gi, oidx[b].astype(np.int32)), f"batch {b}: {int((gi != oidx[b]).sum())} correspondences differ on identical inputs"
        if b in (0, 25, 50):
            ri, rd, rc = ora.closest(opos[b], mode=orc.MODE_REFERENCE)
            ties += int((ri != oidx[b]).sum())
            ci, cd, cc = ora.closest(opos[b], mode=orc.MODE_CANON)
            assert np.array_equal(gd, cd) and np.array_equal(gc, cc)
    _note(f"cfg3 decoupled: 51 x 100k indices equal; exact ties (reference-order != canonical) in batches 0/25/50: {ties}")
    t0 = time.time()
    tr, err, corr = e.icp_align(src, min_iterations=50, max_iterations=50)          # default: TGQ_OPT_ICP_REFERENCE_ORDER = 1
    dt = time.time() - t0
    assert corr.shape == (51, 100_000)
    bad = [int((corr[b] != oidx[b]).sum()) for b in range(51)]
    assert sum(bad) == 0, f"per-iteration correspondences differ end to end: {bad}"
    assert np.array_equal(tr, otr) and err == oerr
    _note(f"cfg3 end to end, reference-order sums: 51 x 100k correspondences and the final transform bit-equal ({dt:.3f}s)")
    e.set_option(eng_mod.OPT_ICP_REFERENCE_ORDER, 0)
    try:
        t0 = time.time()
        tr2, err2, corr2 = e.icp_align(src, min_iterations=50, max_iterations=50)
        dt = time.time() - t0
    finally:
        e.set_option(eng_mod.OPT_ICP_REFERENCE_ORDER, 1)
    # exact (fixed-point) sums do not reproduce the ~1e-4 summation noise of the reference's 100k-term fp32 chains: same optimum,
    # not the same bits -- which is why the reference order is the default
    assert np.allclose(tr2, otr, rtol=2e-3, atol=2e-3)
    same = np.mean([np.mean(corr2[b] == oidx[b]) for b in range(51)])
    _note(f"cfg3 end to end, fixed-point sums (opt-out): max |transform difference| {np.abs(tr2 - otr).max():.2e}, "
          f"{100 * same:.3f}% of correspondences equal ({dt:.3f}s)")


def test_cfg4_shape_diameter_100k_x_30_rays(torus1m):
    """cfg4: 100,000 samples x 30 cone rays (3,000,000 first-hit rays) on the 1M-triangle torus: every ray's (t, idx) and every
    sample's SDF value."""
    e, mesh, ora, (lo, hi) = torus1m
    p, n, _ = e.area_sample(100_000, gen.SEED_SAMPLES)
    scale = float(np.linalg.norm((hi - lo).astype(np.float32)))
    sdf, rt, ri = e.sdf(p, n, scale, True)
    t0 = time.time()
    osdf, ro, rd, ohi, oht = ora.sdf(p, n, scale, True, mode=orc.MODE_CANON)
    assert np.array_equal(ri.reshape(-1), ohi.astype(np.int32)), f"{int((ri.reshape(-1) != ohi).sum())} ray indices differ"
    assert np.array_equal(rt.reshape(-1), oht) and np.array_equal(sdf, osdf)
    rsdf, _, _, rhi, _ = ora.sdf(p[:20000], n[:20000], scale, True, mode=orc.MODE_REFERENCE)
    _note(f"cfg4: 3M rays and 100k SDF values equal (oracle {time.time() - t0:.1f}s); reference-order rays that differ from canonical in the "
          f"first 600k: {int((rhi != ohi[:600000]).sum())}; SDF values that differ: {int((rsdf != osdf[:20000]).sum())}")


def test_cfg5_stride_and_prefix(built):
    """cfg5: 10M-triangle torus; of the 50M closest-point queries and the 50M rays of the bench streams, every 100th (500k each) and
    the first 1M are compared with the oracle.  The engine answers the WHOLE 50M batches (the sample is cut out of
    the full-batch result, so the Morton-ordered launch, chunking and seeding are the bench's), generated block by block."""
    torch = pytest.importorskip("torch")
    from thea_b200 import Engine
    import bench
    mesh = gen.torus_mesh(gen.TORUS_10M)
    assert len(mesh) == 9_999_392
    bounds = gen.mesh_bounds(mesh)
    dev = torch.device("cuda", 0)
    e = Engine(0)
    dm = torch.from_numpy(mesh).to(dev)
    e.build_triangles(dm)
    del dm
    N = 50_000_000
    q = torch.empty((N, 3), dtype=torch.float32, device=dev)
    for b, n in bench.blocks(N):
        q[b:b + n] = bench.torch_closest_queries(n, bounds, gen.SEED_CLOSEST, b, dev)
    idx, d2, cp = e.closest(q)
    sel = np.unique(np.concatenate([np.arange(0, N, 100), np.arange(0, 1_000_000)]))
    sel_t = torch.from_numpy(sel).to(dev)
    qs = q[sel_t].cpu().numpy()
    gi, gd, gc = idx[sel_t].cpu().numpy(), d2[sel_t].cpu().numpy(), cp[sel_t].cpu().numpy()
    # the device-generated stream is the same stream the host generator makes
    assert np.array_equal(sel[:1000], np.arange(1000)) and np.array_equal(qs[:1000], gen.closest_queries(1000, bounds))
    del q, idx, d2, cp
    ora = orc.Oracle(tris9=mesh)
    t0 = time.time()
    oi, od, oc = ora.closest(qs, mode=orc.MODE_CANON)
    t1 = time.time()
    assert np.array_equal(gi, oi.astype(np.int32)), f"{int((gi != oi).sum())} of {len(sel)} closest-point indices differ"
    assert np.array_equal(gd, od) and np.array_equal(gc, oc)
    ri, _, _ = ora.closest(qs[:200_000], mode=orc.MODE_REFERENCE)
    _note(f"cfg5 closest: {len(sel)} sampled queries of the 50M batch equal (oracle {t1 - t0:.1f}s); exact ties in 200k: {int((ri != oi[:200_000]).sum())}")
    ro = torch.empty((N, 3), dtype=torch.float32, device=dev)
    rd = torch.empty((N, 3), dtype=torch.float32, device=dev)
    for b, n in bench.blocks(N):
        o, d = bench.torch_random_rays(n, bounds, gen.SEED_RAYS, b, dev)
        ro[b:b + n] = o
        rd[b:b + n] = d
    t, ridx, _ = e.ray_first(ro, rd)
    os_, ds_ = ro[sel_t].cpu().numpy(), rd[sel_t].cpu().numpy()
    gt, gri = t[sel_t].cpu().numpy(), ridx[sel_t].cpu().numpy()
    del ro, rd, t, ridx
    t0 = time.time()
    ori, ort = ora.ray_first(os_, ds_, mode=orc.MODE_CANON)
    assert np.array_equal(gri, ori.astype(np.int32)), f"{int((gri != ori).sum())} of {len(sel)} ray indices differ"
    assert np.array_equal(gt, ort)
    rri, _ = ora.ray_first(os_, ds_, mode=orc.MODE_REFERENCE)
    _note(f"cfg5 rays: {len(sel)} sampled rays of the 50M batch equal, hit rate {np.mean(gri >= 0):.3f} (oracle {time.time() - t0:.1f}s); "
          f"exact ties: {int((rri != ori).sum())}")
    e.close()
