"""Multi-rank host logic on CPU (gloo, world_size 2): sharding, result gather and the pooled ICP sums.  The per-rank
"engine" here is the CPU oracle standing in for thea_b200.Engine (same method shapes), so what is tested is the N>1
plumbing: 1-rank and 2-rank results must be byte-identical (SURVEY.md 8e determinism requirement)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_tile_the_batch():
    from thea_b200.sharded import shard_range
    for total in (0, 1, 7, 100, 50_000_001):
        for world in (1, 2, 3, 4, 8):
            pos = 0
            for r in range(world):
                b, c = shard_range(total, r, world)
                assert b == pos and c >= 0
                pos += c
            assert pos == total
            sizes = [shard_range(total, r, world)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_space_shards_partition_the_batch_and_are_dense():
    from thea_b200 import gen
    from thea_b200.sharded import space_shard_indices
    import torch
    mesh = gen.torus_mesh(16)
    b = gen.mesh_bounds(mesh)
    lo, hi = gen.box_scaled(*b, 1.5)
    q = gen.closest_queries(20000, b)
    q[:3] = [[100, 100, 100], [-100, 0, 0], [0, 0, 1e9]]          # outside the box: border cells
    for world in (1, 2, 3, 8):
        parts = [space_shard_indices(q, lo, hi, r, world) for r in range(world)]
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(len(q)))
        sizes = np.array([len(p) for p in parts])
        assert sizes.max() <= 1.25 * sizes.mean() + 50                # round-robin cells balance a uniform batch
        tparts = [space_shard_indices(torch.from_numpy(q), lo, hi, r, world).numpy() for r in range(world)]
        assert all(np.array_equal(a, c) for a, c in zip(parts, tparts))


def test_combine_icp_sums_matches_pooled_statistics():
    from thea_b200.sharded import combine_icp_sums
    rng = np.random.default_rng(0)
    p, q = rng.normal(size=(1000, 3)), rng.normal(size=(1000, 3))
    L = np.hstack([np.linalg.qr(rng.normal(size=(3, 3)))[0], rng.normal(size=(3, 1))])

    def sums(p, q):
        pm, qm = p.mean(0), q.mean(0)
        cov = (p - pm).T @ (q - qm)
        r = (L[:, :3] @ pm + L[:, 3] - qm)
        e = r @ r + (((p - pm) @ L[:, :3].T - (q - qm)) ** 2).sum()
        return np.concatenate([pm, qm, cov.reshape(9), [e]])

    whole = sums(p, q)
    parts = [(300, sums(p[:300], q[:300])), (0, np.zeros(16)), (700, sums(p[300:], q[300:]))]
    assert np.allclose(combine_icp_sums(parts, L), whole, rtol=1e-10, atol=1e-10)


class _OracleEngine:
    """Oracle behind the Engine method shapes (tests only)."""

    def __init__(self, mesh):
        from oracle import orc
        self.orc = orc
        self.o = orc.Oracle(tris9=mesh)

    def closest(self, q):
        i, d, c = self.o.closest(q, mode=self.orc.MODE_CANON, threads=2)
        return i.astype(np.int32), d, c

    def ray_first(self, o, d):
        i, t = self.o.ray_first(o, d, mode=self.orc.MODE_CANON, threads=2)
        return t, i.astype(np.int32), None

    def visibility(self, p, dirs, scale):
        return self.o.visibility(p, dirs, scale, mode=self.orc.MODE_CANON, threads=2)

    def closest_segment(self, a, b, dist_bound=-1.0, exclude=None):
        i, d, e, s = self.o.closest_segment(a, b, dist_bound, exclude, mode=self.orc.MODE_CANON, threads=2)
        return i.astype(np.int32), d, e, s

    def closest_triangle(self, q9, dist_bound=-1.0):
        i, d, e, q = self.o.closest_triangle(q9, dist_bound, mode=self.orc.MODE_BRUTE, threads=2)
        return i.astype(np.int32), d, e, q

    def icp_sums(self, p, q, tr=None):
        p, q = np.asarray(p, np.float64), np.asarray(q, np.float64)
        L = np.eye(3, 4) if tr is None else np.asarray(tr, np.float64).reshape(3, 4)
        pm, qm = p.mean(0), q.mean(0)
        r = L[:, :3] @ pm + L[:, 3] - qm
        e = r @ r + (((p - pm) @ L[:, :3].T - (q - qm)) ** 2).sum()
        return np.concatenate([pm, qm, ((p - pm).T @ (q - qm)).reshape(9), [e]])


class _OraclePointEngine:
    """Point oracle behind the Engine method shapes of the local-feature calls (tests only)."""

    def __init__(self, pts, nrm):
        from oracle import orc
        self.orc = orc
        self.o = orc.Oracle(points=pts, normals=nrm)

    def ball_stats(self, c, radius, qn=None):
        d = self.o.ball_stats(c, radius, qn, mode=self.orc.MODE_BRUTE, threads=2)
        return np.concatenate([d["n"][:, None].astype(np.float64), d["avg_dist"][:, None], d["num_offsets"][:, None].astype(np.float64),
                               d["curvature"][:, None], d["centroid"], d["cov"], np.zeros((len(d["n"]), 3))], axis=1)

    def ball_histogram(self, c, radius, nb):
        return self.o.ball_hist(c, radius, nb, mode=self.orc.MODE_BRUTE, threads=2)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from thea_b200 import gen
    from thea_b200.sharded import ShardedEngine, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mesh = gen.torus_mesh(24)
    eng = ShardedEngine(_OracleEngine(mesh))
    b = gen.mesh_bounds(mesh)
    q = gen.closest_queries(3001, b)              # odd size: uneven shards
    o, d = gen.random_rays(2000, b)
    idx, d2, cp = eng.closest(q)
    t, ri = eng.ray_first(o, d)
    s0, c0 = shard_range(len(q), rank, world)
    sums = eng.icp_sums(q[s0:s0 + c0], cp[s0:s0 + c0])
    # the same batch sharded by space: Morton cells of the x1.5 box dealt round-robin
    lo, hi = gen.box_scaled(*b, 1.5)
    sidx, sd2, scp = eng.closest(q, space_box=(lo, hi))
    st_, sri = eng.ray_first(o, d, space_box=(lo, hi))
    assert sidx.tobytes() == idx.tobytes() and sd2.tobytes() == d2.tobytes() and scp.tobytes() == cp.tobytes()
    assert st_.tobytes() == t.tobytes() and sri.tobytes() == ri.tobytes()
    dirs = d[:9].copy()
    vis, nesc = eng.visibility(q[:501], dirs, 3.0)
    # a point structure behind the same front end: per-ball statistics and histograms with per-query radii
    pts, nrm, _ = gen.area_samples(mesh, 1500)
    peng = ShardedEngine(_OraclePointEngine(pts, nrm))
    rad = (0.05 + 0.2 * gen.u01(77, 701)).astype(np.float32)
    st = peng.ball_stats(pts[:701], rad, nrm[:701])
    hist = peng.ball_histogram(pts[:701], np.float32(0.2), 8)
    # query objects: segments with per-query exclusions, sharded by position and by space; triangles; a mesh pair
    sa, sb = q[:1201], (q[:1201] + 0.1 * d[:1201]).astype(np.float32)
    ex = (np.arange(1201 * 2).reshape(1201, 2) % len(mesh)).astype(np.int64)
    seg = eng.closest_segment(sa, sb, 0.3, ex)
    seg_sp = eng.closest_segment(sa, sb, 0.3, ex, space_box=(lo, hi))
    assert all(x.tobytes() == y.tobytes() for x, y in zip(seg, seg_sp))
    q9 = (mesh[::7] + np.float32(0.013)).astype(np.float32)
    tq = eng.closest_triangle(q9)
    pair = eng.closest_pair_of_triangles(q9)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), idx=idx, d2=d2, cp=cp, t=t, ri=ri, sums=sums, vis=vis, nesc=nesc, st=st, hist=hist,
             seg_i=seg[0], seg_d=seg[1], seg_e=seg[2], seg_s=seg[3], tq_i=tq[0], tq_d=tq[1], tq_e=tq[2], tq_q=tq[3],
             pair=np.array([pair[0], pair[1], np.float32(pair[2]).view(np.uint32)], np.int64))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_is_byte_identical_to_one_rank(built, tmp_path):
    import torch.multiprocessing as mp
    from thea_b200 import gen
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.start_processes(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    mesh = gen.torus_mesh(24)
    one = _OracleEngine(mesh)
    b = gen.mesh_bounds(mesh)
    q = gen.closest_queries(3001, b)
    o, d = gen.random_rays(2000, b)
    idx, d2, cp = one.closest(q)
    t, ri, _ = one.ray_first(o, d)
    whole = one.icp_sums(q, cp)
    vis, nesc = one.visibility(q[:501], d[:9].copy(), 3.0)
    pts, nrm, _ = gen.area_samples(mesh, 1500)
    pone = _OraclePointEngine(pts, nrm)
    rad = (0.05 + 0.2 * gen.u01(77, 701)).astype(np.float32)
    st = pone.ball_stats(pts[:701], rad, nrm[:701])
    hist = pone.ball_histogram(pts[:701], np.float32(0.2), 8)
    sa, sb = q[:1201], (q[:1201] + 0.1 * d[:1201]).astype(np.float32)
    ex = (np.arange(1201 * 2).reshape(1201, 2) % len(mesh)).astype(np.int64)
    seg = one.closest_segment(sa, sb, 0.3, ex)
    q9 = (mesh[::7] + np.float32(0.013)).astype(np.float32)
    tq = one.closest_triangle(q9)
    k = np.lexsort((np.arange(len(tq[1])), tq[0], tq[1]))[0]
    for r in range(2):
        z = np.load(tmp_path / f"r{r}.npz")
        assert all(z[n].tobytes() == x.tobytes() for n, x in zip(("seg_i", "seg_d", "seg_e", "seg_s"), seg))
        assert all(z[n].tobytes() == x.tobytes() for n, x in zip(("tq_i", "tq_d", "tq_e", "tq_q"), tq))
        assert z["pair"].tolist() == [int(tq[0][k]), int(k), int(np.float32(tq[1][k]).view(np.uint32))]
        assert z["vis"].tobytes() == vis.tobytes() and z["nesc"].tobytes() == nesc.tobytes()
        assert z["st"].tobytes() == st.tobytes() and z["hist"].tobytes() == hist.tobytes()
        assert z["idx"].tobytes() == idx.tobytes() and z["d2"].tobytes() == d2.tobytes() and z["cp"].tobytes() == cp.tobytes()
        assert z["t"].tobytes() == t.tobytes() and z["ri"].tobytes() == ri.tobytes()
        assert np.allclose(z["sums"], whole, rtol=1e-9, atol=1e-12)
    assert np.load(tmp_path / "r0.npz")["sums"].tobytes() == np.load(tmp_path / "r1.npz")["sums"].tobytes()
