"""GPU parity at BASELINE.json's full sizes: what the oracle can still check directly (cfg1: the whole 1M batch), and
size-independent properties where it cannot (cfg5-size structure): margin-width invariance, batch-order invariance,
geometric consistency of the returned closest points and hit times."""
import os

import numpy as np
import pytest

from oracle import orc
from thea_b200 import gen

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WIDE = os.path.join(ROOT, "thea_b200", "libtgq_wide.so")


def test_cfg1_full_batch_against_oracle(built):
    """cfg1: 1,000,000 closest-point queries against the 70,688-triangle torus, every one checked (canonical kd-tree oracle)."""
    from thea_b200 import Engine
    mesh = gen.torus_mesh(gen.TORUS_70K)
    assert len(mesh) == 70688
    q = gen.closest_queries(1_000_000, gen.mesh_bounds(mesh))
    e = Engine(0)
    e.build_triangles(mesh)
    idx, d2, cp = e.closest(q)
    oi, od, oc = orc.Oracle(tris9=mesh).closest(q, mode=orc.MODE_CANON)
    assert np.array_equal(idx, oi.astype(np.int32))
    assert np.array_equal(d2, od) and np.array_equal(cp, oc)


def test_cfg5_size_margin_width_invariance(built):
    """10M-triangle structure, 4M closest-point queries of the cfg5 stream: the shipped build (tight, budgeted pruning
    margins) and the test build with every absolute margin x4 must return identical bits -- a margin that pruned a true
    winner anywhere would show up as a difference."""
    torch = pytest.importorskip("torch")
    from thea_b200 import Engine
    mesh = gen.torus_mesh(gen.TORUS_10M)
    assert len(mesh) == 9999392
    b = gen.mesh_bounds(mesh)
    dm = torch.from_numpy(mesh).cuda()
    q = torch.from_numpy(gen.closest_queries(4_000_000, b)).cuda()
    res = []
    for lib in (None, WIDE):
        e = Engine(0, lib=lib)
        e.build_triangles(dm)
        idx, d2, cp = e.closest(q)
        res.append((idx.cpu().numpy(), d2.cpu().numpy(), cp.cpu().numpy()))
        if lib is None:
            # batch-order invariance: unsorted launch order, same answers
            e.set_option(1, 0)
            i2, dd2, c2 = e.closest(q)
            assert np.array_equal(res[0][0], i2.cpu().numpy()) and np.array_equal(res[0][1], dd2.cpu().numpy())
            e.set_option(1, 2)
            # rays: hit point lies on the plane of the reported triangle, misses are consistent
            o, d = gen.random_rays(2_000_000, b)
            t, ri, _ = e.ray_first(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda())
            t, ri = t.cpu().numpy(), ri.cpu().numpy()
            hit = ri >= 0
            assert np.array_equal(hit, t >= 0) and 0.2 < hit.mean() < 0.9
            tri = mesh[ri[hit]].reshape(-1, 3, 3).astype(np.float64)
            n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
            n /= np.linalg.norm(n, axis=1, keepdims=True)
            ph = o[hit].astype(np.float64) + t[hit, None].astype(np.float64) * d[hit].astype(np.float64)
            off = np.abs(((ph - tri[:, 0]) * n).sum(axis=1))
            assert off.max() < 1e-4 * max(1.0, float(np.abs(t[hit]).max()))
        e.close()
    assert np.array_equal(res[0][0], res[1][0]), f"{int((res[0][0] != res[1][0]).sum())} indices differ between margin widths"
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    # geometric consistency of a sample: the closest point lies on the reported triangle and realises the distance
    idx, d2, cp = res[0]
    qs = q.cpu().numpy()
    s = np.arange(0, len(idx), 400)
    tri = mesh[idx[s]].reshape(-1, 3, 3).astype(np.float64)
    n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    assert np.abs(((cp[s] - tri[:, 0]) * n).sum(axis=1)).max() < 1e-5
    assert np.allclose(((cp[s].astype(np.float64) - qs[s]) ** 2).sum(axis=1), d2[s], rtol=1e-5, atol=1e-10)
    # the oracle on a 1-in-2000 stride (strided sample of the cfg5 stream)
    s2 = np.arange(0, len(idx), 2000)
    oi, od, oc = orc.Oracle(tris9=mesh).closest(qs[s2], mode=orc.MODE_CANON)
    assert np.array_equal(idx[s2], oi.astype(np.int32)) and np.array_equal(d2[s2], od) and np.array_equal(cp[s2], oc)
