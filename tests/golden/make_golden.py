"""Generates the golden fixtures of tests/golden/ (run once, in the development container: python tests/golden/make_golden.py).

PARITY UNPINNED BY THE REFERENCE: sidch/Thea has no golden vectors or asserting tests on the BvhN query path
(Test/TestBvh.cpp only prints) and cannot be built here (Eigen is absent), so these vectors are produced by the CPU
oracle in brute-force mode (argmin over ALL elements of the restated per-element arithmetic; no tree involved) on
  (a) the reference's own test recipe -- Test/TestBvh.cpp:120-126,143,315-325: 10,000 points and 1,000 triangles in
      [0,1]^3 drawn from Thea's Random(0xFFFF) (restated MT19937 with the G3D skip, Random.cpp:113-125), query
      (0.5, 0.5, 0.5), dist_bound 0.25, k = 3, rays from the origin;
  (b) the bumpy torus of SURVEY.md 8(d) at grid 48 with the seeded query streams of thea_b200/gen.py.
They pin the oracle against regressions (CPU test) and the CUDA engine against the oracle (GPU test) without either test
needing /root/reference at run time.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import orc            # noqa: E402
from thea_b200 import gen         # noqa: E402


def testbvh():
    u = orc.thea_random_uniform01(0xFFFF, 30000 + 9000)
    pts = u[:30000].reshape(10000, 3).copy()
    tris = u[30000:].reshape(1000, 9).copy()
    q = np.vstack([np.array([[0.5, 0.5, 0.5]], np.float32), gen.u01(0xB2000F01, 3 * 255).reshape(255, 3)]).astype(np.float32)
    d = gen.u01(0xB2000F02, 3 * 256).reshape(256, 3).astype(np.float32)
    o = np.zeros((256, 3), np.float32)
    po, to = orc.Oracle(points=pts), orc.Oracle(tris9=tris)
    out = {"pts": pts, "tris": tris, "q": q, "ray_o": o, "ray_d": d}
    out["p_idx"], out["p_d2"], _ = po.closest(q, mode=orc.MODE_BRUTE)
    out["p_idx_b025"], out["p_d2_b025"], _ = po.closest(q, dist_bound=0.25, mode=orc.MODE_BRUTE)
    out["p_knn3_idx"], out["p_knn3_d2"] = po.knn(q, 3, mode=orc.MODE_BRUTE)
    out["t_idx"], out["t_d2"], out["t_cp"] = to.closest(q, mode=orc.MODE_BRUTE)
    out["r_idx"], out["r_t"] = to.ray_first(o, d, mode=orc.MODE_BRUTE)
    np.savez_compressed(os.path.join(HERE, "testbvh_recipe.npz"), **out)
    print("testbvh_recipe: closest idx of (0.5,0.5,0.5) on points", out["p_idx"][0], "on triangles", out["t_idx"][0],
          "ray hits", int((out["r_idx"] >= 0).sum()), "of 256")


def torus48():
    mesh = gen.torus_mesh(48)
    b = gen.mesh_bounds(mesh)
    ora = orc.Oracle(tris9=mesh)
    q = gen.closest_queries(2000, b)
    o, d = gen.random_rays(2000, b)
    out = {}
    out["c_idx"], out["c_d2"], out["c_cp"] = ora.closest(q, mode=orc.MODE_BRUTE)
    out["r_idx"], out["r_t"] = ora.ray_first(o, d, mode=orc.MODE_BRUTE)
    p, n, t = gen.area_samples(mesh, 1000)
    out["s_p"], out["s_n"], out["s_tri"] = p, n, t
    po = orc.Oracle(points=p, normals=n)
    out["k_idx"], out["k_d2"] = po.knn(p, 16, exclude_self=True, mode=orc.MODE_BRUTE)
    seeds = (gen.splitmix64(0x5EED, 100) % np.uint64(1000)).astype(np.int64)
    avg = po.avg_separation(seeds)
    out["g_seeds"], out["g_avg"] = seeds, np.float32(avg)
    out["g_idx"], out["g_sep"], out["g_cnt"] = po.ball_graph(16, float(np.float32(np.sqrt(np.float32(16))) * np.float32(avg)), mode=orc.MODE_BRUTE)
    lo, hi = b
    scale = float(np.linalg.norm((hi - lo).astype(np.float32)))
    out["sdf_scale"] = np.float32(scale)
    out["sdf"], _, _, out["sdf_ridx"], out["sdf_rt"] = ora.sdf(p[:200], n[:200], scale, True, mode=orc.MODE_BRUTE)
    R, tt = gen.rigid_pose()
    src = gen.apply_pose(p, R, tt)
    out["icp_tr"], err, corr = ora.icp(src, mode=orc.MODE_BRUTE)
    out["icp_err"], out["icp_corr0"] = np.float32(err), corr[0]
    n3, dd, ax, ar = ora.tri_cache()
    out["cache_n"], out["cache_d"], out["cache_axis"], out["cache_area"] = n3[::16], dd[::16], ax[::16], ar[::16]
    np.savez_compressed(os.path.join(HERE, "torus48.npz"), **out)
    print("torus48: ray hits", int((out["r_idx"] >= 0).sum()), "of 2000; sdf valid", int((out["sdf"] >= 0).sum()), "of 200; icp batches", len(corr))


def torus48_features():
    """SURVEY 8f rows (local surface features, graph with the surface, transform) on the same torus: torus48_features.npz."""
    mesh = gen.torus_mesh(48)
    lo, hi = gen.mesh_bounds(mesh)
    ext = (hi - lo).astype(np.float32)
    scale = float(np.sqrt(np.float32(ext[0] * ext[0]) + (np.float32(ext[1] * ext[1]) + np.float32(ext[2] * ext[2]))))
    ora = orc.Oracle(tris9=mesh)
    p, n, _ = gen.area_samples(mesh, 1000)
    po = orc.Oracle(points=p, normals=n)
    out = {"scale": np.float32(scale)}
    r = np.float32(np.float32(0.1) * np.float32(scale))
    out["radius"] = r
    st = po.ball_stats(p, r, n, mode=orc.MODE_BRUTE)
    out["bs_n"], out["bs_avg_dist"], out["bs_num_offsets"], out["bs_curvature"] = st["n"], st["avg_dist"], st["num_offsets"], st["curvature"]
    out["bs_centroid"], out["bs_cov"] = st["centroid"], st["cov"]
    out["hist16"] = po.ball_hist(p, r, 16, mode=orc.MODE_BRUTE)
    dirs = gen.u01(0xB2000F10, 3 * 24).reshape(24, 3).astype(np.float32) * np.float32(2) - np.float32(1)
    dirs = (dirs / np.sqrt((dirs.astype(np.float64) ** 2).sum(axis=1, keepdims=True))).astype(np.float32)
    out["vis_dirs"] = dirs
    out["vis_pts"] = (p[:300] + np.float32(0.02) * n[:300]).astype(np.float32)
    out["vis"], out["vis_escaped"] = ora.visibility(out["vis_pts"], dirs, scale, mode=orc.MODE_BRUTE)
    seeds = (gen.splitmix64(0x5EED, 100) % np.uint64(1000)).astype(np.int64)
    avg = po.avg_separation(seeds)
    out["gs_r0"] = np.float32(np.float32(np.sqrt(np.float32(16))) * np.float32(avg))
    out["gs_idx"], out["gs_sep"], out["gs_cnt"] = po.ball_graph_surface(ora, 16, float(out["gs_r0"]), mode=orc.MODE_BRUTE)
    tr = np.array([[0.9, 0.1, 0.0, 0.3], [-0.1, 1.1, 0.05, -0.2], [0.02, 0.0, 0.8, 0.1]], np.float32)
    out["tr"] = tr
    v = mesh.reshape(-1, 3)
    moved = np.empty_like(v)
    for c in range(3):      # AffineTransformN.hpp:160 with Eigen's row . vector order t0 + (t1 + t2), then + translation
        moved[:, c] = (tr[c, 0] * v[:, 0] + (tr[c, 1] * v[:, 1] + tr[c, 2] * v[:, 2])) + tr[c, 3]
    out["tr_mesh_sample"] = moved.reshape(-1, 9)[::64]
    q = gen.closest_queries(1000, (lo, hi))
    out["tr_idx"], out["tr_d2"], out["tr_cp"] = orc.Oracle(tris9=moved.reshape(-1, 9)).closest(q, mode=orc.MODE_BRUTE)
    np.savez_compressed(os.path.join(HERE, "torus48_features.npz"), **out)
    print("torus48_features: mean ball population", float(out["bs_n"].mean()), "mean visibility", float(out["vis"].mean()),
          "graph edges", int(out["gs_cnt"].sum()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "features":
        torus48_features()
    else:
        testbvh()
        torus48()
        torus48_features()
