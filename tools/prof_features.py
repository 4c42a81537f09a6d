"""Development aid: the local-feature kernels (k_ball_stats, k_ball_hist, k_vis_rays, k_ball_graph<true>) on the cfg2-shaped case for ncu captures."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from thea_b200 import Engine, gen


def main(grid=708, P=1_000_000):
    mesh = gen.torus_mesh(grid); lo, hi = gen.mesh_bounds(mesh)
    scale = float(np.linalg.norm((hi - lo).astype(np.float32)))
    e = Engine(0); e.build_triangles(torch.from_numpy(mesh).cuda())
    p, n, _ = e.area_sample(P, gen.SEED_SAMPLES)
    pe = Engine(0); pe.build_points(p, n)
    dp, dn = torch.from_numpy(p).cuda(), torch.from_numpy(n).cuda()
    rad = torch.full((P,), 0.02 * scale, dtype=torch.float32, device="cuda")
    out = torch.empty((P, 16), dtype=torch.float64, device="cuda")
    res = {}
    for _ in range(2):
        pe.ball_stats(dp, rad, dn, out=out)
    res["ball_stats_ms"] = pe.stats()["last_query_ms"]
    for _ in range(2):
        pe.ball_histogram(p[:200_000], 0.02 * scale, 32)
    res["ball_hist_200k_ms"] = pe.stats()["last_query_ms"]
    rng = np.random.default_rng(1)
    dirs = rng.normal(size=(100, 3)); dirs = (dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32)
    for _ in range(2):
        e.visibility(p[:100_000], dirs, scale)
    res["visibility_10M_rays_ms"] = e.stats()["last_query_ms"]
    seeds = (gen.splitmix64(0x5EED, 100) % np.uint64(P)).astype(np.int64)
    avg = pe.avg_separation(seeds)
    r0 = float(np.float32(4.0) * np.float32(avg))
    for _ in range(2):
        gi, gs, gc = pe.ball_graph_surface(e, 16, r0)
    res["graph_with_surface_ms"] = pe.stats()["last_query_ms"]; res["graph_with_surface_mean_degree"] = float(gc.mean())
    pi, ps, pc = pe.ball_graph(16, r0)
    res["graph_plain_ms"] = pe.stats()["last_query_ms"]; res["graph_plain_mean_degree"] = float(pc.mean())
    res["edges_removed_by_reachability"] = int((pc.astype(np.int64) - gc).sum())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
