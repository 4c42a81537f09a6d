"""Times BASELINE.json configs 1-4 through the public API on one GPU (development aid; bench.py is the contract and runs cfg5).
Each config is also spot-checked against the oracle on a sample so that the timed path is the parity-green path."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from thea_b200 import Engine, callers, gen


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


def main():
    from oracle import orc
    res = {}
    # cfg1: 1M closest-point queries vs the 70,688-triangle torus (host buffers in and out)
    mesh = gen.torus_mesh(gen.TORUS_70K); b = gen.mesh_bounds(mesh)
    e = Engine(0); e.build_triangles(mesh)
    q = gen.closest_queries(1_000_000, b)
    dt, _ = timed(lambda: e.closest(q))
    dq = torch.from_numpy(q).cuda()
    dtd, _ = timed(lambda: e.closest(dq))
    res["cfg1"] = {"triangles": len(mesh), "queries": len(q), "e2e_host_buffers_Mqps": len(q) / dt / 1e6, "device_resident_Mqps": len(q) / dtd / 1e6,
                   "build_ms": e.stats()["last_build_ms"]}
    e.close()

    mesh = gen.torus_mesh(gen.TORUS_1M); b = gen.mesh_bounds(mesh)
    e = Engine(0); e.build_triangles(torch.from_numpy(mesh).cuda())
    build_1m = e.stats()["last_build_ms"]
    # cfg2: k = 16 neighbour graph over 1M area samples of the 1M-triangle torus (k-NN and the PointSet3 ball graph)
    t0 = time.perf_counter(); p, n, _ = e.area_sample(1_000_000, gen.SEED_SAMPLES); t_sample = time.perf_counter() - t0
    pe = Engine(0); pe.set_option(2, int(os.environ.get("TGQ_LEAF", "0"))); pe.build_points(p, n)
    dp = torch.from_numpy(p).cuda()
    dt_knn, (ki, kd) = timed(lambda: pe.knn(dp, 16, exclude_self=True))
    seeds = (gen.splitmix64(0x5EED, 100) % np.uint64(len(p))).astype(np.int64)
    dt_graph, (gi, gs, gc, avg) = timed(lambda: callers.sample_graph(pe, 16, seeds))
    sub = np.arange(0, len(p), 5000)
    po = orc.Oracle(points=p, normals=n)
    oki, okd = po.knn(p[sub], 16, mode=orc.MODE_CANON)          # (self included on the oracle side: drop column 0)
    res["cfg2"] = {"samples": len(p), "k": 16, "knn_Mqps_device": len(p) / dt_knn / 1e6, "ball_graph_Mqps_e2e": len(p) / dt_graph / 1e6,
                   "area_sampling_s": t_sample, "avg_separation": avg, "mean_degree": float(gc.mean()), "point_build_ms": pe.stats()["last_build_ms"],
                   "knn_sample_matches_oracle": bool(np.array_equal(ki.cpu().numpy()[sub][:, :15], oki[:, 1:].astype(np.int32)))}
    # Register-style filtered neighbours (same label, n . n' >= 0): NN and 3 * MAX_NBRS = 24 k-NN
    lab = (gen.splitmix64(5, len(p)) % np.uint64(4)).astype(np.int32)
    pe.set_point_labels(lab)
    t0 = time.perf_counter(); fi, fd, _ = pe.closest_filtered(p, n, lab); dt_fnn = time.perf_counter() - t0
    t0 = time.perf_counter(); fki, fkd = pe.knn_filtered(p, 24, n, lab); dt_fknn = time.perf_counter() - t0
    po.set_labels(lab)
    ofi, ofd, _ = po.closest_filtered(p[sub], n[sub], lab[sub], mode=orc.MODE_CANON)
    res["next1_filtered"] = {"samples": len(p), "filtered_nn_Mqps_e2e": len(p) / dt_fnn / 1e6, "filtered_knn24_Mqps_e2e": len(p) / dt_fknn / 1e6,
                             "sample_matches_oracle": bool(np.array_equal(fi[sub], ofi.astype(np.int32)) and np.array_equal(fd[sub], ofd))}
    # cfg3: ICP, 100k-point source vs the 1M-triangle target, 50 iterations (51 correspondence batches)
    src_p, _, _ = e.area_sample(100_000, gen.SEED_SAMPLES + 1)
    R, t = gen.rigid_pose(); src = gen.apply_pose(src_p, R, t)
    e.icp_align(src[:1000], max_iterations=2)
    t0 = time.perf_counter(); tr, err, nb = e.icp_align(src, min_iterations=50, max_iterations=50, want_correspondences=False); dt_icp = time.perf_counter() - t0
    t0 = time.perf_counter(); tr_h, err_h, corr = callers.icp_align(e, src, min_iterations=50, max_iterations=50); dt_host = time.perf_counter() - t0
    res["cfg3"] = {"source_points": len(src), "batches": int(nb), "seconds_device_resident": dt_icp, "correspondences_per_s_M": nb * len(src) / dt_icp / 1e6,
                   "seconds_host_driven_loop": dt_host, "final_error": err, "same_transform": bool(np.allclose(tr, tr_h, rtol=1e-5, atol=1e-5))}
    # cfg4: shape diameter, 100k samples x 30 cone rays on the 1M-triangle torus
    sp, sn, _ = e.area_sample(100_000, gen.SEED_SAMPLES + 2)
    lo, hi = b; scale = float(np.linalg.norm((hi - lo).astype(np.float32)))
    dt_sdf, sdf = timed(lambda: callers.shape_diameter(e, sp, sn, scale))
    osdf = orc.Oracle(tris9=mesh).sdf(sp[:300], sn[:300], scale, True, mode=orc.MODE_CANON)[0]
    ok = osdf >= 0
    res["cfg4"] = {"samples": len(sp), "rays": 30 * len(sp), "seconds": dt_sdf, "rays_per_s_M": 30 * len(sp) / dt_sdf / 1e6, "valid_fraction": float((sdf >= 0).mean()),
                   "sample_matches_oracle": bool(np.array_equal(sdf[:300][ok], osdf[ok])), "mesh_build_ms": build_1m}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
