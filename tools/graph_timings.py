"""Development aid: timings of the SURVEY 8f rank 2-3 additions on cfg2-shaped data (1 M area samples of the 1 M-triangle torus):
sample-graph contraction, geodesic neighbourhoods, furthest point sampling, spin images.  Wall clock around the C-ABI call with device
buffers where the call takes them (the graph rows stay on the device)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from thea_b200 import Engine, gen


def timed(fn, reps=2):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best, out


def main(P=1_000_000):
    mesh = gen.torus_mesh(708)
    e = Engine(0); e.build_triangles(torch.from_numpy(mesh).cuda())
    p, n, _ = e.area_sample(P, gen.SEED_SAMPLES)
    lo, hi = p.min(0), p.max(0); scale = float(np.linalg.norm(hi - lo))
    pe = Engine(0); pe.build_points(p, n)
    seeds = (gen.splitmix64(0x5EED, 100) % np.uint64(P)).astype(np.int64)
    avg = pe.avg_separation(seeds)
    res = {"P": P, "scale": scale}
    K = 8
    di = torch.empty((P, K), dtype=torch.int32, device="cuda"); ds = torch.empty((P, K), dtype=torch.float32, device="cuda"); dc = torch.empty(P, dtype=torch.int32, device="cuda")
    pe.ball_graph(K, float(np.float32(np.sqrt(np.float32(K))) * np.float32(avg)), out=(di, ds, dc))
    # contraction: the first quarter of the samples are the originals (3x oversampling)
    n_orig = P // 4
    t, (ci, cs, cc) = timed(lambda: pe.graph_contract(di, ds, n_orig), 1)
    res["contract_s"] = t; res["contract_Msources_per_s"] = n_orig / t / 1e6; res["contract_mean_degree"] = float(cc.mean())
    # geodesic neighbourhoods of 100k seeds, radius 0.02 x scale
    sd = torch.arange(0, P, P // 100_000, dtype=torch.int32, device="cuda")[:100_000].contiguous()
    md = np.float32(0.02 * scale)
    t, (s_, c_, h_) = timed(lambda: pe.graph_geodesic(di, ds, sd, md, md, 32), 1)
    res["geodesic_100k_s"] = t; res["geodesic_mean_visited"] = float(c_.mean()); res["geodesic_Mvisits_per_s"] = float(c_.sum()) / t / 1e6
    # furthest point sampling: 30k area samples -> 10k (sampleEvenlyBySeparation's default 3x oversampling)
    p2, n2, _ = e.area_sample(30_000, 77)
    pf = Engine(0); pf.build_points(p2)
    s2 = (gen.splitmix64(0x5EED, 100) % np.uint64(30_000)).astype(np.int64)
    a2 = pf.avg_separation(s2)
    gi, gs, gc = pf.ball_graph(8, float(np.float32(np.sqrt(np.float32(8))) * np.float32(a2)))
    t, sel = timed(lambda: pf.graph_fps(gi, gs, 10_000), 1)
    res["fps_30k_to_10k_s"] = t; res["fps_selected"] = int(len(sel))
    # spin images: 10k queries x 1M samples, 15 x 15 bins
    Q = 10_000
    t, img = timed(lambda: pe.spin_image(p[:Q], n[:Q], 15, 15, np.float32(0.75) * np.float32(scale)), 2)
    res["spin_10k_x_1M_s"] = t; res["spin_Gpairs_per_s"] = Q * P / t / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    main()
