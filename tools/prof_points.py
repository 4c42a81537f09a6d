"""Development aid: cfg2-shaped point-structure queries (k-NN k=16, PointSet3 ball graph) for ncu captures."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from thea_b200 import Engine, gen, callers
from thea_b200 import engine as _eng

def main(grid=708, P=1_000_000):
    mesh = gen.torus_mesh(grid)
    e = Engine(0); e.build_triangles(torch.from_numpy(mesh).cuda())
    p, n, _ = e.area_sample(P, gen.SEED_SAMPLES)
    pe = Engine(0, lib=(os.environ.get("TGQ_LIB") or None)); pe.set_option(_eng.OPT_SET_BUFFER, int(os.environ.get("TGQ_SET_BUFFER", "1"))); pe.set_option(_eng.OPT_LEAF_SIZE, int(os.environ.get("TGQ_LEAF", "0"))); pe.build_points(p, n)
    dp = torch.from_numpy(p).cuda()
    for _ in range(2):
        pe.knn(dp, 16, exclude_self=True)
    s1 = pe.stats()
    seeds = (gen.splitmix64(0x5EED, 100) % np.uint64(P)).astype(np.int64)
    avg = pe.avg_separation(seeds)
    di = torch.empty((P, 16), dtype=torch.int32, device="cuda"); ds = torch.empty((P, 16), dtype=torch.float32, device="cuda"); dc = torch.empty(P, dtype=torch.int32, device="cuda")
    for _ in range(2):
        pe.ball_graph(16, float(np.float32(4.0) * np.float32(avg)), out=(di, ds, dc))
    s2 = pe.stats()
    lab = (gen.splitmix64(5, P) % np.uint64(3)).astype(np.int32)
    pe.set_point_labels(lab)
    dn = torch.from_numpy(n).cuda(); dl = torch.from_numpy(lab).cuda()
    for _ in range(2):
        pe.knn_filtered(dp, 24, qn3=dn, qlabel=dl)
    s3 = pe.stats()
    print(json.dumps({"leaf": int(os.environ.get("TGQ_LEAF", "0")), "set_buffer": int(os.environ.get("TGQ_SET_BUFFER", "1")), "filtered_knn24_ms": s3["last_query_ms"], "filtered_knn24_Mqps": P / s3["last_query_ms"] / 1e3, "P": P, "knn_ms": s1["last_query_ms"], "knn_Mqps": P / s1["last_query_ms"] / 1e3, "graph_ms": s2["last_query_ms"],
                      "graph_Mqps": P / s2["last_query_ms"] / 1e3, "mean_degree": float(dc.float().mean().item())}))

if __name__ == "__main__":
    main()
