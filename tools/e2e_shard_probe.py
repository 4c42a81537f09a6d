"""Development aid: the host-buffer (e2e) leg of ONE rank of an 8-rank cfg5 run on one GPU -- the rank's contiguous 1/8 of the closest and
ray batches through pinned host buffers, as bench.py's step_host does -- for chunk-count experiments (TGQ_DEV_MIN_CHUNK)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from thea_b200 import Engine, gen

def main(W=8, reps=5, N=50_000_000):
    dev = torch.device("cuda", 0)
    mesh = gen.torus_mesh(gen.TORUS_10M); bounds = gen.mesh_bounds(mesh)
    e = Engine(0); e.build_triangles(torch.from_numpy(mesh).to(dev))
    n = N // W
    q = bench.torch_closest_queries(n, bounds, gen.SEED_CLOSEST, 0, dev)
    o, d = bench.torch_random_rays(n, bounds, gen.SEED_RAYS, 0, dev)
    hq, ho, hd = (t.cpu().pin_memory().numpy() for t in (q, o, d))
    h_idx = torch.empty(n, dtype=torch.int32).pin_memory().numpy(); h_d2 = torch.empty(n, dtype=torch.float32).pin_memory().numpy()
    h_cp = torch.empty((n, 3), dtype=torch.float32).pin_memory().numpy()
    h_t = torch.empty(n, dtype=torch.float32).pin_memory().numpy(); h_ri = torch.empty(n, dtype=torch.int32).pin_memory().numpy()
    e.set_option(3, 1)
    def step():
        e.closest(hq, out=(h_idx, h_d2, h_cp)); e.ray_first(ho, hd, out=(h_t, h_ri, None)); e.synchronize()
    step(); step()
    t0 = time.perf_counter()
    for _ in range(reps): step()
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"min_chunk": os.environ.get("TGQ_DEV_MIN_CHUNK", "default"), "queries_per_rank": 2 * n, "ms_per_step": round(dt * 1e3, 2),
                      "Mqps_per_rank": round(2 * n / dt / 1e6, 1), "checksum": int(h_idx.astype(np.int64).sum() + h_ri.astype(np.int64).sum())}), flush=True)

if __name__ == "__main__":
    main()
