"""Development aid: one rank's share of the cfg5 batches on ONE GPU (tgq_*_records with shard_world = W, compact output), timed
per call with CUDA events -- what a rank of an N-GPU run spends in its two C-ABI calls, without the other ranks."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from thea_b200 import Engine, gen

def main(W=8, reps=5, N=50_000_000):
    dev = torch.device("cuda", 0)
    s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s)
    mesh = gen.torus_mesh(gen.TORUS_10M); bounds = gen.mesh_bounds(mesh)
    e = Engine(0); e.set_stream(s.cuda_stream)
    if os.environ.get("TGQ_BENCH_CELL_BITS"): e.set_option(8, int(os.environ["TGQ_BENCH_CELL_BITS"]))
    e.build_triangles(torch.from_numpy(mesh).to(dev))
    q = torch.empty((N, 3), dtype=torch.float32, device=dev); ro = torch.empty_like(q); rd = torch.empty_like(q)
    for b, n in bench.blocks(N):
        q[b:b + n] = bench.torch_closest_queries(n, bounds, gen.SEED_CLOSEST, b, dev)
        o, d = bench.torch_random_rays(n, bounds, gen.SEED_RAYS, b, dev); ro[b:b + n] = o; rd[b:b + n] = d
    cs = torch.empty((int(N / min(W, 8) * 1.25) + 4096) * 8 if W > 1 else N * 8, dtype=torch.float32, device=dev)
    rs = torch.empty_like(cs)
    e.set_option(3, 1)
    out = {"world": W}
    for rank in sorted(set([0, W // 2, W - 1])):
        tc, tr = [], []
        for it in range(reps + 2):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record(); nc = e.closest_records(q, [cs], rank, W, compact=W > 1); b.record(); nr = e.ray_first_records(ro, rd, [rs], rank, W, compact=W > 1); c.record()
            torch.cuda.synchronize()
            if it >= 2: tc.append(a.elapsed_time(b)); tr.append(b.elapsed_time(c))
        out[f"rank{rank}"] = {"closest_share": nc, "ray_share": nr, "closest_call_ms": round(float(np.median(tc)), 3), "ray_call_ms": round(float(np.median(tr)), 3),
                              "closest_Mqps": round(nc / np.median(tc) / 1e3, 1), "ray_Mqps": round(nr / np.median(tr) / 1e3, 1)}
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
