"""Development aid: the per-field (tgq_closest / tgq_ray_first) and the record-output entry points on the same 50 M cfg5 batches."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from thea_b200 import Engine, gen

def main(N=50_000_000):
    dev = torch.device("cuda", 0)
    s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s)
    mesh = gen.torus_mesh(gen.TORUS_10M); bounds = gen.mesh_bounds(mesh)
    e = Engine(0); e.set_stream(s.cuda_stream)
    if os.environ.get("TGQ_RAY_ORDER"): e.set_option(9, int(os.environ["TGQ_RAY_ORDER"]))
    e.build_triangles(torch.from_numpy(mesh).to(dev))
    q = torch.empty((N, 3), dtype=torch.float32, device=dev); ro = torch.empty_like(q); rd = torch.empty_like(q)
    for b, n in bench.blocks(N):
        q[b:b + n] = bench.torch_closest_queries(n, bounds, gen.SEED_CLOSEST, b, dev)
        o, d = bench.torch_random_rays(n, bounds, gen.SEED_RAYS, b, dev); ro[b:b + n] = o; rd[b:b + n] = d
    crec = torch.empty((N, 8), dtype=torch.float32, device=dev); rrec = torch.empty((N, 2), dtype=torch.float32, device=dev)
    idx = torch.empty(N, dtype=torch.int32, device=dev); d2 = torch.empty(N, dtype=torch.float32, device=dev); cp = torch.empty((N, 3), dtype=torch.float32, device=dev)
    t = torch.empty(N, dtype=torch.float32, device=dev); ri = torch.empty(N, dtype=torch.int32, device=dev)
    out = {}
    tests = (("closest_fields", lambda: e.closest(q, out=(idx, d2, cp))), ("closest_records", lambda: e.closest_records(q, [crec])),
                     ("ray_fields", lambda: e.ray_first(ro, rd, out=(t, ri, None))), ("ray_records", lambda: e.ray_first_records(ro, rd, [rrec])))
    if os.environ.get("TGQ_RAY_ORDER"): tests = tests[2:3]
    for name, fn in tests:
        ms = []
        for it in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            if it >= 2: ms.append(a.elapsed_time(b))
        out[name] = {"call_ms": round(float(np.median(ms)), 2), "Mqps": round(N / np.median(ms) / 1e3, 1)}
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    main()
