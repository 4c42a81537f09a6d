"""Development aid: timings of the SURVEY 8f rank 4 query objects (tgq_closest_segment, tgq_closest_triangle, tgq_closest_pair between
meshes) on the 250 k-triangle torus and on 1 M surface samples.  Kernel time from the engine's own CUDA events (tgq_get_stats:
last_query_ms includes the Morton ordering of the batch); device-resident inputs, host outputs."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from thea_b200 import Engine, gen


def main(out):
    res = {}
    mesh = gen.torus_mesh(354)                       # 250,632 triangles
    e = Engine(0); e.build_triangles(mesh)
    p, n, _ = e.area_sample(1_000_000, gen.SEED_SAMPLES)
    pe = Engine(0); pe.build_points(p)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    Q = 1_000_000
    base = torch.from_numpy(p).cuda()[torch.randint(0, len(p), (Q,), device="cuda", generator=g)]
    a = (base + 0.02 * torch.randn((Q, 3), device="cuda", generator=g)).contiguous()
    b = (a + 0.02 * torch.randn((Q, 3), device="cuda", generator=g)).contiguous()
    for name, eng in (("triangles", e), ("points", pe)):
        eng.closest_segment(a[:1000], b[:1000])
        t0 = time.perf_counter(); r = eng.closest_segment(a, b); wall = time.perf_counter() - t0
        ms = eng.stats()["last_query_ms"]
        res[f"segment_queries_on_{name}"] = {"Q": Q, "kernel_ms": ms, "Mqueries_per_s": Q / ms / 1e3, "wall_s_incl_download": wall, "found": int((r[0] >= 0).sum())}
    Qt = 250_000
    c = a[:Qt].reshape(Qt, 1, 3)
    q9 = (c + 0.01 * torch.randn((Qt, 3, 3), device="cuda", generator=g)).reshape(Qt, 9).contiguous()
    for name, eng in (("triangles", e), ("points", pe)):
        eng.closest_triangle(q9[:1000])
        r = eng.closest_triangle(q9)
        ms = eng.stats()["last_query_ms"]
        res[f"triangle_queries_on_{name}"] = {"Q": Qt, "kernel_ms": ms, "Mqueries_per_s": Qt / ms / 1e3, "intersecting": int((r[1] == 0).sum())}
    # mesh against a copy pushed 0.01 off the surface: 250 k x 250 k triangles
    nrm = e.triangle_cache()[0]
    other = (mesh.reshape(-1, 3, 3) + 0.01 * nrm[:, None, :]).astype(np.float32).reshape(-1, 9)
    eo = Engine(0); eo.build_triangles(other)
    torch.cuda.synchronize(); t0 = time.perf_counter(); pair = e.closest_pair(eo); torch.cuda.synchronize()
    res["mesh_to_mesh_closest_pair"] = {"triangles": [len(mesh), len(other)], "wall_s": time.perf_counter() - t0, "d2": pair[2]}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/query_object_timings.json")
