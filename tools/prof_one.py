"""Development aid: one query kind on one mesh, a few launches (target of ncu captures)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from thea_b200 import Engine, gen

def main(kind, grid, Q, reps=3):
    mesh = gen.torus_mesh(grid); b = gen.mesh_bounds(mesh)
    e = Engine(0, lib=os.environ.get("TGQ_LIB") or None)
    if os.environ.get("TGQ_POOL") is not None:
        e.set_option(5, int(os.environ["TGQ_POOL"]))
    if os.environ.get("TGQ_WIDE") is not None:
        e.set_option(7, int(os.environ["TGQ_WIDE"]))
    e.build_triangles(torch.from_numpy(mesh).cuda())
    if kind == "closest":
        q = torch.from_numpy(gen.closest_queries(Q, b)).cuda()
        for _ in range(reps):
            e.closest(q)
    else:
        o, d = gen.random_rays(Q, b)
        o = torch.from_numpy(o).cuda(); d = torch.from_numpy(d).cuda()
        for _ in range(reps):
            e.ray_first(o, d)
    s = e.stats()
    print(json.dumps({"kind": kind, "tris": len(mesh), "Q": Q, "ms": s["last_query_ms"], "sort_ms": s["last_sort_ms"],
                      "Mqps": Q / (s["last_query_ms"] - s["last_sort_ms"]) / 1e3, "build_ms": s["last_build_ms"]}))

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
