"""Quick device-side throughput probe (development aid; bench.py is the contract)."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from thea_b200 import Engine, gen

def run(n_grid, Q, leaf=4, sort=2, what=("closest", "ray")):
    mesh = gen.torus_mesh(n_grid)
    b = gen.mesh_bounds(mesh)
    e = Engine(0, lib=os.environ.get("TGQ_LIB") or None)
    e.set_option(2, leaf); e.set_option(1, sort)
    if os.environ.get("TGQ_POOL") is not None:
        e.set_option(5, int(os.environ["TGQ_POOL"]))
    if os.environ.get("TGQ_WIDE") is not None:
        e.set_option(7, int(os.environ["TGQ_WIDE"]))
    e.build_triangles(torch.from_numpy(mesh).cuda())
    st = e.stats()
    out = {"wide": os.environ.get("TGQ_WIDE", "default"), "pool": os.environ.get("TGQ_POOL", "default"), "tris": len(mesh), "Q": Q, "leaf": leaf, "sort": sort, "nodes": st["num_nodes"], "build_ms": round(st["last_build_ms"], 3)}
    if "closest" in what:
        q = torch.from_numpy(gen.closest_queries(Q, b)).cuda()
        for _ in range(3):
            e.closest(q)
        s = e.stats()
        out["closest_ms"] = round(s["last_query_ms"], 3); out["closest_sort_ms"] = round(s["last_sort_ms"], 3)
        out["closest_Mqps"] = round(Q / s["last_query_ms"] / 1e3, 2)
    if "ray" in what:
        o, d = gen.random_rays(Q, b)
        o = torch.from_numpy(o).cuda(); d = torch.from_numpy(d).cuda()
        for _ in range(3):
            e.ray_first(o, d)
        s = e.stats()
        out["ray_ms"] = round(s["last_query_ms"], 3); out["ray_sort_ms"] = round(s["last_sort_ms"], 3)
        out["ray_Mqps"] = round(Q / s["last_query_ms"] / 1e3, 2)
    print(json.dumps(out), flush=True)
    e.close()

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "rays":
        run(2236, 50_000_000, what=("ray",))
    elif len(sys.argv) > 1 and sys.argv[1] == "big":
        run(2236, 50_000_000, what=("closest",))
    elif len(sys.argv) > 1 and sys.argv[1] == "short":
        run(708, 4_000_000)
        run(2236, 8_000_000)
    elif len(sys.argv) > 1 and sys.argv[1] == "leaf":
        for leaf in (1, 2, 4):
            run(708, 4_000_000, leaf=leaf)
            run(2236, 8_000_000, leaf=leaf)
    else:
        run(188, 1_000_000)
        run(708, 4_000_000)
        for leaf in (2, 8):
            run(708, 1_000_000, leaf=leaf)
        run(2236, 8_000_000)
