"""Development aid: cfg3-shaped ICP (100k-point source vs the 1M-triangle torus, 51 correspondence batches), timed a few times
per library build / closest-point kernel, to separate kernel effects from one-time costs."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from thea_b200 import Engine, gen


def main():
    mesh = gen.torus_mesh(gen.TORUS_1M)
    dm = torch.from_numpy(mesh).cuda()
    libs = [None] + [os.path.join(ROOT, "build", f) for f in sorted(os.listdir(os.path.join(ROOT, "build"))) if f.startswith("libtgq_")]
    for lib in libs:
        for pool in (1, 0):
            e = Engine(0, lib=lib); e.set_option(5, pool); e.build_triangles(dm)
            src_p, _, _ = e.area_sample(100_000, gen.SEED_SAMPLES + 1)
            R, t = gen.rigid_pose(); src = gen.apply_pose(src_p, R, t)
            times = []
            for _ in range(4):
                t0 = time.perf_counter(); tr, err, nb = e.icp_align(src, min_iterations=50, max_iterations=50, want_correspondences=False); times.append(time.perf_counter() - t0)
            q = torch.from_numpy(gen.closest_queries(100_000, gen.mesh_bounds(mesh))).cuda()
            for _ in range(3):
                e.closest(q)
            ms = e.stats()["last_query_ms"]
            print(json.dumps({"lib": os.path.basename(lib) if lib else "installed", "pool": pool, "icp_s": [round(x, 4) for x in times], "closest_100k_ms": round(ms, 4), "err": err}), flush=True)
            e.close()


if __name__ == "__main__":
    main()
