"""Times the SURVEY 8f rank-3 local surface features through the public API on one GPU (development aid, like config_timings.py):
ball statistics (AverageDistance / Curvature / LocalPca in one pass) over area samples of the 1M-triangle torus, and Visibility
ray fans against the mesh.  Each is spot-checked against the oracle on a sample."""
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
    mesh = gen.torus_mesh(gen.TORUS_1M); b = gen.mesh_bounds(mesh)
    e = Engine(0); e.build_triangles(torch.from_numpy(mesh).cuda())
    p, n, _ = e.area_sample(1_000_000, gen.SEED_SAMPLES)
    pe = Engine(0); pe.set_option(2, int(os.environ.get("TGQ_LEAF", "0"))); pe.build_points(p, n)
    lo, hi = b; scale = float(np.linalg.norm((hi - lo).astype(np.float32)))
    dp, dn = torch.from_numpy(p).cuda(), torch.from_numpy(n).cuda()
    po = orc.Oracle(points=p, normals=n)
    for name, Q, frac in (("r=0.02*scale, 1M balls", 1_000_000, 0.02), ("r=0.1*scale (reference default), 100k balls", 100_000, 0.1)):
        r = np.float32(frac * scale)
        rad = torch.full((Q,), float(r), dtype=torch.float32, device="cuda")
        out = torch.empty((Q, 16), dtype=torch.float64, device="cuda")
        dt, _ = timed(lambda: pe.ball_stats(dp[:Q], rad, dn[:Q], out=out))
        st = out.cpu().numpy()
        sub = np.arange(0, Q, max(1, Q // 40))
        ref = po.ball_stats(p[sub], r, n[sub])
        curv = callers.projected_curvature(None, p[sub], n[sub], r, stats=st[sub])
        res[name] = {"balls": Q, "mean_points_per_ball": float(st[:, 0].mean()), "seconds_device_resident": dt, "balls_per_s_M": Q / dt / 1e6,
                     "point_visits_per_s_G": float(st[:, 0].sum()) / dt / 1e9,
                     "sample_counts_match_oracle": bool(np.array_equal(st[sub, 0].astype(np.int64), ref["n"])),
                     "sample_curvature_max_abs_diff": float(np.abs(curv - ref["curvature"]).max())}
    # Visibility: 100k samples x 100 fixed directions against the 1M-triangle mesh
    rng = np.random.default_rng(1)
    dirs = rng.normal(size=(100, 3)); dirs = (dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32)
    S = 100_000
    dt, (vis, nesc) = timed(lambda: e.visibility(p[:S], dirs, scale))
    ovis, _ = orc.Oracle(tris9=mesh).visibility(p[:100], dirs, scale, mode=orc.MODE_CANON)
    res["visibility"] = {"samples": S, "rays": 100 * S, "seconds_e2e_host_buffers": dt, "rays_per_s_M": 100 * S / dt / 1e6, "mean_visibility": float(vis.mean()),
                         "sample_matches_oracle": bool(np.array_equal(vis[:100], ovis))}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
