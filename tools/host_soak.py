"""Development aid (CPU only): randomized soaks of the HOST build of the engine's traversal headers (tests/emu) against the oracle's
exhaustive minimum, over scenes of many scales and positions.  What DESIGN.md 4.1 / 4.3 quote comes from here
(profiles/r04_e_host_soak.txt).

  python tools/host_soak.py main    [n]   closest point (seeded / unseeded), k-NN, first-hit rays
  python tools/host_soak.py points  [n]   sample graph, ball range, k-NN on point structures
  python tools/host_soak.py objects [n]   line-segment and triangle query objects on both structure kinds
  python tools/host_soak.py floor         triangle queries on small-scale scenes with TGQ_TRI_FLOOR_FACTOR = 1, 2, 4, 8
"""
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import emu_lib
from oracle import orc
from thea_b200 import gen


def bits(x):
    return np.ascontiguousarray(x, np.float32).view(np.uint32)


def same(g, w):
    return np.array_equal(g[0], np.asarray(w[0]).astype(np.int32)) and np.array_equal(bits(g[1]), bits(w[1]))


def scene(rng, scales, offsets, grids=(6, 12, 20, 33, 48), soup_every=4, seed=0):
    grid, scale, off = int(rng.choice(grids)), float(rng.choice(scales)), float(rng.choice(offsets))
    mesh = (gen.torus_mesh(grid) * np.float32(scale) + np.float32(off)).astype(np.float32)
    if soup_every and seed % soup_every == 0:
        mesh = (rng.uniform(-1, 1, (int(rng.integers(50, 3000)), 9)) * scale + off).astype(np.float32)
    return mesh, scale, off


def soak_main(n):
    runs = bad = 0
    for seed in range(3000, 3000 + n):
        rng = np.random.default_rng(seed)
        mesh, scale, off = scene(rng, [1e-3, 1e-2, 1.0, 30.0, 1e3], [0.0, 7.0, -300.0, 5000.0], seed=seed)
        v = mesh.reshape(-1, 3); lo, hi = v.min(0), v.max(0); ext = float((hi - lo).max())
        Q = 1500
        q = rng.uniform(lo - 0.5 * ext, hi + 0.5 * ext, (Q, 3)).astype(np.float32)
        q[::7] = (v[rng.integers(0, len(v), len(q[::7]))] + rng.normal(0, 1e-4 * ext, (len(q[::7]), 3))).astype(np.float32)
        e, o = emu_lib.Emu(tris9=mesh), orc.Oracle(tris9=mesh)
        w = o.closest(q, mode=orc.MODE_BRUTE)
        for seeded in (True, False):
            g = e.closest(q, seeded=seeded); runs += 1
            if not (same(g, w) and np.array_equal(g[2], w[2])): bad += 1; print("MISMATCH closest", seed, scale, off, flush=True)
        oo = rng.uniform(lo - 0.5 * ext, hi + 0.5 * ext, (Q, 3)).astype(np.float32)
        d = rng.standard_normal((Q, 3)).astype(np.float32); d[::5] *= np.float32(scale)
        gt, gi = e.ray(oo, d)[:2]; wi, wt = o.ray_first(oo, d, mode=orc.MODE_BRUTE); runs += 1
        if not (np.array_equal(gi, wi.astype(np.int32)) and np.array_equal(bits(gt), bits(wt))): bad += 1; print("MISMATCH ray", seed, scale, off, flush=True)
        pts = v[::2].copy(); ep, op = emu_lib.Emu(points=pts), orc.Oracle(points=pts)
        runs += 1
        if not same(ep.closest(q), op.closest(q, mode=orc.MODE_BRUTE)): bad += 1; print("MISMATCH point closest", seed, flush=True)
        runs += 1
        if not same(ep.knn(q[:500], 9), op.knn(q[:500], 9, mode=orc.MODE_BRUTE)): bad += 1; print("MISMATCH knn", seed, flush=True)
    return runs, bad


def soak_points(n):
    runs = bad = 0
    for seed in range(5000, 5000 + n):
        rng = np.random.default_rng(seed)
        scale, off = float(rng.choice([1e-3, 1.0, 30.0, 1e3])), float(rng.choice([0.0, 7.0, -300.0, 5000.0]))
        P = int(rng.integers(300, 4000))
        p, nr, _ = gen.area_samples(gen.torus_mesh(int(rng.choice([12, 20, 33]))), P, seed=seed)
        p = (p * np.float32(scale) + np.float32(off)).astype(np.float32)
        if seed % 5 == 0: p[P // 2: P // 2 + 20] = p[:20]
        e, o = emu_lib.Emu(points=p, normals=nr), orc.Oracle(points=p, normals=nr)
        avg = o.avg_separation((gen.splitmix64(seed, 100) % np.uint64(P)).astype(np.int64))
        K = int(rng.choice([4, 8, 16])); r0 = float(np.float32(np.sqrt(np.float32(K))) * np.float32(avg))
        gi, gs, _ = e.ball_graph(K, r0); oi, os_, _ = o.ball_graph(K, r0); runs += 1
        if not (np.array_equal(gi, oi.astype(np.int32)) and np.array_equal(bits(gs), bits(os_))): bad += 1; print("MISMATCH graph", seed, scale, off, flush=True)
        Q = 500
        q = (p[rng.integers(0, P, Q)] + rng.normal(0, avg, (Q, 3))).astype(np.float32)
        rad = (avg * rng.uniform(0.5, 4.0, Q)).astype(np.float32)
        ri, rc = e.ball_range(q, rad, max_out=32); wi, wc = o.ball_range(q, rad, max_out=32); runs += 1
        if not (np.array_equal(rc, wc.astype(np.int32)) and np.array_equal(ri, wi.astype(np.int32))): bad += 1; print("MISMATCH range", seed, flush=True)
        for k in (1, 7, 24):
            runs += 1
            if not same(e.knn(q, k), o.knn(q, k, mode=orc.MODE_BRUTE)): bad += 1; print("MISMATCH knn", seed, k, flush=True)
    return runs, bad


def soak_objects(n):
    runs = bad = 0
    for seed in range(1000, 1000 + n):
        rng = np.random.default_rng(seed)
        mesh, scale, off = scene(rng, [1e-2, 1.0, 30.0], [0.0, 7.0, -300.0], grids=(6, 12, 20, 33), soup_every=3, seed=seed)
        v = mesh.reshape(-1, 3); lo, hi = v.min(0), v.max(0); ext = float((hi - lo).max())
        Q = 600
        a = rng.uniform(lo - 0.4 * ext, hi + 0.4 * ext, (Q, 3)).astype(np.float32)
        d = rng.standard_normal((Q, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        b = (a + d * rng.choice([0.0, 1e-4, 0.01, 0.2, 2.0], (Q, 1)) * ext).astype(np.float32)
        c = rng.uniform(lo - 0.3 * ext, hi + 0.3 * ext, (Q, 1, 3))
        q9 = (c + rng.normal(0, 1, (Q, 3, 3)) * rng.choice([0.001, 0.02, 0.2, 1.0], (Q, 1, 1)) * ext).reshape(Q, 9).astype(np.float32)
        q9[:40] = (mesh[rng.integers(0, len(mesh), 40)] + np.float32(1e-3 * ext)).astype(np.float32)
        pts = v[::2].copy()
        excl = rng.integers(0, len(mesh), (Q, 4)); excl[::2, 2:] = -1
        bound = float(rng.choice([-1.0, 0.05, 0.3])) * ext if rng.random() < 0.7 else -1.0
        if bound < 0: bound = -1.0
        for e, o, m, kind in ((emu_lib.Emu(tris9=mesh), orc.Oracle(tris9=mesh), len(mesh), "triangles"), (emu_lib.Emu(points=pts), orc.Oracle(points=pts), len(pts), "points")):
            ex = excl % m
            for g, w, tag in ((e.closest_segment(a, b, bound, ex), o.closest_segment(a, b, bound, ex, mode=orc.MODE_BRUTE), "segment"),
                              (e.closest_triangle(q9, bound), o.closest_triangle(q9, bound, mode=orc.MODE_BRUTE), "triangle")):
                runs += 1
                if not same(g, w):
                    bad += 1
                    k = np.flatnonzero((g[0] != w[0].astype(np.int32)) | (bits(g[1]) != bits(w[1])))
                    print("MISMATCH %s queries on %s: seed %d scale %g offset %g, %d queries, reference d2 of those: %s" % (tag, kind, seed, scale, off, k.size, sorted(set(w[1][k].tolist()))[:3]), flush=True)
    return runs, bad


def floor_table():
    emu_dir = os.path.join(ROOT, "tests", "emu")
    for scale, seeds in ((0.01, 300), (0.03, 120), (0.1, 120), (0.3, 120)):
        for factor in (1, 2, 4, 8):
            so = "/tmp/libtgq_emu_floor%d.so" % factor
            if not os.path.exists(so):
                subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unused-function", "-DTGQ_TRI_FLOOR_FACTOR=%d.0f" % factor,
                                       "-o", so, os.path.join(emu_dir, "emu.cpp")])
            code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
                    "import numpy as np, emu_lib\nemu_lib._NAME = %r\nfrom oracle import orc\nfrom thea_b200 import gen\n"
                    "bad = tot = 0\n"
                    "for seed in range(1000, 1000 + %d):\n"
                    "    rng = np.random.default_rng(seed); grid = int(rng.choice([6, 12, 20, 33])); off = float(rng.choice([0.0, 7.0]))\n"
                    "    mesh = (gen.torus_mesh(grid) * np.float32(%r) + np.float32(off)).astype(np.float32)\n"
                    "    v = mesh.reshape(-1, 3); lo, hi = v.min(0), v.max(0); ext = float((hi - lo).max()); Q = 600\n"
                    "    c = rng.uniform(lo - 0.3 * ext, hi + 0.3 * ext, (Q, 1, 3))\n"
                    "    q9 = (c + rng.normal(0, 1, (Q, 3, 3)) * rng.choice([0.001, 0.02, 0.2, 1.0], (Q, 1, 1)) * ext).reshape(Q, 9).astype(np.float32)\n"
                    "    g = emu_lib.Emu(tris9=mesh).closest_triangle(q9); w = orc.Oracle(tris9=mesh).closest_triangle(q9, mode=orc.MODE_BRUTE)\n"
                    "    bad += int(((g[0] != w[0].astype(np.int32)) | (g[1].view(np.uint32) != w[1].view(np.uint32))).sum()); tot += Q\n"
                    "print(bad, tot)\n") % (ROOT, os.path.join(ROOT, "tests"), os.path.relpath(so, emu_dir), seeds, scale)
            out = subprocess.check_output([sys.executable, "-c", code]).decode().split()
            print("scene scale %-5g floor factor %d: %s disagreements in %s triangle queries" % (scale, factor, out[0], out[1]), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "main"
    t0 = time.time()
    if what == "floor":
        floor_table()
    else:
        n = int(sys.argv[2]) if len(sys.argv) > 2 else {"main": 150, "points": 80, "objects": 600}[what]
        runs, bad = {"main": soak_main, "points": soak_points, "objects": soak_objects}[what](n)
        print("%s: %d runs, %d with a disagreement" % (what, runs, bad))
    print("%.0f s" % (time.time() - t0))
