"""The header-only C++ adaptor (include/thea_gpu/GpuBvh.hpp, Callers.hpp): compiles against the C ABI on any box; on
the GPU box a C++ driver written like a Thea caller must reproduce the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from thea_b200 import gen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "cpp", "adaptor_driver")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s"])


def test_adaptor_compiles_and_links(built):
    _build()
    assert os.path.exists(DRIVER)


def test_adaptor_headers_are_self_contained(built, tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text('#include "thea_gpu/Callers.hpp"\nint main() { return sizeof(TheaGpu::GpuMeshBvh) > 0 ? 0 : 1; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src)])


def test_local_feature_headers_and_host_epilogues(built, tmp_path):
    """LocalFeatures.hpp compiles on its own; its 3x3 symmetric eigen solver (the stand-in for Eigen's SelfAdjointEigenSolver in
    LocalPca) and the graph file reader / writer of Callers.hpp run on the host without an engine."""
    src = tmp_path / "lf.cpp"
    src.write_text(r"""
#include "thea_gpu/LocalFeatures.hpp"
#include "thea_gpu/Callers.hpp"
#include <cstdio>
#include <cmath>
using namespace TheaGpu;
int main(int argc, char ** argv)
{
  // eigen pairs of a matrix with known spectrum: Q diag(5, 2, 0.5) Q^T
  double c = std::cos(0.7), s = std::sin(0.7), c2 = std::cos(-0.3), s2 = std::sin(-0.3);
  double Q[9] = {c, -s * c2, s * s2, s, c * c2, -c * s2, 0, s2, c2}, L[3] = {5, 2, 0.5}, A[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { A[3 * i + j] = 0; for (int k = 0; k < 3; ++k) A[3 * i + j] += Q[3 * i + k] * L[k] * Q[3 * j + k]; }
  double w[3], v[9];
  SurfaceFeatures::Local::detail_lf::symmetricEigen3(A, w, v);
  for (int k = 0; k < 3; ++k)
  {
    if (std::fabs(w[k] - L[k]) > 1e-12) { std::printf("eigenvalue %d: %.17g\n", k, w[k]); return 1; }
    for (int i = 0; i < 3; ++i)
    {
      double av = 0; for (int j = 0; j < 3; ++j) av += A[3 * i + j] * v[3 * k + j];
      if (std::fabs(av - w[k] * v[3 * k + i]) > 1e-12) { std::printf("eigenvector %d\n", k); return 1; }
    }
  }
  double Z[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  SurfaceFeatures::Local::detail_lf::symmetricEigen3(Z, w, v);
  if (w[0] != 0 || w[1] != 0 || w[2] != 0) return 1;
  // graph file: write, read back, separations recomputed when absent, malformed input reported
  int32_t idx[12] = {1, 2, -1, -1, 0, -1, -1, -1, -1, -1, -1, -1}; float sep[12] = {0.5f, 1.5f, -1, -1, 0.5f, -1, -1, -1, -1, -1, -1, -1};
  std::string path = argv[1];
  if (!saveGraph(path, 3, 4, idx, sep, true)) return 2;
  int32_t md = 0; std::vector<int32_t> li; std::vector<float> ls; std::string err;
  if (!loadGraph(path, 3, nullptr, md, li, ls, &err)) { std::printf("%s\n", err.c_str()); return 2; }
  if (md != 4 || li != std::vector<int32_t>(idx, idx + 12) || ls[0] != 0.5f || ls[1] != 1.5f || ls[4] != 0.5f || ls[2] != -1.0f) return 2;
  float pos[9] = {0, 0, 0, 3, 4, 0, 0, 0, 1};
  if (!saveGraph(path, 3, 4, idx) || !loadGraph(path, 3, pos, md, li, ls, &err)) return 3;
  if (li[0] != 2 || ls[0] != 1.0f || li[1] != 1 || ls[1] != 5.0f || li[4] != 0 || ls[4] != 5.0f) return 3;   // re-sorted by (separation, index)
  // the same file for a two-sample set names neighbour 2: refused with the reference's message (PointSet3.cpp:271-277)
  if (loadGraph(path, 2, pos, md, li, ls, &err) || err.find("valid neighbor") == std::string::npos) return 4;
  return 0;
}
""")
    exe = tmp_path / "lf"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", os.path.join(ROOT, "thea_b200"), "-ltgq", "-Wl,-rpath," + os.path.join(ROOT, "thea_b200"),
                           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    r = subprocess.run([str(exe), str(tmp_path / "g.graph")], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout)


def test_driver_fails_loudly_without_a_gpu(built, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _build()
    inp = tmp_path / "in.bin"
    inp.write_bytes(struct.pack("<5q", 0, 0, 0, 0, 0))
    r = subprocess.run([DRIVER, str(inp), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode != 0 and "no engine" in r.stderr


@pytest.mark.gpu
def test_adaptor_driver_matches_oracle(built, tmp_path):
    from oracle import orc
    _build()
    mesh = gen.torus_mesh(48)
    T = len(mesh)
    b = gen.mesh_bounds(mesh)
    Q, R, S, NS = 5000, 5000, 1500, 600
    q = gen.closest_queries(Q, b)
    o, d = gen.random_rays(R, b)
    p, n, _ = gen.area_samples(mesh, S)
    Rm, t = gen.rigid_pose()
    src = gen.apply_pose(p, Rm, t)
    sp, sn, _ = gen.area_samples(mesh, NS, seed=99)
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<5q", T, Q, R, S, NS))
        for a in (mesh, q, o, d, src, sp, sn):
            f.write(np.ascontiguousarray(a, np.float32).tobytes())
    r = subprocess.run([DRIVER, str(inp), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    buf = open(out, "rb").read()
    pos = [0]

    def take(dtype, count):
        a = np.frombuffer(buf, dtype=dtype, count=count, offset=pos[0])
        pos[0] += a.nbytes
        return a

    ora = orc.Oracle(tris9=mesh)
    oi, od, oc = ora.closest(q, mode=orc.MODE_BRUTE)
    assert np.array_equal(take(np.int32, Q), oi.astype(np.int32))
    assert np.array_equal(take(np.float32, Q), od)
    assert np.array_equal(take(np.float32, 3 * Q).reshape(-1, 3), oc)
    ns = 16
    assert np.array_equal(take(np.int64, ns), oi[:ns])
    dist = take(np.float64, ns)
    assert np.array_equal(dist, np.sqrt(od[:ns].astype(np.float64)))        # BvhN.hpp:1058
    assert np.array_equal(take(np.float64, ns), dist)
    assert np.array_equal(take(np.float32, 3 * ns).reshape(-1, 3), oc[:ns])

    ori, ort = ora.ray_first(o, d, mode=orc.MODE_BRUTE)
    rt, ri = take(np.float32, R), take(np.int32, R)
    rn, rh = take(np.float32, 3 * R).reshape(-1, 3), take(np.uint8, R)
    assert np.array_equal(rt, ort) and np.array_equal(ri, ori.astype(np.int32))
    on = ora.tri_cache()[0]
    hit = ori >= 0
    assert np.array_equal(rh, hit.astype(np.uint8))
    assert np.array_equal(rn[hit], on[ori[hit]]) and not rn[~hit].any()
    assert np.array_equal(take(np.float32, ns), ort[:ns])
    assert np.array_equal(take(np.int64, ns), ori[:ns])
    assert np.array_equal(take(np.float32, 3 * ns).reshape(-1, 3), rn[:ns])
    assert np.array_equal(take(np.float32, ns), ort[:ns])
    assert np.array_equal(take(np.uint8, ns), hit[:ns].astype(np.uint8))

    trv = take(np.float32, 13)
    batches = int(take(np.int64, 1)[0])
    corr = take(np.int32, batches * S).reshape(batches, S)
    otr, oerr, ocorr = ora.icp(src, mode=orc.MODE_BRUTE)
    assert batches == ocorr.shape[0]
    assert np.array_equal(corr[0], ocorr[0].astype(np.int32))
    assert np.allclose(trv[:12].reshape(3, 4), otr, rtol=1e-5, atol=1e-5)
    assert abs(trv[12] - oerr) <= 1e-4 * abs(oerr) + 1e-9

    diag = float(take(np.float32, 1)[0])
    sdf = take(np.float64, NS)
    osdf = ora.sdf(sp, sn, diag, True, mode=orc.MODE_BRUTE)[0]
    bad = osdf < 0
    if bad.any():
        osdf = osdf.copy()
        osdf[bad] = ora.sdf(sp[bad], sn[bad], diag, False, mode=orc.MODE_BRUTE)[0]
    assert np.array_equal(sdf, osdf)

    po = orc.Oracle(points=sp, normals=sn)
    K = 8
    oki, okd = po.knn(sp, K, exclude_self=True, mode=orc.MODE_BRUTE)
    assert np.array_equal(take(np.int32, NS * K).reshape(NS, K), oki.astype(np.int32))
    assert np.array_equal(take(np.float32, NS * K).reshape(NS, K), okd)
    ski, skd = take(np.int64, 12).reshape(4, 3), take(np.float64, 12).reshape(4, 3)
    o3i, o3d = po.knn(sp[:4], 3, mode=orc.MODE_BRUTE)
    assert np.array_equal(ski, o3i) and np.array_equal(skd, o3d.astype(np.float64))
    seeds = np.array([(j * 7919) % NS for j in range(100)], np.int64)
    avg = float(take(np.float32, 1)[0])
    assert avg == po.avg_separation(seeds)
    ogi, ogs, ogc = po.ball_graph(6, float(np.float32(np.sqrt(np.float32(6))) * np.float32(avg)), mode=orc.MODE_BRUTE)
    assert np.array_equal(take(np.int32, NS * 6).reshape(NS, 6), ogi.astype(np.int32))
    assert np.array_equal(take(np.float32, NS * 6).reshape(NS, 6), ogs)
    assert np.array_equal(take(np.int32, NS), ogc)
    labels = (np.arange(NS) % 3).astype(np.int32)
    po.set_labels(labels)
    ofi, ofd, _ = po.closest_filtered(sp, sn, labels, mode=orc.MODE_BRUTE)
    assert np.array_equal(take(np.int32, NS), ofi.astype(np.int32)) and np.array_equal(take(np.float32, NS), ofd)
    oki, okd = po.knn_filtered(sp, 4, sn, labels, mode=orc.MODE_BRUTE)
    assert np.array_equal(take(np.int32, NS * 4).reshape(NS, 4), oki.astype(np.int32))
    assert np.array_equal(take(np.float32, NS * 4).reshape(NS, 4), okd)
    osi, oss, osc = po.ball_graph_surface(ora, 6, float(np.float32(np.sqrt(np.float32(6))) * np.float32(avg)))
    assert np.array_equal(take(np.int32, NS * 6).reshape(NS, 6), osi.astype(np.int32))
    assert np.array_equal(take(np.float32, NS * 6).reshape(NS, 6), oss)
    assert np.array_equal(take(np.int32, NS), osc)
    # local surface features through the SurfaceFeatures::Local classes (tolerances: tests/test_local_features.py)
    fa, fc = take(np.float64, NS), take(np.float64, NS)
    fe, fv = take(np.float32, 3 * NS).reshape(NS, 3), take(np.float32, 9 * NS).reshape(NS, 3, 3)
    dg = np.float32(diag)
    ra, rc = np.float32(np.float32(0.1) * dg), np.float32(np.float32(0.05) * dg)
    ref_a = po.ball_stats(sp, ra, sn)
    assert np.allclose(fa, ref_a["avg_dist"] / np.float64(ra), rtol=1e-12)
    ref_c = po.ball_stats(sp, rc, sn)
    assert np.abs(fc - ref_c["curvature"]).max() <= 2e-5
    c6 = ref_c["cov"]
    cov = np.stack([c6[:, 0], c6[:, 1], c6[:, 2], c6[:, 1], c6[:, 3], c6[:, 4], c6[:, 2], c6[:, 4], c6[:, 5]], axis=1).reshape(NS, 3, 3)
    w = np.linalg.eigvalsh(cov)[:, ::-1]
    assert np.abs(fe - w).max() <= 2e-5 * float(rc) ** 2
    rec = np.einsum("qk,qki,qkj->qij", fe.astype(np.float64), fv.astype(np.float64), fv.astype(np.float64))
    assert np.abs(rec - cov).max() <= 4e-5 * float(rc) ** 2
    assert np.array_equal(take(np.float64, 8 * NS).reshape(NS, 8), po.ball_hist(sp, ra, 8))
    vscale = float(take(np.float32, 1)[0])
    assert vscale == diag
    vis = take(np.float64, NS)
    ovis, _ = ora.visibility(sp, d[:32], vscale, mode=orc.MODE_BRUTE)
    assert np.array_equal(vis, ovis)
    # SpinImage, the GEODESIC features, the oversampled graph and furthest point sampling
    assert np.array_equal(take(np.float64, 20 * NS).reshape(NS, 5, 4), po.spin_image(sp, sn, 5, 4, np.float32(0.75) * dg))
    ga, gh = take(np.float64, NS), take(np.float64, 8 * NS).reshape(NS, 8)
    gseed = po.closest(sp, mode=orc.MODE_BRUTE)[0]
    rs, rcnt, rh = orc.graph_geodesic(ogi, ogs, gseed, ra, ra, 8)
    assert np.array_equal(gh, rh)
    assert np.allclose(ga, np.where(rcnt > 0, rs / np.maximum(rcnt, 1), 0.0) / np.float64(ra), rtol=1e-12, atol=0)
    NO = NS // 3
    oci, ocs, occ = orc.graph_contract(ogi, ogs, NO)
    assert np.array_equal(take(np.int32, NO * 6).reshape(NO, 6), oci.astype(np.int32))
    assert np.array_equal(take(np.float32, NO * 6).reshape(NO, 6), ocs)
    assert np.array_equal(take(np.int32, NO), occ)
    g8i, g8s, _ = orc.Oracle(points=sp).ball_graph(8, float(np.float32(np.sqrt(np.float32(8))) * np.float32(avg)), mode=orc.MODE_BRUTE)
    ofps = orc.graph_fps(g8i, g8s, NS // 10)
    assert int(take(np.int64, 1)[0]) == len(ofps)
    assert np.array_equal(take(np.int32, NS // 10)[:len(ofps)], ofps.astype(np.int32))
    assert pos[0] == len(buf)


@pytest.mark.gpu
def test_reference_caller_templates_run_on_the_adaptor(built, tmp_path):
    """tests/cpp/thea_callers_driver: Thea's OWN ShapeDiameter<Mesh, ExternalBvhT>, Icp3<Real>::align(n, from, structure) and the
    range / nearest-neighbour call shapes of Test/TestBvh.cpp, compiled unmodified with TheaGpu::GpuMeshBvhT<Thea::Vector3> /
    GpuPointBvhT<Thea::Vector3> as the structure, against the same templates on Thea's BvhN.  The binary is built where
    /root/reference exists and travels to the GPU box."""
    exe = os.path.join(ROOT, "tests", "cpp", "thea_callers_driver")
    if not os.path.exists(exe):
        pytest.skip("thea_callers_driver was not built (needs /root/reference at build time)")
    mesh = gen.torus_mesh(40)
    S, NI, P = 400, 1500, 6000
    sp, sn, _ = gen.area_samples(mesh, S)
    p, _, _ = gen.area_samples(mesh, NI, seed=5)
    src = gen.apply_pose(p, *gen.rigid_pose())
    pts, _, _ = gen.area_samples(mesh, P, seed=6)
    inp = tmp_path / "in.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<4q", len(mesh), S, NI, P))
        for a in (mesh, sp, sn, src, pts):
            f.write(np.ascontiguousarray(a, np.float32).tobytes())
    r = subprocess.run([exe, str(inp)], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert '"fails": 0' in r.stdout
