"""Deterministic synthetic workloads for the spatial-query path.

Counter-based splitmix64 so that any slice of a stream can be generated independently (per rank, per
block) with identical values: ``u01(seed)[i] = float32(splitmix64(seed + (i+1)*GOLDEN) >> 40) * 2**-24``.
Meshes are bumpy tori on an n x n quad grid (2 n^2 triangles, closed, outward oriented).  The reference's
own tools read meshes and sample files from disk (Tools/MeshLocalFeatures/MeshLocalFeatures.cpp:241-257,
Tools/SampleGraph/SampleGraph.cpp:371-444); the engine takes flat float32 arrays, which is what these return.
"""
from __future__ import annotations

import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
SEED_MESH = 0xB2000001
SEED_CLOSEST = 0xB2000002
SEED_RAYS = 0xB2000003
SEED_SAMPLES = 0xB2000004
SEED_ICP = 0xB2000005

# named grid sizes of the configs
TORUS_70K = 188      # 70,688 triangles
TORUS_1M = 708       # 1,002,528 triangles
TORUS_10M = 2236     # 9,999,392 triangles


def splitmix64(seed: int, n: int, offset: int = 0) -> np.ndarray:
    """Values offset .. offset+n-1 of the counter-based splitmix64 stream ``seed`` (uint64)."""
    with np.errstate(over="ignore"):
        i = np.arange(offset + 1, offset + n + 1, dtype=np.uint64)
        z = np.uint64(seed) + i * GOLDEN
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def u01(seed: int, n: int, offset: int = 0) -> np.ndarray:
    """float32 uniforms in [0, 1) with 24 random bits."""
    return (splitmix64(seed, n, offset) >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)


def torus_vertices(n: int) -> np.ndarray:
    """(n*n, 3) float32 vertices of the bumpy torus, row-major in (i = major angle, j = minor angle)."""
    u = (np.arange(n, dtype=np.float64) * (2.0 * np.pi / n))[:, None]
    v = (np.arange(n, dtype=np.float64) * (2.0 * np.pi / n))[None, :]
    r = 0.4 * (1.0 + 0.15 * np.sin(3 * u) * np.cos(5 * v) + 0.05 * np.sin(11 * u + 2 * v))
    x = (1.0 + r * np.cos(v)) * np.cos(u)
    y = (1.0 + r * np.cos(v)) * np.sin(u)
    z = r * np.sin(v)
    return np.stack([x, y, z], axis=-1).reshape(n * n, 3).astype(np.float32)


def torus_indices(n: int) -> np.ndarray:
    """(2 n^2, 3) int32: per quad (i,j): (i,j)-(i+1,j)-(i+1,j+1) then (i,j)-(i+1,j+1)-(i,j+1)."""
    i = np.arange(n, dtype=np.int64)[:, None]
    j = np.arange(n, dtype=np.int64)[None, :]
    i1 = (i + 1) % n
    j1 = (j + 1) % n
    a = (i * n + j) + 0 * j
    b = (i1 * n + j) + 0 * j
    c = i1 * n + j1
    d = i * n + j1
    tris = np.stack([np.stack([a, b, c], -1), np.stack([a, c, d], -1)], axis=2)  # (n, n, 2, 3)
    return tris.reshape(2 * n * n, 3).astype(np.int32)


def torus_mesh(n: int) -> np.ndarray:
    """(2 n^2, 9) float32 triangle soup, vertex positions shared bit-exactly between neighbours."""
    v = torus_vertices(n)
    f = torus_indices(n)
    return np.ascontiguousarray(v[f].reshape(-1, 9))


def mesh_bounds(tris9: np.ndarray):
    p = tris9.reshape(-1, 3)
    return p.min(axis=0), p.max(axis=0)


def box_scaled(lo, hi, s: float):
    lo = np.asarray(lo, np.float32)
    hi = np.asarray(hi, np.float32)
    c = np.float32(0.5) * (lo + hi)
    h = np.float32(0.5 * s) * (hi - lo)
    return (c - h).astype(np.float32), (c + h).astype(np.float32)


def uniform_points(count: int, lo, hi, seed: int, offset: int = 0) -> np.ndarray:
    """(count, 3) float32 uniform in the box; point k uses draws 3k..3k+2 (offset in points)."""
    u = u01(seed, 3 * count, 3 * offset).reshape(count, 3)
    lo = np.asarray(lo, np.float32)
    hi = np.asarray(hi, np.float32)
    return (lo + u * (hi - lo)).astype(np.float32)


def closest_queries(count: int, tris9_bounds, seed: int = SEED_CLOSEST, offset: int = 0) -> np.ndarray:
    """cfg1/cfg5 closest-point queries: uniform in the mesh AABB scaled x1.5 about its centre."""
    lo, hi = box_scaled(*tris9_bounds, 1.5)
    return uniform_points(count, lo, hi, seed, offset)


def random_rays(count: int, tris9_bounds, seed: int = SEED_RAYS, offset: int = 0):
    """cfg5 rays: origin uniform in the x1.5 AABB, direction = normalised 3 x Gaussian (Box-Muller).

    Ray k uses draws 9k..9k+8: 3 for the origin, 6 for three Box-Muller pairs (cos branch only).
    """
    lo, hi = box_scaled(*tris9_bounds, 1.5)
    u = u01(seed, 9 * count, 9 * offset).reshape(count, 9)
    o = (lo + u[:, 0:3] * (hi - lo)).astype(np.float32)
    u1 = np.maximum(u[:, 3:6].astype(np.float64), 2.0 ** -25)
    u2 = u[:, 6:9].astype(np.float64)
    g = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    nrm = np.sqrt((g * g).sum(axis=1, keepdims=True))
    nrm[nrm == 0] = 1.0
    d = (g / nrm).astype(np.float32)
    return o, d


def tri_cache_np(tris9: np.ndarray):
    """float32 numpy restatement of the per-triangle cache used by the samplers (unit normal, area).

    Same operation order as Triangle3.hpp:143-155 / Hyperplane3.hpp:34-41 (each numpy float32 op is one IEEE op).
    """
    t = tris9.reshape(-1, 3, 3)
    v0, v1, v2 = t[:, 0], t[:, 1], t[:, 2]
    e01 = v1 - v0
    e02 = v2 - v0

    def cross(l, r):
        return np.stack([l[:, 1] * r[:, 2] - l[:, 2] * r[:, 1],
                         l[:, 2] * r[:, 0] - l[:, 0] * r[:, 2],
                         l[:, 0] * r[:, 1] - l[:, 1] * r[:, 0]], axis=1)

    def dot(a, b):
        return a[:, 0] * b[:, 0] + (a[:, 1] * b[:, 1] + a[:, 2] * b[:, 2])

    c = cross(e01, e02)
    z = dot(c, c)
    with np.errstate(divide="ignore", invalid="ignore"):
        n = np.where((z > 0)[:, None], c / np.sqrt(z)[:, None], c).astype(np.float32)
    area = (np.float32(0.5) * np.abs(dot(n, cross(e01, e02)))).astype(np.float32)
    return n, area, e01, e02, v0


def area_samples(tris9: np.ndarray, count: int, seed: int = SEED_SAMPLES):
    """MeshSampler::sampleEvenlyByArea EXACT mode (Algorithms/MeshSampler.hpp:147-185, Triangle3.hpp:179-195)
    with the counter-based generator: sample i consumes draws 3i (triangle), 3i+1 (s), 3i+2 (t).

    Returns positions (count,3) f32, face normals (count,3) f32, triangle index (count,) int64.
    """
    n, area, e01, e02, v0 = tri_cache_np(tris9)
    cum = np.cumsum(area.astype(np.float64))          # sequential double accumulation
    total = cum[-1]
    u = u01(seed, 3 * count).reshape(count, 3)
    t = u[:, 0].astype(np.float64) * total
    tri = np.minimum(np.searchsorted(cum, t, side="left"), len(cum) - 1).astype(np.int64)
    s = u[:, 1].copy()
    w = u[:, 2].copy()
    flip = (s + w) > np.float32(1.0)
    s[flip] = np.float32(1.0) - s[flip]
    w[flip] = np.float32(1.0) - w[flip]
    p = (s[:, None] * e01[tri] + w[:, None] * e02[tri]) + v0[tri]
    return p.astype(np.float32), n[tri].astype(np.float32), tri


def rigid_pose(axis=(1.0, 1.0, 1.0), degrees: float = 5.0, t=(0.02, -0.01, 0.015)):
    """cfg3 ICP start pose: rotation about ``axis`` by ``degrees`` then translation; float32 3x3 and 3-vector."""
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    th = np.deg2rad(degrees)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
    return R.astype(np.float32), np.asarray(t, np.float32)


def apply_pose(p: np.ndarray, R: np.ndarray, t: np.ndarray) -> np.ndarray:
    return (p.astype(np.float64) @ R.astype(np.float64).T + t.astype(np.float64)).astype(np.float32)
