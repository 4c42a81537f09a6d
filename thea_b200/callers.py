"""Batched rewrites of the reference's serial query loops on top of Engine (the Python twin of include/thea_gpu/Callers.hpp).

    icp_align        <- Algorithms::Icp3<Real>::align               (Algorithms/Icp3.hpp:110-126,193-275)
    shape_diameter   <- loop over ShapeDiameter::compute (+ retry)   (Tools/MeshLocalFeatures/MeshLocalFeatures.cpp:747-764)
    sample_graph     <- PointSet3::updateGraph                       (Algorithms/PointSet3.hpp:324-391)

Every query batch and every reduction runs on the GPU through the C ABI; only the O(1) epilogues (3x3 polar
decomposition, transform composition, convergence test) are host arithmetic.  Nothing here imports the oracle.
"""
from __future__ import annotations

import numpy as np

F32_MIN = np.finfo(np.float32).tiny


def polar_rotation(M: np.ndarray) -> np.ndarray:
    """R = U V^T of a 3x3 matrix (what Eigen::JacobiSVD gives Icp3.hpp:369-370), via one-sided Jacobi in float64."""
    A = np.array(M, np.float64).reshape(3, 3).copy()
    V = np.eye(3)
    for _ in range(64):
        worst = 0.0
        for p in range(2):
            for q in range(p + 1, 3):
                app, aqq, apq = A[:, p] @ A[:, p], A[:, q] @ A[:, q], A[:, p] @ A[:, q]
                den = np.sqrt(app * aqq)
                if den > 0:
                    worst = max(worst, abs(apq) / den)
                if apq == 0:
                    continue
                zeta = (aqq - app) / (2.0 * apq)
                t = (1.0 if zeta >= 0 else -1.0) / (abs(zeta) + np.sqrt(1.0 + zeta * zeta))
                c = 1.0 / np.sqrt(1.0 + t * t)
                s = c * t
                for X in (A, V):
                    x, y = X[:, p].copy(), X[:, q].copy()
                    X[:, p], X[:, q] = c * x - s * y, s * x + c * y
        if worst < 1e-15:
            break
    sv = np.sqrt((A * A).sum(axis=0))
    U = np.where(sv > 0, A / np.where(sv > 0, sv, 1.0), 0.0)
    mn = int(np.argmin(sv))
    if sv[mn] <= 1e-14 * sv.max():
        a, b = (mn + 1) % 3, (mn + 2) % 3
        U[:, mn] = np.cross(U[:, a], U[:, b])
    return U @ V.T


def _compose(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """a * b for row-major 3x4 [L|t] in fp32 with Eigen's t0 + (t1 + t2) order (AffineTransformN.hpp:154-157)."""
    a = a.astype(np.float32)
    b = b.astype(np.float32)
    r = np.empty((3, 4), np.float32)
    for i in range(3):
        for j in range(3):
            r[i, j] = a[i, 0] * b[0, j] + (a[i, 1] * b[1, j] + a[i, 2] * b[2, j])
        r[i, 3] = (a[i, 0] * b[0, 3] + (a[i, 1] * b[1, 3] + a[i, 2] * b[2, 3])) + a[i, 3]
    return r


def icp_align(eng, from3, min_iterations=3, max_iterations=10, fractional_error_threshold=1e-4, want_error=True, reduce_sums=None):
    """Icp3<Real>::align(n, from, to_structure, &error) against the triangle structure held by `eng`.

    Returns (transform 3x4 float32, error, correspondences [batches x n] int32).  `reduce_sums(out16) -> out16` lets a
    multi-GPU driver combine the per-shard sums (see thea_b200.sharded); default: single GPU.
    """
    orig = np.ascontiguousarray(from3, np.float32).reshape(-1, 3)
    n = orig.shape[0]
    ident = np.eye(3, 4, dtype=np.float32)
    if n == 0:
        return ident, 0.0, np.empty((0, 0), np.int32)
    cur = orig.copy()
    corr = []
    state = {}

    def sums(tr):
        s = eng.icp_sums(cur, state["to"], tr)
        return reduce_sums(s) if reduce_sums is not None else s

    def measure(tr):                                    # Icp3::measureError, Icp3.hpp:387-464
        idx, _, cp = eng.closest(cur)
        if (idx < 0).any():
            raise RuntimeError(f"Icp3: Couldn't get nearest neighbor of source point {int(np.argmax(idx < 0))}")
        corr.append(idx.copy())
        state["to"] = cp
        return np.float32(sums(tr)[15])

    old_tr = ident.copy()
    old_err = measure(old_tr)
    if old_err <= F32_MIN:
        return old_tr, float(old_err), np.stack(corr)
    new_err, new_tr = old_err, old_tr
    for i in range(max_iterations):
        # alignOneStep(from_points, to_points): the CURRENT source positions against the targets found by the last
        # measureError, which ran before the positions were updated (Icp3.hpp:228-237,260-261) -- reproduced as is
        s = sums(None)
        R = polar_rotation(s[6:15].reshape(3, 3)).astype(np.float32)
        pm, qm = s[0:3].astype(np.float32), s[3:6].astype(np.float32)
        inc = np.empty((3, 4), np.float32)
        inc[:, :3] = R
        for r in range(3):
            inc[r, 3] = qm[r] - (R[r, 0] * pm[0] + (R[r, 1] * pm[1] + R[r, 2] * pm[2]))
        new_tr = _compose(inc, old_tr)
        if i < max_iterations - 1 or want_error:
            new_err = measure(inc)
        if i < max_iterations - 1:
            frac = (old_err - new_err) / old_err
            if i >= min_iterations and frac < fractional_error_threshold:
                if frac > 0:
                    return new_tr, float(new_err), np.stack(corr)
                return old_tr, float(old_err), np.stack(corr)
            cur = eng.transform_points(orig, new_tr)
        old_err, old_tr = new_err, new_tr
    return new_tr, float(new_err), np.stack(corr)


def shape_diameter(eng, p3, n3, scale, retry=True):
    """compute(p, n, true) for every sample and, like MeshLocalFeatures.cpp:752-756, compute(p, n, false) where that
    gave a negative value."""
    p3 = np.ascontiguousarray(p3, np.float32).reshape(-1, 3)
    n3 = np.ascontiguousarray(n3, np.float32).reshape(-1, 3)
    sdf, _, _ = eng.sdf(p3, n3, scale, True, want_rays=False)
    if retry:
        bad = np.flatnonzero(sdf < 0)
        if len(bad):
            again, _, _ = eng.sdf(p3[bad], n3[bad], scale, False, want_rays=False)
            sdf[bad] = again
    return sdf


def sample_graph(eng, max_degree, seeds):
    """PointSet3::updateGraph's neighbour search over the point structure held by `eng`; returns (idx, sep, count, avg_sep)."""
    avg = eng.avg_separation(seeds)
    radius = np.float32(np.sqrt(np.float32(max_degree))) * np.float32(avg)
    idx, sep, cnt = eng.ball_graph(max_degree, float(radius))
    return idx, sep, cnt, avg
