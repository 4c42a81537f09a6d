"""Batched rewrites of the reference's serial query loops on top of Engine (the Python twin of include/thea_gpu/Callers.hpp).

    icp_align        <- Algorithms::Icp3<Real>::align               (Algorithms/Icp3.hpp:110-126,193-275)
    shape_diameter   <- loop over ShapeDiameter::compute (+ retry)   (Tools/MeshLocalFeatures/MeshLocalFeatures.cpp:747-764)
    sample_graph     <- PointSet3::updateGraph                       (Algorithms/PointSet3.hpp:324-391)
    average_distance <- loop over AverageDistance::compute (EUCLIDEAN) (Algorithms/SurfaceFeatures/Local/AverageDistance.cpp:63-81)
    projected_curvature <- loop over Curvature::computeProjectedCurvature (…/Local/Curvature.cpp:69-98)
    local_pca        <- loop over LocalPca::compute                   (…/Local/LocalPca.cpp:49-62, Algorithms/PcaN.hpp:150-197)
    visibility       <- loop over Visibility::compute                 (…/Local/Visibility.hpp:104-125)
    save_graph / load_graph / save_samples / load_samples <- PointSet3::save / load (Algorithms/PointSet3.cpp:180-375)
    sample_graph_oversampled <- PointSet3::updateGraph with a dense oversampling (PointSet3.hpp:329-386, PointSet3.cpp:96-178)
    geodesic_average_distance / geodesic_distance_histogram <- AverageDistance / LocalDistanceHistogram, GEODESIC
                                                                (…/Local/AverageDistance.cpp:84-118, LocalDistanceHistogram.cpp:90-137)
    spin_images      <- loop over SpinImage::compute                  (…/Local/SpinImage.cpp:30-75)
    furthest_point_subsample <- FurthestPointSampling::subsample     (Algorithms/FurthestPointSampling.cpp:48-125)
    sample_evenly_by_separation <- MeshSampler::sampleEvenlyBySeparation (Algorithms/MeshSampler.hpp:270-329)

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


def sample_graph(eng, max_degree, seeds, surface=None):
    """PointSet3::updateGraph's neighbour search over the point structure held by `eng`; returns (idx, sep, count, avg_sep).
    `surface`: a triangle Engine for the reachability test of updateGraph(surface) (PointSet3.hpp:639-648), or None."""
    avg = eng.avg_separation(seeds)
    radius = np.float32(np.sqrt(np.float32(max_degree))) * np.float32(avg)
    if surface is not None:
        idx, sep, cnt = eng.ball_graph_surface(surface, max_degree, float(radius))
    else:
        idx, sep, cnt = eng.ball_graph(max_degree, float(radius))
    return idx, sep, cnt, avg


def sample_graph_oversampled(eng_all, n_orig, max_degree, seeds):
    """PointSet3::updateGraph with a dense oversampling: `eng_all` holds the n_orig original samples followed by the dense ones; the
    proximity graph over all of them is contracted to the originals by shortest paths (extractOriginalAdjacencies).
    Returns (idx [n_orig x max_degree], sep, count, avg_sep)."""
    idx, sep, _, avg = sample_graph(eng_all, max_degree, seeds)
    ci, cs, cc = eng_all.graph_contract(idx, sep, int(n_orig))
    return ci, cs, cc, avg


def _geodesic(eng, graph_idx, graph_sep, p3, max_distance, scale, num_bins):
    seed = eng.closest(p3, want_cp=False)[0]
    rng = np.float32(scale) if max_distance < 0 else np.float32(max_distance)
    return eng.graph_geodesic(graph_idx, graph_sep, np.ascontiguousarray(seed, np.int32), np.float32(max_distance), rng, num_bins), rng


def geodesic_average_distance(eng, graph_idx, graph_sep, p3, max_distance, scale):
    """AverageDistance::compute(p, GEODESIC, max_distance): seed = nearest sample, mean graph distance over the samples within
    max_distance of it (< 0: the whole component), divided by max_distance (or the scale)."""
    (s, c, _), rng = _geodesic(eng, graph_idx, graph_sep, p3, max_distance, scale, 0)
    return np.where(c > 0, s / np.maximum(c, 1), 0.0) / np.float64(rng)


def geodesic_distance_histogram(eng, graph_idx, graph_sep, p3, num_bins, max_distance, scale):
    """LocalDistanceHistogram::compute(p, histogram, GEODESIC, max_distance) with every sample accepted: [Q x num_bins] counts."""
    (_, _, h), _ = _geodesic(eng, graph_idx, graph_sep, p3, max_distance, scale, num_bins)
    return h.astype(np.float64)


def spin_images(eng, p3, n3, num_radial_bins, num_height_bins, scale, sample_normals=None):
    """SpinImage::compute(position[, normal], ...) for every query: [Q x radial x height] float64.  n3 None: the normal of the sample
    nearest to each query (sample_normals, P x 3, must then be given)."""
    if n3 is None:
        if sample_normals is None:
            raise ValueError("SpinImage: Cannot infer query normal from surface lacking normals")
        seed = eng.closest(p3)[0]
        n3 = np.where((seed >= 0)[:, None], np.asarray(sample_normals, np.float32)[np.maximum(seed, 0)], np.float32(0))
    img = eng.spin_image(p3, n3, num_radial_bins, num_height_bins, np.float32(0.75) * np.float32(scale))
    return img.astype(np.float64)


def furthest_point_subsample(points3, num_desired, seeds, max_degree=8, device=0):
    """FurthestPointSampling::subsample(..., GEODESIC): graph over the points (PointSet3 default max_degree 8), then sample 0 and
    repeatedly the sample furthest from the selected set along the graph.  Returns the selected indices."""
    from .engine import Engine
    e = Engine(device)
    e.build_points(points3)
    idx, sep, _, _ = sample_graph(e, max_degree, seeds)
    return e.graph_fps(idx, sep, int(num_desired))


def sample_evenly_by_separation(eng_mesh, desired, seed, graph_seeds, oversampling_factor=3.0, device=0):
    """MeshSampler::sampleEvenlyBySeparation: area oversampling on the device, furthest point subsampling, gather.
    Returns (positions, face normals, triangle indices) or None when the subsampling comes up short (the reference returns 0)."""
    n_over = int(np.ceil(np.float32(oversampling_factor) * desired))
    p, n, t = eng_mesh.area_sample(n_over, seed)
    sel = furthest_point_subsample(p, desired, graph_seeds, device=device)
    if len(sel) < desired:
        return None
    return p[sel], n[sel], t[sel]


# ---- local surface features: processRange calls with an accumulating functor ------------------------

def t_junction_candidates(eng_boundary_vertices, p0, p1, tol, exclude=None):
    """The query half of MeshFix's T-junction repair (Tools/MeshFix/MeshFix.cpp:716-790) for E boundary edges at once: per edge the
    segment shrunk by tol at both ends (segFromEdge, :699-713; edges shorter than 2 tol are skipped), one closestElement query with
    dist_bound = 2 tol against the structure over the boundary vertices that ignores exclude[e] (VertexFilter, :678-697, as positions in
    the structure, -1 = unused), and the acceptance test of :766-771.  Returns (split_vertex [E] or -1, segment ends a, b)."""
    f = np.float32
    p0 = np.ascontiguousarray(p0, f).reshape(-1, 3)
    p1 = np.ascontiguousarray(p1, f).reshape(-1, 3)
    d = p1 - p0
    ln = np.sqrt(d[:, 0] * d[:, 0] + (d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])).astype(f)        # LineSegmentN::length
    ok = ~(ln.astype(np.float64) < 2 * tol)
    s = np.where(ok, tol / np.where(ok, ln, 1).astype(np.float64), 0.0)
    t = 1 - s
    sf, tf = s.astype(f)[:, None], t.astype(f)[:, None]
    a = ((f(1) - sf) * p0 + sf * p1).astype(f)
    b = ((f(1) - tf) * p0 + tf * p1).astype(f)
    idx, d2, vp, sp = eng_boundary_vertices.closest_segment(a, b, 2 * tol, exclude)

    def sq(x, y):
        e = (x - y).astype(f)
        return (e[:, 0] * e[:, 0] + (e[:, 1] * e[:, 1] + e[:, 2] * e[:, 2])).astype(np.float64)

    sqtol = tol * tol
    take = ok & (idx >= 0) & (sq(sp, p0) > sqtol) & (sq(sp, p1) > sqtol) & (sq(sp, vp) < sqtol)
    return np.where(take, idx, -1).astype(np.int32), a, b


class TheaRandom:
    """Thea::Random (Random.cpp:48-158): MT19937 whose regeneration never touches state[227] (the loop bounds of Random::generate,
    :92-118, reproduced as they are), tempering of bits() (:68-90), integer(lo, hi) = floor(lo + (hi - lo + 1) * bits / 0xFFFFFFFF)
    clamped to hi (:147-157).  The reference's callers share one generator, Random::common(), seeded from the clock; a caller that
    wants reproducible features hands one of these around instead."""
    N, M = 624, 397

    def __init__(self, seed):
        st = np.empty(self.N, np.uint64)
        st[0] = np.uint64(seed & 0xFFFFFFFF)
        for i in range(1, self.N):
            st[i] = (np.uint64(1812433253) * (st[i - 1] ^ (st[i - 1] >> np.uint64(30))) + np.uint64(i)) & np.uint64(0xFFFFFFFF)
        self.st = st.astype(np.uint32)
        self.index = self.N

    def _generate(self):
        st, N, M = self.st, self.N, self.M
        U, L, A = np.uint32(0x80000000), np.uint32(0x7FFFFFFF), np.uint32(0x9908B0DF)

        def step(dst, src):
            x = (st[dst] & U) | (st[dst.start + 1:dst.stop + 1] & L)
            st[dst] = st[src] ^ (x >> np.uint32(1)) ^ np.where(x & np.uint32(1), A, np.uint32(0))
        step(slice(0, N - M), slice(M, N))                                   # i = 0 .. 226: reads old state[i + 1] and old state[i + 397]
        step(slice(N - M + 1, 2 * (N - M) + 1), slice(1, N - M + 1))          # i = 228 .. 454: reads state[1 .. 227] (227 is never regenerated)
        step(slice(2 * (N - M) + 1, N - 1), slice(N - M + 1, M - 1))          # i = 455 .. 622: reads state[228 .. 395], new
        y = (st[N - 1] & U) | (st[0] & L)
        st[N - 1] = st[M - 1] ^ (y >> np.uint32(1)) ^ (A if (y & np.uint32(1)) else np.uint32(0))
        self.index = 0

    def bits(self, n=None):
        """n untempered-then-tempered draws as uint32 (a scalar int when n is None)."""
        count = 1 if n is None else int(n)
        out = np.empty(count, np.uint32)
        k = 0
        while k < count:
            if self.index >= self.N:
                self._generate()
            m = min(count - k, self.N - self.index)
            out[k:k + m] = self.st[self.index:self.index + m]
            self.index += m
            k += m
        r = out
        r = r ^ (r >> np.uint32(11))
        r = r ^ ((r << np.uint32(7)) & np.uint32(0x9D2C5680))
        r = r ^ ((r << np.uint32(15)) & np.uint32(0xEFC60000))
        r = r ^ (r >> np.uint32(18))
        return int(r[0]) if n is None else r

    @staticmethod
    def to_integer(bits, lo, hi):
        """Random::integer(lo, hi) from raw draws (arrays broadcast)."""
        r = np.floor(lo + (hi - lo + 1) * bits.astype(np.float64) / np.float64(0xFFFFFFFF)).astype(np.int64)
        return np.minimum(r, hi)


def random_walks(eng_samples, graph_idx, samples3, p3, num_steps, num_walks, rng):
    """SurfaceFeatures::Local::RandomWalks::compute (Algorithms/SurfaceFeatures/Local/RandomWalks.cpp:30-117) for Q positions in turn:
    the seed sample is the structure's closest element to the position (ONE batched closestElement call for all positions), then
    num_walks walks of num_steps steps over the neighbour rows graph_idx [P x K] (-1 padded; what sample_graph returned), all drawing
    from the shared generator `rng` (TheaRandom) in the reference's order -- position-major, walk-major, step-major --; per step the sum of
    the sample positions reached (float coordinates added into doubles in walk order), divided by the number of walks that got that far,
    less the position.  The walks themselves are the caller's own sequential arithmetic over one random stream, as in the reference; they
    stay on the host.  Returns [Q x num_steps x 3] float64."""
    f = np.float32
    p3 = np.ascontiguousarray(p3, f).reshape(-1, 3)
    pts = np.ascontiguousarray(samples3, f).reshape(-1, 3)
    gi = np.asarray(graph_idx).astype(np.int64)
    Q, S, W, K = len(p3), int(num_steps), int(num_walks), gi.shape[1]
    seed_idx = np.asarray(eng_samples.closest(p3)[0]).astype(np.int64)
    deg = (gi >= 0).sum(1)
    assert np.array_equal(deg, np.where((gi >= 0).all(1), K, np.argmin(gi >= 0, axis=1))), "neighbour rows must be -1 padded at the end"
    out = np.empty((Q, S, 3), np.float64)
    if (deg > 0).all():
        # no walk can stop early: walk w of position q takes draws ((q W + w) S + s), so the walks of one position advance together
        for q in range(Q):
            draws = rng.bits(W * S).reshape(W, S)
            cur = np.full(W, seed_idx[q], np.int64)
            for s_ in range(S):
                nxt = TheaRandom.to_integer(draws[:, s_], 0, deg[cur] - 1)
                cur = gi[cur, nxt]
                out[q, s_] = np.cumsum(pts[cur].astype(np.float64), axis=0)[-1] / W - p3[q].astype(np.float64)      # cumsum: left to right
        return out
    for q in range(Q):                    # isolated samples exist: a walk that reaches one stops and leaves its remaining draws to the next walk
        acc = np.zeros((S, 3), np.float64)
        counts = np.zeros(S, np.int64)
        for _ in range(W):
            cur = seed_idx[q]
            for s_ in range(S):
                if deg[cur] == 0:
                    break
                cur = gi[cur, int(TheaRandom.to_integer(np.array([rng.bits()], np.uint32), 0, int(deg[cur]) - 1)[0])]
                acc[s_] += pts[cur].astype(np.float64)
                counts[s_] += 1
        with np.errstate(invalid="ignore", divide="ignore"):
            out[q] = acc / counts[:, None] - p3[q].astype(np.float64)
    return out


def average_distance(eng, p3, max_distance, stats=None):
    """AverageDistance::compute(p, EUCLIDEAN, max_distance) with max_distance >= 0 (absolute) for every p:
    mean |x - p| over the samples in Ball3(p, max_distance), divided by max_distance (AverageDistance.cpp:63-81)."""
    st = eng.ball_stats(p3, max_distance) if stats is None else stats
    n = st[:, 0]
    avg = np.where(n > 0, st[:, 1] / np.maximum(n, 1.0), 0.0)
    return avg / np.float64(np.float32(max_distance))


def projected_curvature(eng, p3, n3, nbd_radius, stats=None):
    """Curvature::computeProjectedCurvature(p, n, r) for every (p, n); nbd_radius is the ABSOLUTE ball radius (the reference
    multiplies its argument by surf->getScale(), Curvature.cpp:88-91).  -sum(offsets) . n / num_offsets, 0 when none."""
    n3 = np.ascontiguousarray(n3, np.float32).reshape(-1, 3)
    st = eng.ball_stats(p3, nbd_radius, qn3=n3) if stats is None else stats
    k = st[:, 2]
    proj = (st[:, 3:6] * n3.astype(np.float64)).sum(axis=1)
    return np.where(k > 0, -proj / np.maximum(k, 1.0), 0.0)


def pca_from_stats(st, centres):
    """(eigenvalues [Q,3] descending, eigenvectors [Q,3,3] rows = unit vectors, centroid [Q,3], cov [Q,3,3]) from tgq_ball_stats
    sums: covariance about the centroid, divided by n - 1 (PcaN.hpp:163-184); zeros when n < 2 (PcaN.hpp:177-182)."""
    n = st[:, 0]
    m = st[:, 6:9] / np.maximum(n, 1.0)[:, None]
    S2 = np.empty((len(st), 3, 3))
    S2[:, 0, 0], S2[:, 0, 1], S2[:, 0, 2] = st[:, 9], st[:, 10], st[:, 11]
    S2[:, 1, 0], S2[:, 1, 1], S2[:, 1, 2] = st[:, 10], st[:, 12], st[:, 13]
    S2[:, 2, 0], S2[:, 2, 1], S2[:, 2, 2] = st[:, 11], st[:, 13], st[:, 14]
    cov = (S2 - n[:, None, None] * m[:, :, None] * m[:, None, :]) / np.maximum(n - 1.0, 1.0)[:, None, None]
    ok = n >= 2
    cov[~ok] = 0.0
    w, v = np.linalg.eigh(cov)                      # ascending; SelfAdjointEigenSolver sorts the same way (PcaN.hpp:193)
    evals = w[:, ::-1].copy()
    evecs = np.transpose(v, (0, 2, 1))[:, ::-1, :].copy()
    evals[~ok] = 0.0
    evecs[~ok] = 0.0
    centroid = np.asarray(centres, np.float64).reshape(-1, 3) + m
    return evals, evecs, centroid, cov


def local_pca(eng, p3, nbd_radius):
    """LocalPca::compute(p, eigenvectors, r) for every p; nbd_radius absolute (LocalPca.cpp:52-55 multiplies by getScale())."""
    st = eng.ball_stats(p3, nbd_radius)
    evals, evecs, _, _ = pca_from_stats(st, p3)
    return evals, evecs


def visibility(eng, p3, dirs3, scale=None):
    """Visibility::compute for every p with the given ray directions; scale defaults to the diagonal of the structure's bounds
    (Visibility.hpp:73,88)."""
    if scale is None:
        _, _, lo, hi = eng.info()
        ext = (np.asarray(hi, np.float32) - np.asarray(lo, np.float32)).astype(np.float32)
        scale = float(np.sqrt(np.float32(ext[0] * ext[0]) + (np.float32(ext[1] * ext[1]) + np.float32(ext[2] * ext[2]))))
    vis, _ = eng.visibility(p3, dirs3, scale)
    return vis


# ---- PointSet3 on-disk formats: plain text, one line per sample ---------------------------

def _fmt(x):
    """operator<< of a float with the default ostream precision (6 significant digits, %g)."""
    return "%g" % float(x)


def save_graph(path, max_degree, idx, sep=None, count=None, write_distances=False):
    """PointSet3::save, graph part (PointSet3.cpp:331-354): first line max_degree, then per sample
    `degree nbr_0 ... nbr_{d-1} [sep_0 ... sep_{d-1}]`."""
    idx = np.asarray(idx)
    if count is None:
        count = (idx >= 0).sum(axis=1)
    with open(path, "wb") as f:
        f.write(b"%d\n" % int(max_degree))
        for i in range(idx.shape[0]):
            d = int(count[i])
            parts = [str(d)] + [str(int(j)) for j in idx[i, :d]]
            if write_distances:
                parts += [_fmt(x) for x in sep[i, :d]]
            f.write((" ".join(parts) + "\n").encode())


def load_graph(path, positions=None):
    """PointSet3::load, graph part (PointSet3.cpp:222-321): returns (max_degree, idx [P x max_degree] int32 padded with -1,
    sep [P x max_degree] float32 padded with -1, count [P]).  Without stored separations they are the Euclidean distances of
    `positions` (PointSet3.cpp:304-311); neighbours are re-sorted by (separation, index) as BoundedSortedArray::insert does."""
    with open(path, "r") as f:
        lines = f.read().split("\n")
    try:
        max_degree = int(lines[0].split()[0])
    except (IndexError, ValueError):
        raise ValueError(f"PointSet3: Could not read valid maximum degree from '{path}'")
    if max_degree < 0:
        raise ValueError(f"PointSet3: Could not read valid maximum degree from '{path}'")
    rows = [ln for ln in lines[1:] if ln.strip() != ""]
    P = len(rows) if positions is None else len(positions)
    if len(rows) < P:
        raise ValueError(f"PointSet3: Could not read neighbors of vertex {len(rows)} from '{path}'")
    idx = np.full((P, max_degree), -1, np.int32)
    sep = np.full((P, max_degree), -1.0, np.float32)
    count = np.zeros(P, np.int32)
    pos = None if positions is None else np.asarray(positions, np.float32).reshape(-1, 3)
    for i in range(P):
        tok = rows[i].split()
        d = int(tok[0])
        if d < 0:
            raise ValueError(f"PointSet3: Could not read valid degree of vertex {i} from '{path}'")
        nb = np.array([int(t) for t in tok[1:1 + d]], np.int64)
        if len(nb) != d or (nb < 0).any() or (nb >= P).any():
            raise ValueError(f"PointSet3: Could not read valid neighbor of vertex {i} from '{path}'")
        rest = tok[1 + d:]
        if rest:
            if len(rest) < d:
                raise ValueError(f"PointSet3: Could not read separation of neighbor {len(rest)} of vertex {i} from '{path}'")
            sp = np.array([float(t) for t in rest[:d]], np.float32)
        elif pos is not None:
            df = pos[nb] - pos[i]
            sp = np.sqrt(df[:, 0] * df[:, 0] + (df[:, 1] * df[:, 1] + df[:, 2] * df[:, 2])).astype(np.float32)
        else:
            sp = np.full(d, -1.0, np.float32)
        order = np.lexsort((nb, sp))[:max_degree]
        k = len(order)
        idx[i, :k], sep[i, :k], count[i] = nb[order], sp[order], k
    return max_degree, idx, sep, count


def save_samples(path, p3, n3=None):
    """PointSet3::save, samples part (PointSet3.cpp:356-372): `x y z [nx ny nz]` per line."""
    p3 = np.asarray(p3, np.float32).reshape(-1, 3)
    with open(path, "wb") as f:
        for i in range(len(p3)):
            vals = list(p3[i]) + ([] if n3 is None else list(np.asarray(n3, np.float32).reshape(-1, 3)[i]))
            f.write((" ".join(_fmt(v) for v in vals) + "\n").encode())


def load_samples(path):
    """PointSet3::load, samples part (PointSet3.cpp:186-216): (positions, normals or None); mixed lines are an error."""
    P, N = [], []
    has_n = None
    with open(path, "r") as f:
        for ln in f.read().split("\n"):
            tok = ln.split()
            if not tok:
                continue
            if len(tok) < 3:
                raise ValueError(f"Could not read sample {len(P)} from '{path}'")
            P.append([float(t) for t in tok[:3]])
            hn = len(tok) >= 6
            if has_n is None:
                has_n = hn
            elif has_n != hn:
                raise ValueError(f"PointSet3: Some samples in '{path}' have normals and some don't")
            if hn:
                N.append([float(t) for t in tok[3:6]])
    p = np.asarray(P, np.float32).reshape(-1, 3)
    return p, (np.asarray(N, np.float32).reshape(-1, 3) if has_n else None)
