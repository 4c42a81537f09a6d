"""ctypes binding of oracle/_ref/libthea_ref.so: the reference's OWN sources (sidch/Thea, /root/reference/Code/Source) compiled
unmodified against the Eigen stand-in of oracle/eigen_shim/ (recipe: oracle/Makefile target `ref`).

TEST INFRASTRUCTURE ONLY -- used by tests/ to pin the restatement in oracle/thea_oracle.hpp to the reference's code, and by
bench.py's CPU legs.  The product package (thea_b200) never imports this module.  The library can only be BUILT where
/root/reference exists; elsewhere (the GPU box) the prebuilt, git-ignored file is used if it travelled with the snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libthea_ref.so")
_LIB = None


def available(build: bool = True) -> bool:
    """True when the reference library exists (building it first when the reference sources are present)."""
    if build and os.path.exists("/root/reference/Code/Source/Algorithms/BvhN.hpp"):
        try:
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
        except Exception:
            return False
    return os.path.exists(_SO)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise RuntimeError("oracle/_ref/libthea_ref.so is missing and /root/reference is not present to build it")
        L = C.CDLL(_SO)
        vp, i64, f32, f64, i32 = C.c_void_p, C.c_int64, C.c_float, C.c_double, C.c_int
        L.ref_create_tris.restype = vp
        L.ref_create_tris.argtypes = [vp, i64]
        L.ref_create_points.restype = vp
        L.ref_create_points.argtypes = [vp, i64]
        L.ref_destroy.argtypes = [vp]
        L.ref_num_nodes.restype = i64
        L.ref_num_nodes.argtypes = [vp]
        L.ref_closest.argtypes = [vp, vp, i64, f64, i32, vp, vp, vp]
        L.ref_closest_element.argtypes = [vp, vp, i64, f64, vp, vp, vp]
        L.ref_knn.argtypes = [vp, vp, i64, i32, f64, i32, vp, vp]
        L.ref_ray_first.argtypes = [vp, vp, vp, vp, i64, i32, vp, vp, vp]
        L.ref_ray_time.argtypes = [vp, vp, vp, vp, i64, vp, vp]
        L.ref_ball_range.restype = i64
        L.ref_ball_range.argtypes = [vp, vp, f32, i64, vp]
        L.ref_box_range.restype = i64
        L.ref_box_range.argtypes = [vp, vp, vp, i64, vp]
        L.ref_closest_pair_bvh.restype = i32
        L.ref_closest_pair_bvh.argtypes = [vp, vp, f64, vp, vp, vp, vp, vp]
        L.ref_closest_segment.restype = i64
        L.ref_closest_segment.argtypes = [vp, vp, vp, f64, vp, vp, vp]
        L.ref_closest_segment_batch.argtypes = [vp, vp, vp, i64, f64, vp, i32, vp, vp, vp, vp]
        L.ref_seg_point_batch.argtypes = [vp, vp, vp, i64, vp, vp]
        L.ref_tri_seg_batch.argtypes = [vp, vp, vp, i64, vp, vp, vp]
        L.ref_tri_tri_batch.argtypes = [vp, vp, i64, vp, vp, vp]
        L.ref_closest_triangle_batch.argtypes = [vp, vp, i64, f64, vp, vp, vp, vp]
        L.ref_random_walks.argtypes = [vp, i64, i32, vp, vp, vp, i64, i64, i64, C.c_uint32, vp]
        L.ref_tri_cache.argtypes = [vp] * 6
        L.ref_tri_point.restype = f32
        L.ref_tri_point.argtypes = [vp, vp, vp]
        L.ref_tri_point_batch.argtypes = [vp, vp, i64, vp, vp]
        L.ref_ray_tri.restype = f32
        L.ref_ray_tri.argtypes = [vp, vp, vp, f32]
        L.ref_ray_tri_batch.argtypes = [vp, vp, vp, i64, vp]
        L.ref_fast_sqrt.restype = f32
        L.ref_fast_sqrt.argtypes = [f32]
        L.ref_box_point_sqdist.restype = f32
        L.ref_box_point_sqdist.argtypes = [vp, vp, vp]
        L.ref_box_ray_time.restype = f32
        L.ref_box_ray_time.argtypes = [vp, vp, vp, vp, f32]
        L.ref_random_uniform01.argtypes = [C.c_uint32, i64, vp]
        L.ref_tree_depth.restype = i32
        L.ref_tree_depth.argtypes = [i64, i32]
        L.ref_icp.argtypes = [vp, vp, i64, i32, i32, f64, vp, vp]
        L.ref_icp_error.restype = f32
        L.ref_icp_error.argtypes = [vp, i64, vp, vp]
        L.ref_icp_step.argtypes = [i64, vp, vp, vp]
        L.ref_sdf.argtypes = [vp, vp, vp, i64, f32, i32, i32, vp]
        L.ref_graph.argtypes = [vp, vp, i64, i32, f32, vp, vp, vp, vp]
        L.ref_avg_separation.restype = f32
        L.ref_avg_separation.argtypes = [vp, i64, vp, i32]
        L.ref_k_closest_pairs_bvh.restype = i32
        L.ref_k_closest_pairs_bvh.argtypes = [vp, vp, i32, f64, vp, vp, vp]
        L.ref_graph_contract.argtypes = [vp, i64, i64, i32, vp, vp, vp, vp, vp]
        L.ref_geodesic_features.argtypes = [vp, i64, i32, vp, vp, f32, vp, i64, f32, i32, vp, vp]
        L.ref_fps.restype = i64
        L.ref_fps.argtypes = [vp, i64, i32, vp, vp, i64, vp]
        L.ref_spin_image.argtypes = [vp, vp, i64, f32, vp, vp, i64, i32, i32, vp]
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def ncores() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


class Reference:
    """Thea's BvhN (kd-tree, default parameters) over triangles (T, 9) or points (P, 3)."""

    def __init__(self, tris9=None, points=None):
        L = lib()
        if tris9 is not None:
            t = _f32(tris9).reshape(-1, 9)
            self.n, self.is_points = t.shape[0], False
            self.h = L.ref_create_tris(_p(t), self.n)
        else:
            p = _f32(points).reshape(-1, 3)
            self.n, self.is_points = p.shape[0], True
            self.h = L.ref_create_points(_p(p), self.n)

    def __del__(self):
        try:
            if self.h:
                lib().ref_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def num_nodes(self):
        return int(lib().ref_num_nodes(self.h))

    def closest(self, q, dist_bound=-1.0, threads=None):
        """closestPair<MetricL2>(q, dist_bound, UniversalCompatibility(), true): index, squared distance, closest point."""
        q = _f32(q).reshape(-1, 3)
        Q = q.shape[0]
        idx, d2, cp = np.empty(Q, np.int64), np.empty(Q, np.float32), np.empty((Q, 3), np.float32)
        lib().ref_closest(self.h, _p(q), Q, float(dist_bound), threads or ncores(), _p(idx), _p(d2), _p(cp))
        return idx, d2, cp

    def closest_element(self, q, dist_bound=-1.0):
        q = _f32(q).reshape(-1, 3)
        Q = q.shape[0]
        idx, dist, cp = np.empty(Q, np.int64), np.empty(Q, np.float64), np.empty((Q, 3), np.float32)
        lib().ref_closest_element(self.h, _p(q), Q, float(dist_bound), _p(idx), _p(dist), _p(cp))
        return idx, dist, cp

    def knn(self, q, k, dist_bound=-1.0, threads=None):
        q = _f32(q).reshape(-1, 3)
        Q = q.shape[0]
        idx, d2 = np.empty((Q, k), np.int64), np.empty((Q, k), np.float32)
        lib().ref_knn(self.h, _p(q), Q, int(k), float(dist_bound), threads or ncores(), _p(idx), _p(d2))
        return idx, d2

    def ray_first(self, o, d, tmax=None, threads=None, normals=False):
        o, d = _f32(o).reshape(-1, 3), _f32(d).reshape(-1, 3)
        tm = None if tmax is None else _f32(tmax)
        R = o.shape[0]
        idx, t = np.empty(R, np.int64), np.empty(R, np.float32)
        n3 = np.empty((R, 3), np.float32) if normals else None
        lib().ref_ray_first(self.h, _p(o), _p(d), _p(tm), R, threads or ncores(), _p(idx), _p(t), _p(n3))
        return (idx, t, n3) if normals else (idx, t)

    def ray_time(self, o, d, tmax=None):
        o, d = _f32(o).reshape(-1, 3), _f32(d).reshape(-1, 3)
        tm = None if tmax is None else _f32(tmax)
        R = o.shape[0]
        t, hit = np.empty(R, np.float32), np.empty(R, np.uint8)
        lib().ref_ray_time(self.h, _p(o), _p(d), _p(tm), R, _p(t), _p(hit))
        return t, hit

    def ball_range(self, c, radius, cap=1 << 20):
        idx = np.empty(cap, np.int64)
        n = lib().ref_ball_range(self.h, _p(_f32(c).reshape(3)), float(radius), cap, _p(idx))
        return idx[:min(n, cap)].copy(), int(n)

    def box_range(self, lo, hi, cap=1 << 20):
        idx = np.empty(cap, np.int64)
        n = lib().ref_box_range(self.h, _p(_f32(lo).reshape(3)), _p(_f32(hi).reshape(3)), cap, _p(idx))
        return idx[:min(n, cap)].copy(), int(n)

    def k_closest_pairs(self, other, k, dist_bound=-1.0):
        """kClosestPairs<MetricL2>(other structure, set of k): (target indices, query indices, squared distances), ascending."""
        ti, qi, d2 = np.empty(k, np.int64), np.empty(k, np.int64), np.empty(k, np.float32)
        n = lib().ref_k_closest_pairs_bvh(self.h, other.h, int(k), float(dist_bound), _p(ti), _p(qi), _p(d2))
        return ti[:n], qi[:n], d2[:n]

    def closest_pair(self, other, dist_bound=-1.0):
        """BVH-to-BVH closestPair: (found, target index in self, query index in other, d2, query point, target point)."""
        ti, qi = np.zeros(1, np.int64), np.zeros(1, np.int64)
        d2, a, b = np.zeros(1, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
        ok = lib().ref_closest_pair_bvh(self.h, other.h, float(dist_bound), _p(ti), _p(qi), _p(d2), _p(a), _p(b))
        return bool(ok), int(ti[0]), int(qi[0]), float(d2[0]), a, b

    def closest_segment(self, a, b, dist_bound=-1.0):
        d2, ca, cb = np.zeros(1, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
        i = lib().ref_closest_segment(self.h, _p(_f32(a).reshape(3)), _p(_f32(b).reshape(3)), float(dist_bound), _p(d2), _p(ca), _p(cb))
        return int(i), float(d2[0]), ca, cb

    def closest_segments(self, a, b, dist_bound=-1.0, exclude=None):
        """BvhN::closestPair<MetricL2>(LineSegment3(a_i, b_i), dist_bound, ..., true) per query, under a pushed index filter:
        idx, d2, point on the element, point on the segment."""
        a = _f32(a).reshape(-1, 3)
        b = _f32(b).reshape(-1, 3)
        Q = a.shape[0]
        ex = None if exclude is None else np.ascontiguousarray(exclude, np.int64).reshape(Q, -1)
        idx, d2 = np.empty(Q, np.int64), np.empty(Q, np.float32)
        ep, sp = np.empty((Q, 3), np.float32), np.empty((Q, 3), np.float32)
        lib().ref_closest_segment_batch(self.h, _p(a), _p(b), Q, float(dist_bound), _p(ex), 0 if ex is None else ex.shape[1],
                                        _p(idx), _p(d2), _p(ep), _p(sp))
        return idx, d2, ep, sp

    def closest_triangles(self, q9, dist_bound=-1.0):
        """BvhN::closestPair<MetricL2>(Triangle3(q_i), dist_bound, ..., true) per query: idx, d2, point on the element, point on the query."""
        q = _f32(q9).reshape(-1, 9)
        Q = q.shape[0]
        idx, d2 = np.empty(Q, np.int64), np.empty(Q, np.float32)
        ep, qp = np.empty((Q, 3), np.float32), np.empty((Q, 3), np.float32)
        lib().ref_closest_triangle_batch(self.h, _p(q), Q, float(dist_bound), _p(idx), _p(d2), _p(ep), _p(qp))
        return idx, d2, ep, qp

    def icp(self, src, min_iters=3, max_iters=10, frac_thresh=1e-4):
        s = _f32(src).reshape(-1, 3)
        tr, err = np.zeros((3, 4), np.float32), np.zeros(1, np.float32)
        lib().ref_icp(self.h, _p(s), s.shape[0], int(min_iters), int(max_iters), float(frac_thresh), _p(tr), _p(err))
        return tr, float(err[0])

    def sdf(self, p, n, scale, only_interior=True, threads=None):
        p = _f32(p).reshape(-1, 3)
        n = None if n is None else _f32(n).reshape(-1, 3)
        out = np.empty(p.shape[0], np.float64)
        lib().ref_sdf(self.h, _p(p), _p(n), p.shape[0], float(scale), int(only_interior), threads or ncores(), _p(out))
        return out


def graph(points, normals, max_degree, init_radius, surface: Reference | None = None):
    """PointSet3::findSampleNeighbors for every sample: (idx [P, k], sep [P, k], count [P])."""
    p = _f32(points).reshape(-1, 3)
    n = None if normals is None else _f32(normals).reshape(-1, 3)
    P = p.shape[0]
    idx, sep, cnt = np.empty((P, max_degree), np.int64), np.empty((P, max_degree), np.float32), np.empty(P, np.int32)
    lib().ref_graph(_p(p), _p(n), P, int(max_degree), float(init_radius), surface.h if surface is not None else None, _p(idx), _p(sep), _p(cnt))
    return idx, sep, cnt


def avg_separation(points, seeds):
    p = _f32(points).reshape(-1, 3)
    s = np.ascontiguousarray(seeds, np.int64)
    return float(lib().ref_avg_separation(_p(p), p.shape[0], _p(s), len(s)))


def icp_error(tr12, src, dst):
    s, d, t = _f32(src).reshape(-1, 3), _f32(dst).reshape(-1, 3), _f32(tr12).reshape(12)
    return float(lib().ref_icp_error(_p(t), s.shape[0], _p(s), _p(d)))


def icp_step(src, dst):
    s, d = _f32(src).reshape(-1, 3), _f32(dst).reshape(-1, 3)
    tr = np.zeros((3, 4), np.float32)
    lib().ref_icp_step(s.shape[0], _p(s), _p(d), _p(tr))
    return tr


def tri_point_batch(v9, p3):
    v, p = _f32(v9).reshape(-1, 9), _f32(p3).reshape(-1, 3)
    d2, cp = np.empty(len(v), np.float32), np.empty((len(v), 3), np.float32)
    lib().ref_tri_point_batch(_p(v), _p(p), len(v), _p(d2), _p(cp))
    return d2, cp


def ray_tri_batch(v9, o3, d3):
    v, o, d = _f32(v9).reshape(-1, 9), _f32(o3).reshape(-1, 3), _f32(d3).reshape(-1, 3)
    t = np.empty(len(v), np.float32)
    lib().ref_ray_tri_batch(_p(v), _p(o), _p(d), len(v), _p(t))
    return t


def seg_point_batch(a3, b3, p3):
    a, b, p = (_f32(x).reshape(-1, 3) for x in (a3, b3, p3))
    n = a.shape[0]
    d2, sp = np.empty(n, np.float32), np.empty((n, 3), np.float32)
    lib().ref_seg_point_batch(_p(a), _p(b), _p(p), n, _p(d2), _p(sp))
    return d2, sp


def tri_seg_batch(v9, a3, b3):
    v = _f32(v9).reshape(-1, 9)
    a, b = (_f32(x).reshape(-1, 3) for x in (a3, b3))
    n = a.shape[0]
    d2, tp, sp = np.empty(n, np.float32), np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
    lib().ref_tri_seg_batch(_p(v), _p(a), _p(b), n, _p(d2), _p(tp), _p(sp))
    return d2, tp, sp


def tri_tri_batch(va9, vb9):
    """Triangle3 a_i.squaredDistance(b_i, &pa, &pb): squared distance, point on a, point on b."""
    a, b = (_f32(x).reshape(-1, 9) for x in (va9, vb9))
    n = a.shape[0]
    d2, pa, pb = np.empty(n, np.float32), np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
    lib().ref_tri_tri_batch(_p(a), _p(b), n, _p(d2), _p(pa), _p(pb))
    return d2, pa, pb


def tri_cache(v9):
    n, d, ax, ar, c = np.empty(3, np.float32), np.empty(1, np.float32), np.empty(1, np.int32), np.empty(1, np.float32), np.empty(3, np.float32)
    lib().ref_tri_cache(_p(_f32(v9).reshape(9)), _p(n), _p(d), _p(ax), _p(ar), _p(c))
    return n, float(d[0]), int(ax[0]), float(ar[0]), c


def fast_sqrt(x):
    return float(lib().ref_fast_sqrt(float(x)))


def box_point_sqdist(lo, hi, p):
    return float(lib().ref_box_point_sqdist(_p(_f32(lo)), _p(_f32(hi)), _p(_f32(p))))


def box_ray_time(lo, hi, o, d, max_time=-1.0):
    return float(lib().ref_box_ray_time(_p(_f32(lo)), _p(_f32(hi)), _p(_f32(o)), _p(_f32(d)), float(max_time)))


def random_uniform01(seed, n):
    out = np.empty(n, np.float32)
    lib().ref_random_uniform01(seed, n, _p(out))
    return out


def tree_depth(n, leaf):
    return int(lib().ref_tree_depth(n, leaf))


# ---- sample-graph consumers and SpinImage: the reference's own PointSet3 / ShortestPaths / SurfaceFeatures::Local code on a GIVEN graph ----
def _graph_args(idx, sep):
    idx = np.ascontiguousarray(idx, np.int64)
    sep = np.ascontiguousarray(sep, np.float32)
    return idx, sep, idx.shape[0], idx.shape[1]


def graph_contract(points, n_orig, idx, sep):
    """PointSet3::extractOriginalAdjacencies on a graph over n_orig original + the remaining dense samples: (idx, sep, count) of the originals."""
    p = _f32(points).reshape(-1, 3)
    idx, sep, P, K = _graph_args(idx, sep)
    oi, os_, oc = np.empty((n_orig, K), np.int64), np.empty((n_orig, K), np.float32), np.empty(n_orig, np.int32)
    lib().ref_graph_contract(_p(p), int(n_orig), int(P - n_orig), K, _p(idx), _p(sep), _p(oi), _p(os_), _p(oc))
    return oi, os_, oc


def geodesic_features(points, idx, sep, scale, q, max_distance, nbins):
    """AverageDistance / LocalDistanceHistogram ::compute(q, GEODESIC, max_distance): (avg [Q] float64, hist [Q x nbins] float64)."""
    p = _f32(points).reshape(-1, 3)
    q = _f32(q).reshape(-1, 3)
    idx, sep, P, K = _graph_args(idx, sep)
    avg, hist = np.empty(len(q), np.float64), np.empty((len(q), nbins), np.float64)
    lib().ref_geodesic_features(_p(p), P, K, _p(idx), _p(sep), float(scale), _p(q), len(q), float(max_distance), int(nbins), _p(avg), _p(hist))
    return avg, hist


def fps(points, idx, sep, n_desired):
    """FurthestPointSampling::subsample's selection loop around the reference's ShortestPaths, on a given graph."""
    p = _f32(points).reshape(-1, 3)
    idx, sep, P, K = _graph_args(idx, sep)
    sel = np.full(int(n_desired), -1, np.int64)
    n = lib().ref_fps(_p(p), P, K, _p(idx), _p(sep), int(n_desired), _p(sel))
    return sel[:n]


def random_walks(points, idx, sep, q, num_steps, num_walks, seed):
    """The reference's RandomWalks::compute per position on an injected graph, Random::common() put into the state of Random(seed)."""
    p = _f32(points).reshape(-1, 3)
    gi, gs, _, _ = _graph_args(idx, sep)
    q = _f32(q).reshape(-1, 3)
    out = np.empty((q.shape[0], int(num_steps), 3), np.float64)
    lib().ref_random_walks(_p(p), p.shape[0], gi.shape[1], _p(gi), _p(gs), _p(q), q.shape[0], int(num_steps), int(num_walks), int(seed), _p(out))
    return out


def spin_image(points, normals, scale, q, qn, nr, nh):
    """SpinImage::compute(position, normal, nr, nh): [Q x nr x nh] float64."""
    p, n, q, qn = _f32(points).reshape(-1, 3), _f32(normals).reshape(-1, 3), _f32(q).reshape(-1, 3), _f32(qn).reshape(-1, 3)
    out = np.empty((len(q), int(nr), int(nh)), np.float64)
    lib().ref_spin_image(_p(p), _p(n), len(p), float(scale), _p(q), _p(qn), len(q), int(nr), int(nh), _p(out))
    return out
