"""ctypes binding of the CPU oracle (oracle/liborc.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (thea_b200) never imports this module.  Pinned to the reference's
own code by tests/test_ref_pin.py (oracle/ref.py, oracle/_ref); what remains assumed is Eigen's evaluation order
(see oracle/thea_oracle.hpp).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODE_REFERENCE = 0   # kd-tree of BvhN.hpp:1450-1525, DFS order, first-found wins (the reference's exact behaviour)
MODE_CANON = 1       # canonical tie rule (lowest element index) through the kd-tree
MODE_BRUTE = 2       # canonical rule by exhaustive scan (the definition)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "thea_oracle.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.orc_create_tris.restype = C.c_void_p
        L.orc_create_tris.argtypes = [C.c_void_p, C.c_int64]
        L.orc_create_points.restype = C.c_void_p
        L.orc_create_points.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_num_nodes.restype = C.c_int64
        L.orc_num_nodes.argtypes = [C.c_void_p]
        L.orc_max_depth.restype = C.c_int64
        L.orc_max_depth.argtypes = [C.c_void_p]
        L.orc_tri_cache.argtypes = [C.c_void_p] * 5
        L.orc_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_ray_first.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p]
        L.orc_knn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                              C.c_void_p, C.c_void_p]
        L.orc_closest_segment.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_closest_triangle.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4
        L.orc_tri_tri_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_random_walks.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_uint32, C.c_void_p]
        L.orc_seg_point_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_tri_seg_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_labels.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_closest_filtered.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_double, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_knn_filtered.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_double, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p]
        L.orc_ball_range.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_ball_graph_surface.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_ball_hist.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_graph_contract.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_graph_geodesic.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_graph_fps.restype = C.c_int64
        L.orc_graph_fps.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p]
        L.orc_spin_image.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]
        L.orc_ball_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]
        L.orc_visibility.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_avg_separation.restype = C.c_float
        L.orc_avg_separation.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_ball_graph.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_sdf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_int, C.c_int, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_sdf_stats.restype = C.c_double
        L.orc_sdf_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        L.orc_icp.restype = C.c_int
        L.orc_icp.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_icp_ex.restype = C.c_int
        L.orc_icp_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_icp_error.restype = C.c_float
        L.orc_icp_error.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_icp_step.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_area_sample.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_tri_point.restype = C.c_float
        L.orc_tri_point.argtypes = [C.c_void_p] * 3
        L.orc_ray_tri.restype = C.c_float
        L.orc_ray_tri.argtypes = [C.c_void_p] * 3
        L.orc_fast_sqrt.restype = C.c_float
        L.orc_fast_sqrt.argtypes = [C.c_float]
        L.orc_box_ray_time.restype = C.c_float
        L.orc_box_ray_time.argtypes = [C.c_void_p] * 4 + [C.c_float]
        L.orc_polar_rotation.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_thea_random_uniform01.argtypes = [C.c_uint32, C.c_int64, C.c_void_p]
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def ncores() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


class Oracle:
    """CPU BvhN restatement over triangles (T,9) or points (P,3) [+ normals]."""

    def __init__(self, tris9=None, points=None, normals=None):
        L = lib()
        self._keep = []
        if tris9 is not None:
            t = _f32(tris9).reshape(-1, 9)
            self.n = t.shape[0]
            self.is_points = False
            self.h = L.orc_create_tris(_p(t), self.n)
        else:
            p = _f32(points).reshape(-1, 3)
            nn = None if normals is None else _f32(normals).reshape(-1, 3)
            self.n = p.shape[0]
            self.is_points = True
            self.h = L.orc_create_points(_p(p), self.n, _p(nn))

    def __del__(self):
        try:
            if self.h:
                lib().orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def tri_cache(self):
        n = np.empty((self.n, 3), np.float32)
        d = np.empty(self.n, np.float32)
        ax = np.empty(self.n, np.int32)
        ar = np.empty(self.n, np.float32)
        lib().orc_tri_cache(self.h, _p(n), _p(d), _p(ax), _p(ar))
        return n, d, ax, ar

    def closest(self, q, dist_bound=-1.0, mode=MODE_CANON, exclude_self=False, threads=None):
        q = _f32(q).reshape(-1, 3)
        Q = q.shape[0]
        idx = np.empty(Q, np.int64)
        d2 = np.empty(Q, np.float32)
        cp = np.empty((Q, 3), np.float32)
        lib().orc_closest(self.h, _p(q), Q, float(dist_bound), mode, int(exclude_self), threads or ncores(),
                          _p(idx), _p(d2), _p(cp))
        return idx, d2, cp

    def closest_segment(self, a, b, dist_bound=-1.0, exclude=None, mode=MODE_CANON, threads=None):
        """Closest element to LineSegment3(a_i, b_i): idx, d2, point on the element, point on the segment."""
        a = _f32(a).reshape(-1, 3)
        b = _f32(b).reshape(-1, 3)
        Q = a.shape[0]
        ex = None if exclude is None else np.ascontiguousarray(exclude, np.int64).reshape(Q, -1)
        idx = np.empty(Q, np.int64)
        d2 = np.empty(Q, np.float32)
        ep = np.empty((Q, 3), np.float32)
        sp = np.empty((Q, 3), np.float32)
        lib().orc_closest_segment(self.h, _p(a), _p(b), Q, float(dist_bound), _p(ex), 0 if ex is None else ex.shape[1], mode,
                                  threads or ncores(), _p(idx), _p(d2), _p(ep), _p(sp))
        return idx, d2, ep, sp

    def closest_triangle(self, q9, dist_bound=-1.0, mode=MODE_CANON, plain_bound=False, threads=None):
        """Closest element to Triangle3(q_i): idx, d2, point on the element, point on the query triangle."""
        q = _f32(q9).reshape(-1, 9)
        Q = q.shape[0]
        idx = np.empty(Q, np.int64)
        d2 = np.empty(Q, np.float32)
        ep = np.empty((Q, 3), np.float32)
        qp = np.empty((Q, 3), np.float32)
        lib().orc_closest_triangle(self.h, _p(q), Q, float(dist_bound), int(plain_bound), mode, threads or ncores(), _p(idx), _p(d2), _p(ep), _p(qp))
        return idx, d2, ep, qp

    def random_walks(self, graph_idx, q, num_steps, num_walks, seed):
        """RandomWalks::compute for every position of q in turn on the neighbour rows graph_idx [P x K], one shared Random(seed)."""
        gi = np.ascontiguousarray(graph_idx, np.int64)
        q = _f32(q).reshape(-1, 3)
        out = np.empty((q.shape[0], int(num_steps), 3), np.float64)
        lib().orc_random_walks(self.h, _p(gi), gi.shape[0], gi.shape[1], _p(q), q.shape[0], int(num_steps), int(num_walks), int(seed), _p(out))
        return out

    def ray_first(self, o, d, tmax=None, mode=MODE_CANON, threads=None):
        o = _f32(o).reshape(-1, 3)
        d = _f32(d).reshape(-1, 3)
        tm = None if tmax is None else _f32(tmax)
        R = o.shape[0]
        idx = np.empty(R, np.int64)
        t = np.empty(R, np.float32)
        lib().orc_ray_first(self.h, _p(o), _p(d), _p(tm), R, mode, threads or ncores(), _p(idx), _p(t))
        return idx, t

    def knn(self, q, k, dist_bound=-1.0, exclude_self=False, mode=MODE_CANON, threads=None):
        q = _f32(q).reshape(-1, 3)
        Q = q.shape[0]
        idx = np.empty((Q, k), np.int64)
        d2 = np.empty((Q, k), np.float32)
        lib().orc_knn(self.h, _p(q), Q, k, float(dist_bound), int(exclude_self), mode, threads or ncores(), _p(idx), _p(d2))
        return idx, d2

    def set_labels(self, labels):
        """Per-element labels of a point structure (for filtered queries)."""
        self._labels = None if labels is None else np.ascontiguousarray(labels, np.int32)
        lib().orc_set_labels(self.h, _p(self._labels))

    def closest_filtered(self, q, qn=None, qlabel=None, min_dot=0.0, dist_bound=-1.0, mode=MODE_CANON, threads=None):
        q = _f32(q).reshape(-1, 3)
        qn = None if qn is None else _f32(qn).reshape(-1, 3)
        ql = None if qlabel is None else np.ascontiguousarray(qlabel, np.int32)
        Q = q.shape[0]
        idx = np.empty(Q, np.int64)
        d2 = np.empty(Q, np.float32)
        cp = np.empty((Q, 3), np.float32)
        lib().orc_closest_filtered(self.h, _p(q), _p(qn), _p(ql), Q, float(min_dot), float(dist_bound), mode, threads or ncores(), _p(idx), _p(d2), _p(cp))
        return idx, d2, cp

    def knn_filtered(self, q, k, qn=None, qlabel=None, min_dot=0.0, dist_bound=-1.0, mode=MODE_CANON, threads=None):
        q = _f32(q).reshape(-1, 3)
        qn = None if qn is None else _f32(qn).reshape(-1, 3)
        ql = None if qlabel is None else np.ascontiguousarray(qlabel, np.int32)
        Q = q.shape[0]
        idx = np.empty((Q, k), np.int64)
        d2 = np.empty((Q, k), np.float32)
        lib().orc_knn_filtered(self.h, _p(q), _p(qn), _p(ql), Q, int(k), float(min_dot), float(dist_bound), mode, threads or ncores(), _p(idx), _p(d2))
        return idx, d2

    def ball_range(self, c, radius, max_out=32, exclude_self=False, threads=None):
        c = _f32(c).reshape(-1, 3)
        Q = c.shape[0]
        rad = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, np.float32), (Q,)), np.float32)
        idx = np.empty((Q, max_out), np.int64)
        cnt = np.empty(Q, np.int32)
        lib().orc_ball_range(self.h, _p(c), _p(rad), Q, int(max_out), int(exclude_self), threads or ncores(), _p(idx), _p(cnt))
        return idx, cnt

    def ball_graph_surface(self, surface, max_degree, init_radius, mode=MODE_REFERENCE, threads=None):
        """PointSet3::updateGraph(surface): neighbour search with the reachability ray test against the triangle oracle `surface`."""
        idx = np.empty((self.n, max_degree), np.int64)
        sep = np.empty((self.n, max_degree), np.float32)
        cnt = np.empty(self.n, np.int32)
        lib().orc_ball_graph_surface(self.h, surface.h, max_degree, float(init_radius), mode, threads or ncores(), _p(idx), _p(sep), _p(cnt))
        return idx, sep, cnt

    def spin_image(self, q, qn, nr, nh, max_extent, threads=None):
        """SpinImage::compute(position, normal, nr, nh) for every query: [Q x nr x nh] float64 counts (SpinImage.cpp:47-75)."""
        q = np.ascontiguousarray(q, np.float32); qn = np.ascontiguousarray(qn, np.float32)
        Q = len(q)
        out = np.empty((Q, int(nr), int(nh)), np.float64)
        lib().orc_spin_image(self.h, _p(q), _p(qn), Q, int(nr), int(nh), float(max_extent), threads or ncores(), _p(out))
        return out

    def ball_hist(self, c, radius, nbins, mode=MODE_REFERENCE, threads=None):
        """LocalDistanceHistogram (EUCLIDEAN, explicit max_distance): [Q, nbins] float64 bin contents."""
        c = _f32(c).reshape(-1, 3)
        Q = c.shape[0]
        rad = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, np.float32), (Q,)), np.float32)
        out = np.empty((Q, nbins), np.float64)
        lib().orc_ball_hist(self.h, _p(c), _p(rad), Q, int(nbins), mode, threads or ncores(), _p(out))
        return out

    def ball_stats(self, c, radius, qn=None, mode=MODE_REFERENCE, threads=None):
        """Local-feature functors over ball neighbourhoods (see orc_ball_stats): dict of n, avg_dist, num_offsets, curvature,
        centroid [Q,3], cov [Q,6] (xx xy xz yy yz zz)."""
        c = _f32(c).reshape(-1, 3)
        Q = c.shape[0]
        rad = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, np.float32), (Q,)), np.float32)
        qn = None if qn is None else _f32(qn).reshape(-1, 3)
        out = np.empty((Q, 16), np.float64)
        lib().orc_ball_stats(self.h, _p(c), _p(qn), _p(rad), Q, mode, threads or ncores(), _p(out))
        return {"n": out[:, 0].astype(np.int64), "avg_dist": out[:, 1], "num_offsets": out[:, 2].astype(np.int64), "curvature": out[:, 3],
                "centroid": out[:, 4:7], "cov": out[:, 7:13]}

    def visibility(self, p, dirs, scale, mode=MODE_CANON, threads=None):
        p = _f32(p).reshape(-1, 3)
        S = p.shape[0]
        dirs = _f32(dirs)
        per_sample = dirs.ndim == 3
        D = dirs.shape[-2]
        vis = np.empty(S, np.float64)
        nesc = np.empty(S, np.int32)
        lib().orc_visibility(self.h, _p(p), S, _p(dirs), int(D), int(per_sample), float(scale), mode, threads or ncores(), _p(vis), _p(nesc))
        return vis, nesc

    def avg_separation(self, seeds):
        s = np.ascontiguousarray(seeds, np.int64)
        return float(lib().orc_avg_separation(self.h, _p(s), len(s)))

    def ball_graph(self, max_degree, init_radius, mode=MODE_REFERENCE, threads=None):
        idx = np.empty((self.n, max_degree), np.int64)
        sep = np.empty((self.n, max_degree), np.float32)
        cnt = np.empty(self.n, np.int32)
        lib().orc_ball_graph(self.h, max_degree, float(init_radius), mode, threads or ncores(), _p(idx), _p(sep), _p(cnt))
        return idx, sep, cnt

    def sdf(self, p, n, scale, only_interior=True, mode=MODE_CANON, threads=None):
        p = _f32(p).reshape(-1, 3)
        n = _f32(n).reshape(-1, 3)
        S = p.shape[0]
        ro = np.empty((S * 30, 3), np.float32)
        rd = np.empty((S * 30, 3), np.float32)
        hi = np.empty(S * 30, np.int64)
        ht = np.empty(S * 30, np.float32)
        out = np.empty(S, np.float64)
        lib().orc_sdf(self.h, _p(p), _p(n), S, float(scale), int(only_interior), mode, threads or ncores(),
                      _p(ro), _p(rd), _p(hi), _p(ht), _p(out))
        return out, ro, rd, hi, ht

    def icp(self, src, min_iters=3, max_iters=10, frac_thresh=1e-4, mode=MODE_CANON, threads=None):
        s = _f32(src).reshape(-1, 3)
        n = s.shape[0]
        idx = np.full((max_iters + 1, n), -1, np.int64)
        tr = np.zeros((3, 4), np.float32)
        err = np.zeros(1, np.float32)
        b = lib().orc_icp(self.h, _p(s), n, min_iters, max_iters, float(frac_thresh), mode, threads or ncores(),
                          _p(idx), _p(tr), _p(err))
        return tr, float(err[0]), idx[:b]

    def icp_iterates(self, src, min_iters=3, max_iters=10, frac_thresh=1e-4, mode=MODE_CANON, threads=None):
        """icp() that also returns the source positions each batch queried: (tr, err, idx [B, n], positions [B, n, 3])."""
        s = _f32(src).reshape(-1, 3)
        n = s.shape[0]
        idx = np.full((max_iters + 1, n), -1, np.int64)
        pos = np.zeros((max_iters + 1, n, 3), np.float32)
        tr = np.zeros((3, 4), np.float32)
        err = np.zeros(1, np.float32)
        b = lib().orc_icp_ex(self.h, _p(s), n, min_iters, max_iters, float(frac_thresh), mode, threads or ncores(),
                             _p(idx), _p(tr), _p(err), _p(pos))
        return tr, float(err[0]), idx[:b], pos[:b]

    def area_sample(self, count, seed):
        p = np.empty((count, 3), np.float32)
        n = np.empty((count, 3), np.float32)
        t = np.empty(count, np.int64)
        lib().orc_area_sample(self.h, count, seed, _p(p), _p(n), _p(t))
        return p, n, t


def tri_point(v9, p):
    v = _f32(v9).reshape(9)
    q = _f32(p).reshape(3)
    cp = np.empty(3, np.float32)
    d2 = lib().orc_tri_point(_p(v), _p(q), _p(cp))
    return float(d2), cp


def ray_tri(v9, o, d):
    return float(lib().orc_ray_tri(_p(_f32(v9).reshape(9)), _p(_f32(o).reshape(3)), _p(_f32(d).reshape(3))))


def seg_point(a, b, p):
    """LineSegment3(a_i, b_i) against point p_i: squared distance and the closest point on the segment."""
    a, b, p = (_f32(x).reshape(-1, 3) for x in (a, b, p))
    n = a.shape[0]
    d2 = np.empty(n, np.float32)
    sp = np.empty((n, 3), np.float32)
    lib().orc_seg_point_batch(_p(a), _p(b), _p(p), n, _p(d2), _p(sp))
    return d2, sp


def tri_seg(v9, a, b):
    """Triangle i against LineSegment3(a_i, b_i): squared distance, closest point on the triangle, on the segment."""
    v9 = _f32(v9).reshape(-1, 9)
    a, b = (_f32(x).reshape(-1, 3) for x in (a, b))
    n = a.shape[0]
    d2 = np.empty(n, np.float32)
    tp = np.empty((n, 3), np.float32)
    sp = np.empty((n, 3), np.float32)
    lib().orc_tri_seg_batch(_p(v9), _p(a), _p(b), n, _p(d2), _p(tp), _p(sp))
    return d2, tp, sp


def tri_tri(va9, vb9):
    """Triangle a_i .squaredDistance(b_i): squared distance, closest point on a, on b."""
    a, b = (_f32(x).reshape(-1, 9) for x in (va9, vb9))
    n = a.shape[0]
    d2 = np.empty(n, np.float32)
    pa = np.empty((n, 3), np.float32)
    pb = np.empty((n, 3), np.float32)
    lib().orc_tri_tri_batch(_p(a), _p(b), n, _p(d2), _p(pa), _p(pb))
    return d2, pa, pb


def fast_sqrt(x):
    return float(lib().orc_fast_sqrt(float(x)))


def box_ray_time(lo, hi, o, d, max_time=-1.0):
    return float(lib().orc_box_ray_time(_p(_f32(lo)), _p(_f32(hi)), _p(_f32(o)), _p(_f32(d)), float(max_time)))


def icp_error(tr12, src, dst):
    s, d, t = _f32(src).reshape(-1, 3), _f32(dst).reshape(-1, 3), _f32(tr12).reshape(12)
    return float(lib().orc_icp_error(_p(t), s.shape[0], _p(s), _p(d)))


def icp_step(src, dst):
    s, d = _f32(src).reshape(-1, 3), _f32(dst).reshape(-1, 3)
    tr = np.zeros((3, 4), np.float32)
    lib().orc_icp_step(s.shape[0], _p(s), _p(d), _p(tr))
    return tr


def polar_rotation(M):
    m = np.ascontiguousarray(M, np.float64).reshape(9)
    r = np.empty(9, np.float64)
    lib().orc_polar_rotation(_p(m), _p(r))
    return r.reshape(3, 3)


def thea_random_uniform01(seed, n):
    out = np.empty(n, np.float32)
    lib().orc_thea_random_uniform01(seed, n, _p(out))
    return out


# ---- shortest paths on a given sample graph (idx / sep: P x K rows, -1 padded): ShortestPaths.hpp:166-353 restated ----
def _graph(idx, sep):
    idx = np.ascontiguousarray(idx, np.int64)
    sep = np.ascontiguousarray(sep, np.float32)
    return idx, sep, idx.shape[0], idx.shape[1]


def graph_contract(idx, sep, n_orig, threads=None):
    """PointSet3::extractOriginalAdjacencies: (idx, sep, count) rows of the first n_orig samples."""
    idx, sep, P, K = _graph(idx, sep)
    oi, os_, oc = np.empty((n_orig, K), np.int64), np.empty((n_orig, K), np.float32), np.empty(n_orig, np.int32)
    lib().orc_graph_contract(_p(idx), _p(sep), P, K, int(n_orig), threads or ncores(), _p(oi), _p(os_), _p(oc))
    return oi, os_, oc


def graph_geodesic(idx, sep, seeds, max_distance, hist_range, nbins, threads=None):
    """Geodesic neighbourhood of every seed: (sum [Q] float64, count [Q] int64, hist [Q x nbins] float64)."""
    idx, sep, P, K = _graph(idx, sep)
    seeds = np.ascontiguousarray(seeds, np.int64)
    Q = len(seeds)
    s, c, h = np.empty(Q, np.float64), np.empty(Q, np.int64), np.empty((Q, max(nbins, 1)), np.float64)
    lib().orc_graph_geodesic(_p(idx), _p(sep), P, K, _p(seeds), Q, float(max_distance), float(hist_range), int(nbins), threads or ncores(), _p(s), _p(c), _p(h) if nbins > 0 else None)
    return s, c, h


def graph_fps(idx, sep, n_desired):
    """FurthestPointSampling::subsample's selection on a given graph."""
    idx, sep, P, K = _graph(idx, sep)
    sel = np.full(int(n_desired), -1, np.int64)
    n = lib().orc_graph_fps(_p(idx), _p(sep), P, K, int(n_desired), _p(sel))
    return sel[:n]
