"""ctypes binding of tests/emu/libtgq_emu.so -- the TEST-ONLY host build of the engine headers (see emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_NAME = "libtgq_emu.so"


def use_variant(name):
    """Switches every later Emu to another host build of the same headers (e.g. "libtgq_emu96.so", the 96-byte oriented-box node);
    returns the previous name.  Existing Emu objects must be dropped first."""
    global _LIB, _NAME
    prev, _NAME, _LIB = _NAME, name, None
    return prev


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "emu"), "-s"])
        L = C.CDLL(os.path.join(_HERE, "emu", _NAME))
        vp = C.c_void_p
        L.emu_build.restype = vp
        L.emu_build.argtypes = [vp, C.c_int64, C.c_int, vp, C.c_int]
        L.emu_destroy.argtypes = [vp]
        L.emu_num_nodes.restype = C.c_int64
        L.emu_num_nodes.argtypes = [vp]
        L.emu_num_wide.restype = C.c_int64
        L.emu_num_wide.argtypes = [vp]
        L.emu_set_wide.argtypes = [vp, C.c_int]
        L.emu_set_buffered.argtypes = [C.c_int]
        L.emu_closest.argtypes = [vp, vp, C.c_int64, C.c_double, C.c_int, vp, vp, vp]
        L.emu_knn.argtypes = [vp, vp, C.c_int64, C.c_int, C.c_float, C.c_int, vp, vp]
        L.emu_ball_graph.argtypes = [vp, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, vp, vp, vp]
        L.emu_set_labels.argtypes = [vp, vp]
        L.emu_ball_range.argtypes = [vp, vp, vp, C.c_int64, C.c_int, C.c_int, vp, vp]
        L.emu_ball_graph_surface.argtypes = [vp, vp, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, vp, vp, vp]
        L.emu_ball_hist.argtypes = [vp, vp, vp, C.c_int64, C.c_int, vp]
        L.emu_spin_image.argtypes = [vp, vp, vp, C.c_int64, C.c_int, C.c_int, C.c_float, vp]
        L.emu_ball_stats.argtypes = [vp, vp, vp, vp, C.c_int64, vp]
        L.emu_closest_filtered.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_float, C.c_double, vp, vp, vp]
        L.emu_knn_filtered.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_int, C.c_float, C.c_float, vp, vp]
        L.emu_ray.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_int, vp, vp]
        L.emu_select.argtypes = [vp, C.c_int, C.c_int]
        L.emu_std_nth_element.argtypes = [vp, C.c_int, C.c_int]
        L.emu_sdf_statistics.restype = C.c_double
        L.emu_sdf_statistics.argtypes = [vp, vp, C.c_int, C.c_float]
        L.emu_sdf_rays.argtypes = [vp, vp, C.c_int64, C.c_float, vp, vp]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def set_buffered(on):
    """k-NN / sample graph emulation: candidate buffers (True, what the engine runs by default) or sorted sets (TGQ_OPT_SET_BUFFER = 0)."""
    lib().emu_set_buffered(int(bool(on)))


class Emu:
    def __init__(self, tris9=None, points=None, normals=None, leaf=None):
        L = lib()
        if leaf is None:
            leaf = 4 if tris9 is not None else 15          # the engine's defaults by structure kind (TGQ_OPT_LEAF_SIZE = 0)
        if tris9 is not None:
            self.src = np.ascontiguousarray(tris9, np.float32).reshape(-1, 9)
            self.n = len(self.src)
            self.h = L.emu_build(_p(self.src), self.n, 1, None, leaf)
        else:
            self.src = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
            self.nrm = None if normals is None else np.ascontiguousarray(normals, np.float32)
            self.n = len(self.src)
            self.h = L.emu_build(_p(self.src), self.n, 0, _p(self.nrm), leaf)

    def __del__(self):
        try:
            lib().emu_destroy(self.h)
        except Exception:
            pass

    def set_wide(self, on):
        """Closest point on triangle structures through the 4-wide nodes (default) or the binary oriented-box nodes."""
        lib().emu_set_wide(self.h, int(bool(on)))

    def num_wide(self):
        return lib().emu_num_wide(self.h)

    def num_nodes(self):
        return lib().emu_num_nodes(self.h)

    def closest(self, q, dist_bound=-1.0, exclude_self=False, seeded=True):
        """seeded: every query starts from the previous query's answer (what a kernel lane does on ordered batches)."""
        q = np.ascontiguousarray(q, np.float32)
        Q = len(q)
        idx = np.empty(Q, np.int32)
        d2 = np.empty(Q, np.float32)
        cp = np.empty((Q, 3), np.float32)
        flag = int(exclude_self) if seeded else -int(exclude_self) - 1
        lib().emu_closest(self.h, _p(q), Q, float(dist_bound), flag, _p(idx), _p(d2), _p(cp))
        return idx, d2, cp

    def closest_segment(self, a, b, dist_bound=-1.0, exclude=None):
        """Closest element to LineSegment3(a_i, b_i): idx, d2, point on the element, point on the segment."""
        L = lib()
        L.emu_closest_segment.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_int] + [C.c_void_p] * 4
        a = np.ascontiguousarray(a, np.float32).reshape(-1, 3)
        b = np.ascontiguousarray(b, np.float32).reshape(-1, 3)
        Q = len(a)
        ex = None if exclude is None else np.ascontiguousarray(exclude, np.int32).reshape(Q, -1)
        idx, d2 = np.empty(Q, np.int32), np.empty(Q, np.float32)
        ep, sp = np.empty((Q, 3), np.float32), np.empty((Q, 3), np.float32)
        L.emu_closest_segment(self.h, _p(a), _p(b), Q, float(dist_bound), _p(ex), 0 if ex is None else ex.shape[1], _p(idx), _p(d2), _p(ep), _p(sp))
        return idx, d2, ep, sp

    def closest_triangle(self, q9, dist_bound=-1.0):
        """Closest element to Triangle3(q_i): idx, d2, point on the element, point on the query triangle."""
        L = lib()
        L.emu_closest_triangle.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double] + [C.c_void_p] * 4
        q = np.ascontiguousarray(q9, np.float32).reshape(-1, 9)
        Q = len(q)
        idx, d2 = np.empty(Q, np.int32), np.empty(Q, np.float32)
        ep, qp = np.empty((Q, 3), np.float32), np.empty((Q, 3), np.float32)
        L.emu_closest_triangle(self.h, _p(q), Q, float(dist_bound), _p(idx), _p(d2), _p(ep), _p(qp))
        return idx, d2, ep, qp

    def range_all(self, a, b, box=False):
        """All elements in Ball3(a, b) or AxisAlignedBox3(a, b): ascending element indices."""
        L = lib()
        L.emu_range_all.restype = C.c_int64
        L.emu_range_all.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int64]
        a = np.ascontiguousarray(a, np.float32)
        bb = np.ascontiguousarray(b, np.float32).reshape(3) if box else np.zeros(3, np.float32)
        cap = 1 << 22
        idx = np.empty(cap, np.int32)
        n = L.emu_range_all(self.h, int(box), _p(a), _p(bb), 0.0 if box else float(b), _p(idx), cap)
        return np.sort(idx[:n])

    def ball_range(self, c, radius, max_out=32, exclude_self=False):
        c = np.ascontiguousarray(c, np.float32)
        Q = len(c)
        rad = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, np.float32), (Q,)), np.float32)
        idx = np.empty((Q, max_out), np.int32)
        cnt = np.empty(Q, np.int32)
        lib().emu_ball_range(self.h, _p(c), _p(rad), Q, int(max_out), int(exclude_self), _p(idx), _p(cnt))
        return idx, cnt

    def spin_image(self, q, qn, nr, nh, max_extent):
        q = np.ascontiguousarray(q, np.float32); qn = np.ascontiguousarray(qn, np.float32)
        out = np.empty((len(q), nr, nh), np.uint32)
        lib().emu_spin_image(self.h, _p(q), _p(qn), len(q), int(nr), int(nh), float(max_extent), _p(out))
        return out

    def ball_hist(self, c, radius, nbins):
        c = np.ascontiguousarray(c, np.float32)
        Q = len(c)
        rad = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, np.float32), (Q,)), np.float32)
        out = np.empty((Q, nbins), np.uint32)
        lib().emu_ball_hist(self.h, _p(c), _p(rad), Q, int(nbins), _p(out))
        return out

    def ball_stats(self, c, radius, qn=None):
        c = np.ascontiguousarray(c, np.float32)
        Q = len(c)
        rad = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, np.float32), (Q,)), np.float32)
        qn = None if qn is None else np.ascontiguousarray(qn, np.float32)
        out = np.empty((Q, 16), np.float64)
        lib().emu_ball_stats(self.h, _p(c), _p(qn), _p(rad), Q, _p(out))
        return out

    def set_labels(self, labels):
        self._lab = None if labels is None else np.ascontiguousarray(labels, np.int32)
        lib().emu_set_labels(self.h, _p(self._lab))

    def closest_filtered(self, q, qn=None, qlabel=None, min_dot=0.0, dist_bound=-1.0):
        q = np.ascontiguousarray(q, np.float32)
        qn = None if qn is None else np.ascontiguousarray(qn, np.float32)
        ql = None if qlabel is None else np.ascontiguousarray(qlabel, np.int32)
        Q = len(q)
        idx = np.empty(Q, np.int32)
        d2 = np.empty(Q, np.float32)
        cp = np.empty((Q, 3), np.float32)
        lib().emu_closest_filtered(self.h, _p(q), _p(qn), _p(ql), Q, float(min_dot), float(dist_bound), _p(idx), _p(d2), _p(cp))
        return idx, d2, cp

    def knn_filtered(self, q, k, qn=None, qlabel=None, min_dot=0.0, bound_ub=np.inf):
        q = np.ascontiguousarray(q, np.float32)
        qn = None if qn is None else np.ascontiguousarray(qn, np.float32)
        ql = None if qlabel is None else np.ascontiguousarray(qlabel, np.int32)
        Q = len(q)
        idx = np.empty((Q, k), np.int32)
        d2 = np.empty((Q, k), np.float32)
        lib().emu_knn_filtered(self.h, _p(q), _p(qn), _p(ql), Q, int(k), float(min_dot), float(bound_ub), _p(idx), _p(d2))
        return idx, d2

    def knn(self, q, k, bound_ub=np.inf, exclude_self=False):
        q = np.ascontiguousarray(q, np.float32)
        Q = len(q)
        idx = np.empty((Q, k), np.int32)
        d2 = np.empty((Q, k), np.float32)
        lib().emu_knn(self.h, _p(q), Q, k, float(bound_ub), int(exclude_self), _p(idx), _p(d2))
        return idx, d2

    def ball_graph(self, k, r0, rounds=3, min_dot=-0.5, min_degree=None):
        if min_degree is None:
            min_degree = min(max(int(0.25 * k), 4), k)
        idx = np.empty((self.n, k), np.int32)
        sep = np.empty((self.n, k), np.float32)
        cnt = np.empty(self.n, np.int32)
        lib().emu_ball_graph(self.h, k, float(r0), rounds, float(min_dot), min_degree, _p(idx), _p(sep), _p(cnt))
        return idx, sep, cnt

    def ball_graph_surface(self, surface, k, r0, rounds=3, min_dot=-0.5, min_degree=None):
        if min_degree is None:
            min_degree = min(max(int(0.25 * k), 4), k)
        idx = np.empty((self.n, k), np.int32)
        sep = np.empty((self.n, k), np.float32)
        cnt = np.empty(self.n, np.int32)
        lib().emu_ball_graph_surface(self.h, surface.h, k, float(r0), rounds, float(min_dot), min_degree, _p(idx), _p(sep), _p(cnt))
        return idx, sep, cnt

    def ray(self, o, d, tmax=None, any_hit=False):
        o = np.ascontiguousarray(o, np.float32)
        d = np.ascontiguousarray(d, np.float32)
        tm = None if tmax is None else np.ascontiguousarray(tmax, np.float32)
        R = len(o)
        t = np.empty(R, np.float32)
        idx = np.empty(R, np.int32)
        lib().emu_ray(self.h, _p(o), _p(d), _p(tm), R, int(any_hit), _p(t), _p(idx))
        return t, idx


def select(a, nth):
    b = np.array(a, np.float64)
    lib().emu_select(_p(b), len(b), nth)
    return b


def std_nth_element(a, nth):
    b = np.array(a, np.float64)
    lib().emu_std_nth_element(_p(b), len(b), nth)
    return b


def sdf_statistics(values, weights, scale):
    v = np.array(values, np.float64)
    w = np.array(weights, np.float64)
    return float(lib().emu_sdf_statistics(_p(v), _p(w), len(v), float(scale)))


def sdf_rays(p, n, scale):
    p = np.ascontiguousarray(p, np.float32)
    n = np.ascontiguousarray(n, np.float32)
    S = len(p)
    o = np.empty((S * 30, 3), np.float32)
    d = np.empty((S * 30, 3), np.float32)
    lib().emu_sdf_rays(_p(p), _p(n), S, float(scale), _p(o), _p(d))
    return o, d
