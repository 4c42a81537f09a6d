"""ctypes binding of libtgq.so (include/thea_gpu_query.h).

Arrays may be numpy arrays (host memory, staged by the library) or torch CUDA tensors (device pointers, zero copy).
Nothing here computes: every method forwards to one C-ABI entry point.  There is no fallback -- a missing library or a
missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KIND_NONE, KIND_TRIANGLES, KIND_POINTS = 0, 1, 2
OPT_SORT_QUERIES, OPT_LEAF_SIZE, OPT_ASYNC, OPT_CHUNK, OPT_POOL, OPT_ICP_REFERENCE_ORDER, OPT_WIDE = 1, 2, 3, 4, 5, 6, 7
OPT_SET_BUFFER = 10
MAX_K = 64

SYMBOLS = [
    "tgq_create", "tgq_destroy", "tgq_last_error", "tgq_set_stream", "tgq_set_option",
    "tgq_build_triangles", "tgq_build_triangles_indexed", "tgq_build_points", "tgq_refit_triangles",
    "tgq_info", "tgq_triangle_cache", "tgq_closest", "tgq_knn", "tgq_ball_graph", "tgq_avg_separation",
    "tgq_ray_first", "tgq_ray_any", "tgq_sdf", "tgq_icp_sums", "tgq_transform_points", "tgq_area_sample",
    "tgq_get_stats", "tgq_set_point_labels", "tgq_closest_filtered", "tgq_knn_filtered", "tgq_icp_align", "tgq_ball_range",
    "tgq_ball_stats", "tgq_visibility", "tgq_ball_graph_surface", "tgq_ball_histogram", "tgq_spin_image", "tgq_graph_contract", "tgq_graph_geodesic", "tgq_graph_fps", "tgq_set_transform",
    "tgq_closest_records", "tgq_ray_first_records", "tgq_synchronize", "tgq_scatter_records", "tgq_bucket_records", "tgq_range_query", "tgq_closest_pair", "tgq_k_closest_pairs",
    "tgq_closest_segment", "tgq_closest_triangle",
]


class TgqError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tgq error {code}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [("num_elements", C.c_int64), ("num_nodes", C.c_int64), ("node_bytes", C.c_int64),
                ("element_bytes", C.c_int64), ("last_build_ms", C.c_double), ("last_query_ms", C.c_double),
                ("last_sort_ms", C.c_double), ("last_launches", C.c_int64), ("total_launches", C.c_int64)]


def lib_path() -> str:
    return os.path.join(_HERE, "libtgq.so")


_LIBS = {}


def load_library(path=None):
    """Loads libtgq.so; raises if it has not been built (run ``python -c 'import __graft_entry__ as g; g.build()'``).
    `path` selects another build of the same ABI (the tests load libtgq_wide.so, the wide-margin variant)."""
    global _LIB
    if path is None and _LIB is not None:
        return _LIB
    if path is not None and path in _LIBS:
        return _LIBS[path]
    default = path is None
    path = lib_path() if path is None else path
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `make -C thea_b200/csrc` (there is no CPU fallback)")
    L = C.CDLL(path)
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
    L.tgq_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.tgq_destroy.argtypes = [vp]
    L.tgq_last_error.restype = C.c_char_p
    L.tgq_last_error.argtypes = [vp]
    L.tgq_set_stream.argtypes = [vp, vp]
    L.tgq_set_option.argtypes = [vp, C.c_int, i64]
    L.tgq_build_triangles.argtypes = [vp, vp, i64]
    L.tgq_build_triangles_indexed.argtypes = [vp, vp, i64, vp, i64]
    L.tgq_build_points.argtypes = [vp, vp, i64, vp]
    L.tgq_refit_triangles.argtypes = [vp, vp, i64]
    L.tgq_info.argtypes = [vp, vp, vp, vp, vp]
    L.tgq_triangle_cache.argtypes = [vp, vp, vp, vp, vp]
    L.tgq_closest.argtypes = [vp, vp, i64, f64, i32, vp, vp, vp]
    L.tgq_knn.argtypes = [vp, vp, i64, i32, f64, i32, vp, vp]
    L.tgq_ball_graph.argtypes = [vp, i32, f32, i32, f32, i32, vp, vp, vp]
    L.tgq_avg_separation.argtypes = [vp, vp, i32, vp]
    L.tgq_ray_first.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp]
    L.tgq_ray_any.argtypes = [vp, vp, vp, vp, i64, vp]
    L.tgq_sdf.argtypes = [vp, vp, vp, i64, f32, i32, vp, vp, vp]
    L.tgq_icp_sums.argtypes = [vp, vp, vp, i64, vp, vp]
    L.tgq_transform_points.argtypes = [vp, vp, i64, vp, vp]
    L.tgq_area_sample.argtypes = [vp, i64, C.c_uint64, vp, vp, vp]
    L.tgq_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.tgq_set_point_labels.argtypes = [vp, vp]
    L.tgq_ball_range.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp]
    L.tgq_ball_graph_surface.argtypes = [vp, vp, i32, f32, i32, f32, i32, vp, vp, vp]
    L.tgq_ball_histogram.argtypes = [vp, vp, vp, i64, i32, vp]
    L.tgq_spin_image.argtypes = [vp, vp, vp, i64, i32, i32, C.c_float, vp]
    L.tgq_graph_contract.argtypes = [vp, vp, vp, i64, i32, i64, vp, vp, vp]
    L.tgq_graph_geodesic.argtypes = [vp, vp, vp, i64, i32, vp, i64, C.c_float, C.c_float, i32, vp, vp, vp]
    L.tgq_graph_fps.argtypes = [vp, vp, vp, i64, i32, i64, vp, vp]
    L.tgq_set_transform.argtypes = [vp, vp]
    L.tgq_ball_stats.argtypes = [vp, vp, vp, vp, i64, vp]
    L.tgq_visibility.argtypes = [vp, vp, i64, vp, i32, i32, f32, vp, vp]
    L.tgq_icp_align.argtypes = [vp, vp, i64, i32, i32, f64, i32, vp, vp, vp, vp]
    L.tgq_closest_filtered.argtypes = [vp, vp, vp, vp, i64, f32, f64, vp, vp, vp]
    L.tgq_knn_filtered.argtypes = [vp, vp, vp, vp, i64, i32, f32, f64, vp, vp]
    L.tgq_closest_segment.argtypes = [vp, vp, vp, i32, i64, f64, vp, i32, vp, vp, vp, vp]
    L.tgq_closest_triangle.argtypes = [vp, vp, i64, f64, vp, vp, vp, vp]
    L.tgq_closest_records.argtypes = [vp, vp, i64, f64, i32, i32, vp, i32, i32, vp]
    L.tgq_ray_first_records.argtypes = [vp, vp, vp, vp, i64, i32, i32, vp, i32, i32, vp]
    L.tgq_scatter_records.argtypes = [vp, vp, i32, i64, vp, i32, vp]
    L.tgq_bucket_records.argtypes = [vp, vp, i64, i32, i64, i32, vp, i64, vp]
    L.tgq_range_query.argtypes = [vp, i32, vp, vp, i64, vp, vp, i64, vp]
    L.tgq_closest_pair.argtypes = [vp, vp, f64, vp, vp, vp, vp, vp]
    L.tgq_k_closest_pairs.argtypes = [vp, vp, i32, f64, vp, vp, vp, vp]
    for name in SYMBOLS:
        if name not in ("tgq_last_error",):
            getattr(L, name).restype = C.c_int
    if default:
        _LIB = L
    else:
        _LIBS[path] = L
    return L


def _ptr(a):
    """Raw pointer of a numpy array or torch tensor (None passes through)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(f"unsupported array type {type(a)}")


def _is_torch(a):
    return a is not None and not isinstance(a, np.ndarray) and hasattr(a, "data_ptr")


_TORCH_DTYPE = {np.dtype(np.float32): "float32", np.dtype(np.int32): "int32", np.dtype(np.float64): "float64", np.dtype(np.uint8): "uint8"}


def _as_in(a, dtype=np.float32):
    """Contiguous input of the right dtype: numpy arrays are converted, torch tensors must already match (a tensor of another
    dtype would be reinterpreted, not converted, by the C ABI: refused here)."""
    if a is None:
        return None
    if _is_torch(a):
        if not a.is_contiguous():
            raise ValueError("torch inputs must be contiguous")
        want = _TORCH_DTYPE.get(np.dtype(dtype))
        if want is not None and str(a.dtype) != "torch." + want:
            raise TypeError(f"torch input has dtype {a.dtype}, the engine needs torch.{want}")
        return a
    return np.ascontiguousarray(a, dtype=dtype)


class Engine:
    """One BvhN-like structure on one GPU.  Method names follow the C ABI; see include/thea_gpu_query.h."""

    def __init__(self, device: int = 0, lib: str | None = None):
        self._L = load_library(lib)
        h = C.c_void_p()
        rc = self._L.tgq_create(device, C.byref(h))
        if rc != 0:
            raise TgqError(rc, self._L.tgq_last_error(None).decode())
        self._h = h
        self.device = device
        self._explicit_stream = False

    def _in(self, a, dtype=np.float32):
        """_as_in plus the two things a CUDA tensor needs: it must live on this engine's device, and the handle's work must be
        ordered after whatever produced it.  Unless set_stream() was called explicitly, the handle is therefore moved onto torch's
        CURRENT stream of the device whenever a CUDA tensor is passed (results written into CUDA tensors are ordered the same way)."""
        a = _as_in(a, dtype)
        if _is_torch(a) and getattr(a, "is_cuda", False):
            if a.device.index != self.device:
                raise ValueError(f"tensor lives on {a.device}, the engine on cuda:{self.device}")
            if not self._explicit_stream:
                import torch
                # torch's default stream has handle 0, which tgq_set_stream reads as "the handle's own stream": name it by
                # cudaStreamLegacy (0x1) instead
                self._ck(self._L.tgq_set_stream(self._h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream or 1)))
        return a

    def close(self):
        if getattr(self, "_h", None):
            self._L.tgq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise TgqError(rc, self._L.tgq_last_error(self._h).decode())

    # ---- configuration ----
    def set_option(self, opt, value):
        self._ck(self._L.tgq_set_option(self._h, opt, int(value)))

    def synchronize(self):
        self._ck(self._L.tgq_synchronize(self._h))

    def set_stream(self, cuda_stream_ptr):
        """Pins the handle to one CUDA stream (0 / None: back to following torch's current stream for CUDA tensors, the handle's
        own stream otherwise)."""
        self._explicit_stream = bool(cuda_stream_ptr)
        self._ck(self._L.tgq_set_stream(self._h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    # ---- build ----
    def build_triangles(self, v9):
        v9 = self._in(v9)
        T = int((v9.size if isinstance(v9, np.ndarray) else v9.numel()) // 9)
        self._ck(self._L.tgq_build_triangles(self._h, _ptr(v9), T))
        return T

    def build_triangles_indexed(self, verts, tris):
        verts = self._in(verts)
        tris = self._in(tris, np.int32)
        self._ck(self._L.tgq_build_triangles_indexed(self._h, _ptr(verts), int(verts.shape[0]), _ptr(tris), int(tris.shape[0])))
        return int(tris.shape[0])

    def build_points(self, p3, n3=None):
        p3 = self._in(p3)
        n3 = self._in(n3)
        self._ck(self._L.tgq_build_points(self._h, _ptr(p3), int(p3.shape[0]), _ptr(n3)))
        return int(p3.shape[0])

    def set_transform(self, tr12=None):
        """BvhN::setTransform (3x4 [L|t], row-major) / clearTransform (None): queries see the transformed elements."""
        tr = None if tr12 is None else np.ascontiguousarray(tr12, np.float32).reshape(12)
        self._ck(self._L.tgq_set_transform(self._h, _ptr(tr)))

    def refit_triangles(self, v9):
        v9 = self._in(v9)
        T = int((v9.size if isinstance(v9, np.ndarray) else v9.numel()) // 9)
        self._ck(self._L.tgq_refit_triangles(self._h, _ptr(v9), T))

    def info(self):
        kind = C.c_int32()
        n = C.c_int64()
        lo = (C.c_float * 3)()
        hi = (C.c_float * 3)()
        self._ck(self._L.tgq_info(self._h, C.byref(kind), C.byref(n), lo, hi))
        return kind.value, n.value, np.array(lo[:], np.float32), np.array(hi[:], np.float32)

    def triangle_cache(self):
        _, n, _, _ = self.info()
        nn = np.empty((n, 3), np.float32)
        d = np.empty(n, np.float32)
        ax = np.empty(n, np.int32)
        ar = np.empty(n, np.float32)
        self._ck(self._L.tgq_triangle_cache(self._h, _ptr(nn), _ptr(d), _ptr(ax), _ptr(ar)))
        return nn, d, ax, ar

    # ---- queries (numpy in -> numpy out; torch in -> caller supplies torch outputs) ----
    def closest(self, q3, dist_bound=-1.0, exclude_self=False, out=None, want_cp=True):
        q3 = self._in(q3)
        Q = int(q3.shape[0])
        if out is None:
            if _is_torch(q3):
                import torch
                idx = torch.empty(Q, dtype=torch.int32, device=q3.device)
                d2 = torch.empty(Q, dtype=torch.float32, device=q3.device)
                cp = torch.empty((Q, 3), dtype=torch.float32, device=q3.device) if want_cp else None
            else:
                idx = np.empty(Q, np.int32)
                d2 = np.empty(Q, np.float32)
                cp = np.empty((Q, 3), np.float32) if want_cp else None
        else:
            idx, d2, cp = out
        self._ck(self._L.tgq_closest(self._h, _ptr(q3), Q, float(dist_bound), int(exclude_self), _ptr(idx), _ptr(d2), _ptr(cp)))
        return idx, d2, cp

    def knn(self, q3, k, dist_bound=-1.0, exclude_self=False, out=None):
        q3 = self._in(q3)
        Q = int(q3.shape[0])
        if out is None:
            if _is_torch(q3):
                import torch
                idx = torch.empty((Q, k), dtype=torch.int32, device=q3.device)
                d2 = torch.empty((Q, k), dtype=torch.float32, device=q3.device)
            else:
                idx = np.empty((Q, k), np.int32)
                d2 = np.empty((Q, k), np.float32)
        else:
            idx, d2 = out
        self._ck(self._L.tgq_knn(self._h, _ptr(q3), Q, int(k), float(dist_bound), int(exclude_self), _ptr(idx), _ptr(d2)))
        return idx, d2

    # ---- filtered nearest neighbours (point structures; Register.cpp NNFilter semantics) ----
    def set_point_labels(self, labels):
        lab = None if labels is None else self._in(labels, np.int32)
        self._ck(self._L.tgq_set_point_labels(self._h, _ptr(lab)))

    def closest_filtered(self, q3, qn3=None, qlabel=None, min_dot=0.0, dist_bound=-1.0):
        q3 = self._in(q3)
        qn3 = self._in(qn3)
        qlabel = self._in(qlabel, np.int32)
        Q = int(q3.shape[0])
        idx = np.empty(Q, np.int32)
        d2 = np.empty(Q, np.float32)
        cp = np.empty((Q, 3), np.float32)
        self._ck(self._L.tgq_closest_filtered(self._h, _ptr(q3), _ptr(qn3), _ptr(qlabel), Q, float(min_dot), float(dist_bound), _ptr(idx), _ptr(d2), _ptr(cp)))
        return idx, d2, cp

    def closest_segment(self, a3, b3, dist_bound=-1.0, exclude=None, b_is_direction=False):
        """closestElement / closestPair<MetricL2>(LineSegment3(a_i, b_i), dist_bound) per query (Tools/MeshFix/MeshFix.cpp:756-763):
        (idx, d2, closest point on the element, closest point on the segment).  exclude: Q x (<= 4) element indices each query
        ignores, -1 = unused (MeshFix's VertexFilter as data).  b_is_direction: b3 holds the stored direction vectors, not the second endpoints."""
        a3 = self._in(a3)
        b3 = self._in(b3)
        Q = int(a3.shape[0])
        ex = None
        if exclude is not None:
            ex = np.ascontiguousarray(exclude, np.int32).reshape(Q, -1)
        idx = np.empty(Q, np.int32)
        d2 = np.empty(Q, np.float32)
        ep = np.empty((Q, 3), np.float32)
        sp = np.empty((Q, 3), np.float32)
        self._ck(self._L.tgq_closest_segment(self._h, _ptr(a3), _ptr(b3), int(b_is_direction), Q, float(dist_bound), _ptr(ex), 0 if ex is None else int(ex.shape[1]),
                                             _ptr(idx), _ptr(d2), _ptr(ep), _ptr(sp)))
        return idx, d2, ep, sp

    def closest_triangle(self, q9, dist_bound=-1.0):
        """closestPair<MetricL2>(Triangle3(q_i), dist_bound, ..., true) per query triangle (Q x 9 floats):
        (idx, d2, closest point on the element, closest point on the query triangle)."""
        q9 = self._in(q9)
        Q = int(q9.size // 9) if isinstance(q9, np.ndarray) else int(q9.numel() // 9)
        idx = np.empty(Q, np.int32)
        d2 = np.empty(Q, np.float32)
        ep = np.empty((Q, 3), np.float32)
        qp = np.empty((Q, 3), np.float32)
        self._ck(self._L.tgq_closest_triangle(self._h, _ptr(q9), Q, float(dist_bound), _ptr(idx), _ptr(d2), _ptr(ep), _ptr(qp)))
        return idx, d2, ep, qp

    def knn_filtered(self, q3, k, qn3=None, qlabel=None, min_dot=0.0, dist_bound=-1.0):
        q3 = self._in(q3)
        qn3 = self._in(qn3)
        qlabel = self._in(qlabel, np.int32)
        Q = int(q3.shape[0])
        idx = np.empty((Q, k), np.int32)
        d2 = np.empty((Q, k), np.float32)
        self._ck(self._L.tgq_knn_filtered(self._h, _ptr(q3), _ptr(qn3), _ptr(qlabel), Q, int(k), float(min_dot), float(dist_bound), _ptr(idx), _ptr(d2)))
        return idx, d2

    def ball_range(self, c3, radius, max_out=32, exclude_self=False):
        """rangeQueryIndices with a ball per query: (idx [Q x max_out] lowest indices ascending, count [Q])."""
        c3 = self._in(c3)
        Q = int(c3.shape[0])
        rad = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, np.float32), (Q,)), np.float32)
        idx = np.empty((Q, max_out), np.int32) if max_out > 0 else None
        cnt = np.empty(Q, np.int32)
        self._ck(self._L.tgq_ball_range(self._h, _ptr(c3), _ptr(rad), Q, int(max_out), int(exclude_self), _ptr(idx), _ptr(cnt)))
        return idx, cnt

    def spin_image(self, q3, qn3, num_radial_bins, num_height_bins, max_extent):
        """SpinImage::compute(position, normal, ...) for every query: [Q x radial x height] uint32 counts over ALL samples of the
        point structure (tgq_spin_image); max_extent = 0.75f * scale in the reference."""
        q3 = self._in(q3)
        qn3 = self._in(qn3)
        Q = int(q3.shape[0])
        img = np.zeros((Q, int(num_radial_bins), int(num_height_bins)), np.uint32)
        self._ck(self._L.tgq_spin_image(self._h, _ptr(q3), _ptr(qn3), Q, int(num_radial_bins), int(num_height_bins), float(max_extent), _ptr(img)))
        return img

    # ---- shortest paths on a sample graph (rows as ball_graph returns them) ----
    def graph_contract(self, idx, sep, n_orig):
        """PointSet3::extractOriginalAdjacencies: the graph over originals [0, n_orig) + dense samples contracted to the originals
        (tgq_graph_contract): (idx [n_orig x K], sep, count)."""
        idx = self._in(idx, np.int32); sep = self._in(sep)
        P, K = int(idx.shape[0]), int(idx.shape[1])
        oi, os_, oc = np.empty((n_orig, K), np.int32), np.empty((n_orig, K), np.float32), np.empty(n_orig, np.int32)
        self._ck(self._L.tgq_graph_contract(self._h, _ptr(idx), _ptr(sep), P, K, int(n_orig), _ptr(oi), _ptr(os_), _ptr(oc)))
        return oi, os_, oc

    def graph_geodesic(self, idx, sep, seeds, max_distance, hist_range, num_bins=0):
        """Geodesic neighbourhood of every seed sample (tgq_graph_geodesic): (sum [Q] float64, count [Q] int32, hist [Q x num_bins] or None)."""
        idx = self._in(idx, np.int32); sep = self._in(sep); seeds = self._in(seeds, np.int32)
        P, K, Q = int(idx.shape[0]), int(idx.shape[1]), int(seeds.shape[0])
        s, c = np.empty(Q, np.float64), np.empty(Q, np.int32)
        h = np.empty((Q, num_bins), np.uint32) if num_bins > 0 else None
        self._ck(self._L.tgq_graph_geodesic(self._h, _ptr(idx), _ptr(sep), P, K, _ptr(seeds), Q, float(max_distance), float(hist_range), int(num_bins),
                                            _ptr(s), _ptr(c), _ptr(h)))
        return s, c, h

    def graph_fps(self, idx, sep, n_desired):
        """FurthestPointSampling::subsample on a given graph (tgq_graph_fps): selected sample indices (possibly fewer than asked)."""
        idx = self._in(idx, np.int32); sep = self._in(sep)
        P, K = int(idx.shape[0]), int(idx.shape[1])
        sel = np.full(int(n_desired), -1, np.int32)
        n = C.c_int64(0)
        self._ck(self._L.tgq_graph_fps(self._h, _ptr(idx), _ptr(sep), P, K, int(n_desired), _ptr(sel), C.byref(n)))
        return sel[:n.value]

    def ball_histogram(self, c3, radius, num_bins):
        """LocalDistanceHistogram (EUCLIDEAN) for every ball: [Q x num_bins] uint32 counts (tgq_ball_histogram)."""
        c3 = self._in(c3)
        Q = int(c3.shape[0])
        rad = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, np.float32), (Q,)), np.float32)
        hist = np.empty((Q, num_bins), np.uint32)
        self._ck(self._L.tgq_ball_histogram(self._h, _ptr(c3), _ptr(rad), Q, int(num_bins), _ptr(hist)))
        return hist

    def k_closest_pairs(self, query, k, dist_bound=-1.0):
        """tgq_k_closest_pairs: BvhN::kClosestPairs between this point structure (the target) and the point structure `query`:
        (target indices, query indices, squared distances) of the k closest pairs, ascending."""
        t, q, d = np.empty(k, np.int64), np.empty(k, np.int64), np.empty(k, np.float32)
        n = C.c_int32(0)
        self._ck(self._L.tgq_k_closest_pairs(self._h, query._h, int(k), float(dist_bound), _ptr(t), _ptr(q), _ptr(d), C.byref(n)))
        return t[:n.value], q[:n.value], d[:n.value]

    def closest_pair(self, query, dist_bound=-1.0):
        """tgq_closest_pair: BvhN::closestPair between this structure (the target) and the structure `query` (points or triangles).
        Returns (target index, query index, d2, point on the query element, point on the target element); indices -1 when none."""
        ti, qi, d2 = C.c_int64(), C.c_int64(), C.c_float()
        qp, tp = np.zeros(3, np.float32), np.zeros(3, np.float32)
        self._ck(self._L.tgq_closest_pair(self._h, query._h, float(dist_bound), C.byref(ti), C.byref(qi), C.byref(d2), _ptr(qp), _ptr(tp)))
        return int(ti.value), int(qi.value), float(d2.value), qp, tp

    def range_query(self, a3, b, box=False):
        """tgq_range_query: all elements of a point structure in Q balls (a3 = centres, b = radii) or boxes (a3 = low, b = high corners).
        Returns (offsets [Q + 1] int64, idx [total] int32): CSR, ascending element index per range."""
        a3 = np.ascontiguousarray(a3, np.float32).reshape(-1, 3)
        Q = a3.shape[0]
        b = np.ascontiguousarray(b, np.float32).reshape(-1, 3) if box else np.ascontiguousarray(np.broadcast_to(np.asarray(b, np.float32), (Q,)), np.float32)
        off = np.zeros(Q + 1, np.int64)
        total = C.c_int64()
        self._ck(self._L.tgq_range_query(self._h, 1 if box else 0, _ptr(a3), _ptr(b), Q, _ptr(off), None, 0, C.byref(total)))
        idx = np.empty(max(1, total.value), np.int32)
        if total.value:
            self._ck(self._L.tgq_range_query(self._h, 1 if box else 0, _ptr(a3), _ptr(b), Q, _ptr(off), _ptr(idx), total.value, C.byref(total)))
        return off, idx[:total.value]

    def ball_stats(self, c3, radius, qn3=None, out=None):
        """Per-ball neighbourhood sums (tgq_ball_stats): [Q x 16] doubles, layout in include/thea_gpu_query.h."""
        c3 = self._in(c3)
        Q = int(c3.shape[0])
        if _is_torch(radius):
            rad = radius
        else:
            rad = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, np.float32), (Q,)), np.float32)
        qn3 = self._in(qn3)
        if out is None:
            out = np.empty((Q, 16), np.float64)
        self._ck(self._L.tgq_ball_stats(self._h, _ptr(c3), _ptr(qn3), _ptr(rad), Q, _ptr(out)))
        return out

    def visibility(self, p3, dirs3, scale):
        """Visibility::compute for every sample with caller-supplied directions ([D x 3] shared or [S x D x 3]):
        (visibility [S] float64, num_escaped [S] int32)."""
        p3 = self._in(p3)
        dirs3 = self._in(dirs3)
        S = int(p3.shape[0])
        per_sample = len(dirs3.shape) == 3
        D = int(dirs3.shape[-2])
        vis = np.empty(S, np.float64)
        nesc = np.empty(S, np.int32)
        self._ck(self._L.tgq_visibility(self._h, _ptr(p3), S, _ptr(dirs3), D, int(per_sample), float(scale), _ptr(vis), _ptr(nesc)))
        return vis, nesc

    def ball_graph(self, max_degree, init_radius, rounds=3, min_dot=-0.5, min_degree=None, out=None):
        _, n, _, _ = self.info()
        if min_degree is None:
            min_degree = min(max(int(0.25 * max_degree), 4), max_degree)   # PointSet3.hpp:665
        if out is None:
            idx = np.empty((n, max_degree), np.int32)
            sep = np.empty((n, max_degree), np.float32)
            cnt = np.empty(n, np.int32)
        else:
            idx, sep, cnt = out
        self._ck(self._L.tgq_ball_graph(self._h, int(max_degree), float(init_radius), int(rounds), float(min_dot), int(min_degree),
                                        _ptr(idx), _ptr(sep), _ptr(cnt)))
        return idx, sep, cnt

    def ball_graph_surface(self, surface, max_degree, init_radius, rounds=3, min_dot=-0.5, min_degree=None):
        """PointSet3::updateGraph(surface): the ball graph with the reachability ray test against the triangle Engine `surface`."""
        _, n, _, _ = self.info()
        if min_degree is None:
            min_degree = min(max(int(0.25 * max_degree), 4), max_degree)   # PointSet3.hpp:665
        idx = np.empty((n, max_degree), np.int32)
        sep = np.empty((n, max_degree), np.float32)
        cnt = np.empty(n, np.int32)
        self._ck(self._L.tgq_ball_graph_surface(self._h, surface._h, int(max_degree), float(init_radius), int(rounds), float(min_dot),
                                                int(min_degree), _ptr(idx), _ptr(sep), _ptr(cnt)))
        return idx, sep, cnt

    def avg_separation(self, seeds):
        s = np.ascontiguousarray(seeds, np.int32)
        out = C.c_float()
        self._ck(self._L.tgq_avg_separation(self._h, _ptr(s), len(s), C.byref(out)))
        return float(out.value)

    def ray_first(self, o3, d3, tmax=None, out=None, want_normal=False):
        o3 = self._in(o3)
        d3 = self._in(d3)
        tmax = self._in(tmax)
        R = int(o3.shape[0])
        if out is None:
            if _is_torch(o3):
                import torch
                t = torch.empty(R, dtype=torch.float32, device=o3.device)
                idx = torch.empty(R, dtype=torch.int32, device=o3.device)
                nrm = torch.empty((R, 3), dtype=torch.float32, device=o3.device) if want_normal else None
            else:
                t = np.empty(R, np.float32)
                idx = np.empty(R, np.int32)
                nrm = np.empty((R, 3), np.float32) if want_normal else None
        else:
            t, idx, nrm = out
        self._ck(self._L.tgq_ray_first(self._h, _ptr(o3), _ptr(d3), _ptr(tmax), R, _ptr(t), _ptr(idx), _ptr(nrm)))
        return t, idx, nrm

    # ---- record outputs / sharding of a replicated batch by space (multi-GPU form; device tensors only) ----
    @staticmethod
    def _dest_array(dest):
        ptrs = [int(d.data_ptr()) if hasattr(d, "data_ptr") else int(d) for d in dest]
        return (C.c_void_p * len(ptrs))(*ptrs), len(ptrs)

    def closest_records(self, q3, dest, rank=0, world=1, dist_bound=-1.0, compact=False):
        """tgq_closest_records: q3 = the WHOLE batch (CUDA tensor); dest = the record arrays ([Q, 8] int32 / float32 views of 32-byte
        records) of every rank that must receive this rank's results, own copy included (tensors or raw device pointers); compact:
        one dense destination in processing order.  Returns the number of queries this rank answered."""
        q3 = self._in(q3)
        arr, n = self._dest_array(dest)
        share = C.c_int64()
        self._ck(self._L.tgq_closest_records(self._h, _ptr(q3), int(q3.shape[0]), float(dist_bound), int(rank), int(world), arr, n, int(compact), C.byref(share)))
        return int(share.value)

    def ray_first_records(self, o3, d3, dest, rank=0, world=1, tmax=None, compact=False):
        """tgq_ray_first_records: 8-byte records (t, idx) at the ray's position in every destination array (compact: 16-byte records
        (t, idx, ray index, 0) in processing order in one destination).  Returns the number of rays this rank answered."""
        o3 = self._in(o3)
        d3 = self._in(d3)
        tmax = self._in(tmax)
        arr, n = self._dest_array(dest)
        share = C.c_int64()
        self._ck(self._L.tgq_ray_first_records(self._h, _ptr(o3), _ptr(d3), _ptr(tmax), int(o3.shape[0]), int(rank), int(world), arr, n, int(compact), C.byref(share)))
        return int(share.value)

    def bucket_records(self, compact, n, kind, shard_len, world, buckets, cap, counts):
        """tgq_bucket_records: n compact records split by the rank owning their position (query index // shard_len)."""
        self._ck(self._L.tgq_bucket_records(self._h, C.c_void_p(int(compact.data_ptr())), int(n), int(kind), int(shard_len), int(world),
                                            C.c_void_p(int(buckets.data_ptr())), int(cap), C.c_void_p(int(counts.data_ptr()))))

    def scatter_records(self, compact, slots, cap, counts, kind, dest):
        """tgq_scatter_records: `slots` dense arrays of up to `cap` compact records, counts (int64 CUDA tensor) valid -> dest[query index]."""
        self._ck(self._L.tgq_scatter_records(self._h, C.c_void_p(int(compact.data_ptr())), int(slots), int(cap), C.c_void_p(int(counts.data_ptr())), int(kind),
                                             C.c_void_p(int(dest.data_ptr()))))

    def ray_any(self, o3, d3, tmax=None):
        o3 = self._in(o3)
        d3 = self._in(d3)
        tmax = self._in(tmax)
        R = int(o3.shape[0])
        hit = np.empty(R, np.uint8)
        self._ck(self._L.tgq_ray_any(self._h, _ptr(o3), _ptr(d3), _ptr(tmax), R, _ptr(hit)))
        return hit

    def sdf(self, p3, n3, scale, only_hit_interior=True, want_rays=True):
        p3 = self._in(p3)
        n3 = self._in(n3)
        S = int(p3.shape[0])
        sdf = np.empty(S, np.float64)
        rt = np.empty((S, 30), np.float32) if want_rays else None
        ri = np.empty((S, 30), np.int32) if want_rays else None
        self._ck(self._L.tgq_sdf(self._h, _ptr(p3), _ptr(n3), S, float(scale), int(only_hit_interior), _ptr(sdf), _ptr(rt), _ptr(ri)))
        return sdf, rt, ri

    def icp_sums(self, from3, to3, tr12=None):
        from3 = self._in(from3)
        to3 = self._in(to3)
        tr = None if tr12 is None else np.ascontiguousarray(tr12, np.float32).reshape(12)
        out = np.empty(16, np.float64)
        self._ck(self._L.tgq_icp_sums(self._h, _ptr(from3), _ptr(to3), int(from3.shape[0]), _ptr(tr), _ptr(out)))
        return out

    def icp_align(self, from3, min_iterations=3, max_iterations=10, fractional_error_threshold=1e-4, want_error=True, want_correspondences=True):
        """Icp3<Real>::align on the device (tgq_icp_align): returns (transform 3x4, error, correspondences [batches x n])."""
        from3 = self._in(from3)
        n = int(from3.shape[0])
        tr = np.empty(12, np.float32)
        err = C.c_float()
        nb = C.c_int32()
        corr = np.empty((max_iterations + 1, n), np.int32) if want_correspondences else None
        self._ck(self._L.tgq_icp_align(self._h, _ptr(from3), n, int(min_iterations), int(max_iterations), float(fractional_error_threshold),
                                       int(want_error), _ptr(tr), C.byref(err), C.byref(nb), _ptr(corr)))
        return tr.reshape(3, 4), float(err.value), (corr[:nb.value] if corr is not None else nb.value)

    def transform_points(self, in3, tr12, out=None):
        in3 = self._in(in3)
        tr = np.ascontiguousarray(tr12, np.float32).reshape(12)
        n = int(in3.shape[0])
        if out is None:
            if _is_torch(in3):
                import torch
                out = torch.empty((n, 3), dtype=torch.float32, device=in3.device)
            else:
                out = np.empty((n, 3), np.float32)
        self._ck(self._L.tgq_transform_points(self._h, _ptr(in3), n, _ptr(tr), _ptr(out)))
        return out

    def area_sample(self, count, seed):
        p = np.empty((count, 3), np.float32)
        n = np.empty((count, 3), np.float32)
        t = np.empty(count, np.int32)
        self._ck(self._L.tgq_area_sample(self._h, int(count), int(seed), _ptr(p), _ptr(n), _ptr(t)))
        return p, n, t

    def stats(self) -> dict:
        s = Stats()
        self._ck(self._L.tgq_get_stats(self._h, C.byref(s)))
        return {f[0]: getattr(s, f[0]) for f in Stats._fields_}
