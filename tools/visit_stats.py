"""Development aid: per-query node / exact-triangle-test counts of the closest-point traversal (host emulation)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import emu_lib
from thea_b200 import gen

def main(grid=708, Q=20000, leaf=4, tight=False):
    mesh = gen.torus_mesh(grid); b = gen.mesh_bounds(mesh)
    E = emu_lib.Emu(tris9=mesh, leaf=leaf)
    q = gen.closest_queries(Q, b)
    L = emu_lib.lib()
    L.emu_closest_counts.argtypes = [C.c_void_p] * 2 + [C.c_int64] + [C.c_void_p] * 3
    nodes = np.empty(Q, np.int32); prims = np.empty(Q, np.int32)
    idx, d2, cp = E.closest(q)
    ub = np.nextafter(d2, np.float32(np.inf)) if tight else None
    L.emu_closest_counts(E.h, q.ctypes.data_as(C.c_void_p), Q, None if ub is None else ub.ctypes.data_as(C.c_void_p), nodes.ctypes.data_as(C.c_void_p), prims.ctypes.data_as(C.c_void_p))
    d = np.sqrt(d2)
    for name, a in (("nodes", nodes), ("prims", prims)):
        print(name, "mean %.1f" % a.mean(), "p50 %d p90 %d p99 %d p99.9 %d max %d" % tuple(np.percentile(a, [50, 90, 99, 99.9, 100])))
    # by distance decile
    order = np.argsort(d)
    for k in range(10):
        s = order[k * Q // 10:(k + 1) * Q // 10]
        print("dist %.3f-%.3f nodes %.1f prims %.1f" % (d[s[0]], d[s[-1]], nodes[s].mean(), prims[s].mean()))
    return q, nodes, prims, d

if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 708, int(sys.argv[2]) if len(sys.argv) > 2 else 20000, tight=len(sys.argv) > 3)
