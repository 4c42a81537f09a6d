"""Development aid: per-query work of the k-NN traversals on a cfg2-shaped sample set (host emulation counters): nodes, points tested,
candidates appended and compactions for candidate buffers with seeded / unseeded start radius, and for the sorted sets."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import emu_lib
from thea_b200 import gen


def morton_order(p):
    lo, hi = p.min(0), p.max(0)
    g = np.clip(((p - lo) / (hi - lo) * 1023).astype(np.int64), 0, 1023).astype(np.uint64)
    def spread(x):
        x = (x | (x << np.uint64(16))) & np.uint64(0x030000FF)
        x = (x | (x << np.uint64(8))) & np.uint64(0x0300F00F)
        x = (x | (x << np.uint64(4))) & np.uint64(0x030C30C3)
        x = (x | (x << np.uint64(2))) & np.uint64(0x09249249)
        return x
    return np.argsort((spread(g[:, 0]) << np.uint64(2)) | (spread(g[:, 1]) << np.uint64(1)) | spread(g[:, 2]), kind="stable")


def main(P=200000, k=16, stride=1):
    mesh = gen.torus_mesh(300)
    p, n, _ = gen.area_samples(mesh, P)
    p = np.ascontiguousarray(p[morton_order(p)])
    E = emu_lib.Emu(points=p)
    L = emu_lib.lib()
    a, b = C.c_longlong(), C.c_longlong()
    q = np.ascontiguousarray(p[5000:25000])
    # queries in the order a lane would see them: every `stride`-th
    for name, buffered, qs in (("sorted sets", False, q), ("buffers, seeded by the previous query (stride 1)", True, q),
                               ("buffers, seeded, stride 32", True, np.ascontiguousarray(q[::32])), ("buffers, seeded, stride 8", True, np.ascontiguousarray(q[::8]))):
        emu_lib.set_buffered(buffered)
        L.emu_counters(C.byref(a), C.byref(b), 1); L.emu_set_counters(C.byref(a), C.byref(b), 1)
        E.knn(qs, k, exclude_self=False)
        L.emu_counters(C.byref(a), C.byref(b), 1); nodes, prims = a.value, b.value
        L.emu_set_counters(C.byref(a), C.byref(b), 1)
        print("%-50s nodes %.1f points %.1f appends %.1f compactions %.2f" % (name, nodes / len(qs), prims / len(qs), a.value / len(qs), b.value / len(qs)))


if __name__ == "__main__":
    main()
