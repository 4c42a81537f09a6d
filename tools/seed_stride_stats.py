"""Development aid: closest-point visit counts when query i of the Morton-ordered cfg5 batch is seeded with the answer of query
i - stride (host emulation of the engine headers).  stride 32 ~ queries dealt to idle lanes one at a time (round 1-2 kernels),
stride 1 = a lane working through a run of consecutive queries."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import emu_lib
from thea_b200 import gen


def morton30(q, lo, hi):
    c = 0.5 * (lo + hi); e = hi - lo
    u = (q - (c - e)) * (0.5 / e)
    g = np.clip((u * 1024.0).astype(np.int64), 0, 1023).astype(np.uint64)
    def spread(x):
        x = (x | (x << np.uint64(16))) & np.uint64(0x030000FF)
        x = (x | (x << np.uint64(8))) & np.uint64(0x0300F00F)
        x = (x | (x << np.uint64(4))) & np.uint64(0x030C30C3)
        x = (x | (x << np.uint64(2))) & np.uint64(0x09249249)
        return x
    return (spread(g[:, 0]) << np.uint64(2)) | (spread(g[:, 1]) << np.uint64(1)) | spread(g[:, 2])


def main(grid=2236, total=50_000_000, window=20000, nwin=3):
    t0 = time.time()
    mesh = gen.torus_mesh(grid); b = gen.mesh_bounds(mesh)
    q = gen.closest_queries(total, b)
    code = morton30(q.astype(np.float64), np.asarray(b[0], np.float64), np.asarray(b[1], np.float64))
    order = np.argsort(code, kind="stable")
    print("queries sorted %.0fs" % (time.time() - t0), flush=True)
    if os.environ.get("EMU_VARIANT"): emu_lib.use_variant(os.environ["EMU_VARIANT"])
    E = emu_lib.Emu(tris9=mesh, leaf=4); E.set_wide(False)
    print("emu built %.0fs" % (time.time() - t0), flush=True)
    L = emu_lib.lib()
    L.emu_closest_counts_stride.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    for w in range(nwin):
        start = (w + 1) * total // (nwin + 1)
        qs = np.ascontiguousarray(q[order[start:start + window]])
        for stride in ((32,) if os.environ.get("EMU_ONLY32") else (0, 64, 32, 8, 2, 1)):
            nodes = np.empty(window, np.int32); prims = np.empty(window, np.int32)
            L.emu_closest_counts_stride(E.h, qs.ctypes.data_as(C.c_void_p), window, stride, nodes.ctypes.data_as(C.c_void_p), prims.ctypes.data_as(C.c_void_p))
            s = slice(64, None)
            print("window %d stride %2d: nodes %.1f  screened+tested leaf triangles %.1f" % (w, stride, nodes[s].mean(), prims[s].mean()), flush=True)


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:]))
