"""thea_b200 -- B200-native batched spatial-query engine behind Thea's Algorithms::BvhN query API.

The product is ``libtgq.so`` (hand-written sm_100a CUDA behind the C ABI of ``include/thea_gpu_query.h``) and the
header-only C++ adaptor ``include/thea_gpu/GpuBvh.hpp``.  This package is the thin ctypes binding used by the tests
and ``bench.py``; it holds no compute path of its own and raises if the CUDA library cannot be loaded or no GPU exists.
"""
from .engine import Engine, TgqError, lib_path, load_library  # noqa: F401
