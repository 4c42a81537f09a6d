"""The C-ABI shared library: loads on a CPU-only box, exports exactly what include/thea_gpu_query.h declares, and
refuses to run without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "thea_gpu_query.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(tgq_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(built):
    from thea_b200 import engine
    L = engine.load_library()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/thea_gpu_query.h but not exported by libtgq.so"
    assert sorted(engine.SYMBOLS) == names


def test_product_never_imports_the_oracle():
    import itertools
    for base, _, files in itertools.chain(os.walk(os.path.join(ROOT, "thea_b200")), os.walk(os.path.join(ROOT, "include"))):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(base, f), errors="ignore").read()
                for needle in ("import oracle", "from oracle", "liborc", "orc_", "thea_oracle", "libtgq_emu", "tests/emu"):
                    assert needle not in src, f"{base}/{f} references test infrastructure ({needle})"


def test_no_gpu_means_loud_failure(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from thea_b200 import Engine, TgqError
    with pytest.raises(TgqError) as ei:
        Engine(0)
    assert ei.value.code == 3 and "no CPU path" in str(ei.value)


def test_null_handle_calls_return_einval(built):
    from thea_b200 import engine
    L = engine.load_library()
    assert L.tgq_set_option(None, 1, 1) == 1
    assert L.tgq_info(None, None, None, None, None) == 1
    assert L.tgq_destroy(None) == 0
    assert L.tgq_create(0, None) == 1
