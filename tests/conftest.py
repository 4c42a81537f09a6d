import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Build artefacts exist (oracle, emulation, CUDA library)."""
    import __graft_entry__ as g
    if not (os.path.exists(os.path.join(ROOT, "thea_b200", "libtgq.so"))
            and os.path.exists(os.path.join(ROOT, "oracle", "liborc.so"))
            and os.path.exists(os.path.join(ROOT, "tests", "emu", "libtgq_emu.so"))):
        g.build()
    return True
