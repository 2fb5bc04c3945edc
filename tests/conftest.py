import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the native pieces exist (idempotent; the GPU box uses the prebuilt files)."""
    import __graft_entry__ as g

    so = os.path.join(ROOT, "aac.js_b200", "libaacfb.so")
    emu = os.path.join(ROOT, "aac.js_b200", "libaacfb_emul.so")
    orc = os.path.join(ROOT, "oracle", "libaacfb_oracle.so")
    if not (os.path.exists(so) and os.path.exists(emu) and os.path.exists(orc)):
        g.build()
