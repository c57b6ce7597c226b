import os
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


SESSION_START = time.time()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture
def emu_backend():
    """Route the engine's native calls to the torch restatement in tests/emu.py (CPU host-logic tests only)."""
    import emu
    emu.install()
    yield emu
    emu.uninstall()
