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
    """The C oracle (and, in the build container, oracle/_ref) must exist before any test."""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "oracle", "_build", "liboracle_c.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    yield
