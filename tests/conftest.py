import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native libraries once per session (no-op when up to date)."""
    from minivideo_b200 import build
    build.build_synth()
    build.build_oracle()
    build.build_front()
    build.build_gpu()
    build.build_reference()
    yield
