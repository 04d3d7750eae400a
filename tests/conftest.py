import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library; CPU tests only load it and look at symbols, GPU tests call into it."""
    from rlrep_b200 import _lib, build
    if not _lib.LIB_PATH.exists():
        build.build()
    return _lib.load()
