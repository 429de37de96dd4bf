import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _gpu_count() -> int:
    try:
        from convdr_b200 import _lib
        return int(_lib.load().b2f_device_count())
    except Exception:
        return 0


@pytest.fixture(scope="session")
def gpu_count():
    return _gpu_count()


def pytest_sessionstart(session):
    """The CUDA library and the oracle's C restatement are built in-tree before COLLECTION (test modules bind
    the library at import time), so a fresh clone runs without a separate build step."""
    from convdr_b200 import build
    from oracle import build as oracle_build
    build.build_cuda()
    oracle_build.build_oracle()


def pytest_collection_modifyitems(config, items):
    # a `gpu` test on a box without a GPU is a configuration error, not a skip: fail loudly,
    # unless the run did not ask for gpu tests (-m "not gpu" deselects them anyway)
    pass
