import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI library, compiled in-tree if it is not there yet (nvcc cross-compiles on CPU)."""
    import orb_slam2_chinesenotes_b200 as ob
    if not os.path.exists(ob.LIB_PATH):
        ob.build()
    return ob.lib()
