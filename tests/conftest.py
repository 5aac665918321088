import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """A fresh checkout has no built library (build artefacts are kept out of history): build it once, in-tree, before the
    first test needs it (nvcc cross-compiles sm_100a without a GPU; ~30 s).  With the library present this is a no-op."""
    from jitr_b200 import _cabi

    if not os.path.exists(_cabi.LIB_PATH) and not os.environ.get("JITR_B200_LIB"):
        from jitr_b200 import _build

        _build.build()


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        return cache[name]

    return load
