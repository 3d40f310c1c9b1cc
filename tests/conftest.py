import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """ccg context on cuda:0 — fails loudly (no CPU fallback) when the library or the GPU is missing."""
    import camera_calibrator_b200 as cc

    cc.lib()  # raises if libccg.so was not built
    c = cc.Context(0)
    yield c
    c.close()
