import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """A pn_ctx on cuda:0 (GPU tests only)."""
    import torch
    from peanut_b200 import _lib
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    c = _lib.Context(0)
    yield c
    c.close()
