import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """A tdnnf context on cuda:0 (GPU tests only): fails loudly when the library is missing."""
    import torch

    from tdnnf_nas_b200 import capi

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    c = capi.Context(0)
    c.use_current_stream()
    yield c
    torch.cuda.synchronize()
    c.close()
