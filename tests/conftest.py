import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "h-emcee_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running oracle replay")


@pytest.fixture(scope="session")
def hx():
    """The product package; GPU tests fail loudly (not skip) when the extension is missing."""
    import torch
    import hemcee_b200
    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    hemcee_b200._lib.load()
    return hemcee_b200
