import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")
    config.addinivalue_line("markers", "slow: long-running")


@pytest.fixture(scope='session')
def golden():
    import numpy as np
    return {k: np.load(os.path.join(GOLDEN, k + '.npz'))
            for k in ('bounding', 'multi', 'chains')}


@pytest.fixture
def fake_ops(monkeypatch):
    """CPU stand-in for dynesty_b200.ops built from the oracle (tests only)."""
    import fake_backend
    fake_backend.install(monkeypatch)
    return fake_backend
