import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle

    oracle.build()
    return oracle


def clustered(rs, n, d, ncl=64, sigma=0.3):
    """small clustered dataset so that IVF lists are non-trivial"""
    centers = rs.randn(ncl, d).astype(np.float32)
    lab = rs.randint(0, ncl, size=n)
    return (centers[lab] + sigma * rs.randn(n, d)).astype(np.float32)
