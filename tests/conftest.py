import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _emu_lib():
    """DFX_EMU_LIB=<path to libdfx_emu_full.so>: run the `gpu` tests against the library built for
    the CPU emulator (tests/emu/, the same sources compiled by g++ on a fiber SIMT runtime; the
    tcgen05 paths are absent).  Test infrastructure: slow, small shapes only, no CUDA tensors."""
    return os.environ.get("DFX_EMU_LIB")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")
    if _emu_lib():
        from distributed_faiss_b200 import engine

        engine.LIB_PATH = _emu_lib()
        engine._lib = None


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu() or _emu_lib():
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
