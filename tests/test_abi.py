"""The C-ABI library loads and exports every symbol include/dfx.h declares; no compute here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "dfx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dfx_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    import __graft_entry__ as entry

    entry.build()
    from distributed_faiss_b200 import engine

    lib = ctypes.CDLL(engine.LIB_PATH)
    declared = _header_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dfx.h but not exported"
    assert sorted(engine.EXPORTED_SYMBOLS) == declared
    assert b"sm_100a" in engine.lib().dfx_version()
    # the library carries the hash of the sources it was compiled from (build.py rebuilds on mismatch)
    from distributed_faiss_b200 import build
    assert build.embedded_hash() == build.source_hash()
    assert engine.lib().dfx_version().decode().endswith(build.source_hash())


def test_no_cpu_fallback():
    """without a GPU every computing entry point fails loudly"""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from distributed_faiss_b200 import engine
    import numpy as np

    with pytest.raises(RuntimeError, match="CUDA device"):
        engine.GpuIndex(engine.KIND_FLAT, 8)
    with pytest.raises(RuntimeError):
        engine.merge(np.zeros((1, 1, 1), np.float32), np.zeros((1, 1, 1), np.int64))
    from distributed_faiss_b200.index import Index
    from distributed_faiss_b200.index_cfg import IndexCfg

    ix = Index(IndexCfg(index_builder_type="flat", dim=8))
    ix.add_batch(np.zeros((4, 8), np.float32), None)
    with pytest.raises(RuntimeError):
        ix.train()


def test_sass_is_blackwell_native():
    """the coarse-quantizer object code contains tcgen05 / TMA / TMEM instructions"""
    import shutil
    import subprocess

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    obj = os.path.join(ROOT, "distributed_faiss_b200", "build", "dfx_tc.o")
    if not os.path.exists(obj):
        pytest.skip("object file not kept")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass
