"""Build recipe for libdfx.so (the C-ABI CUDA library, include/dfx.h).

nvcc cross-compiles for sm_100a without a GPU; the .so is built IN-TREE
(distributed_faiss_b200/libdfx.so) so that it travels to the GPU box.
`-fmad=false`: the kernels spell every fused multiply-add as __fmaf_rn so
that the arithmetic is exactly the canonical order of DESIGN.md (bit-exact
parity with oracle/dfx_oracle.c); the compiler must not contract anything else.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libdfx.so")
SOURCES = ["dfx_api.cu", "dfx_search.cu", "dfx_build.cu", "dfx_tc.cu", "dfx_scan_il.cu", "dfx_scan_il2.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-ccbin", "/usr/bin/g++",
    "--expt-relaxed-constexpr",
]


_HASH_MARK = b"dfx-src-sha256="


def source_hash() -> str:
    """sha256 over every source the library is compiled from (csrc/, include/) and the compiler
    flags.  It is compiled into the library (`dfx_version()` ends with it), so a stale binary is
    recognised by content -- file times do not survive a copy of the tree to another machine."""
    h = hashlib.sha256()
    h.update(" ".join(FLAGS).encode())
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    return h.hexdigest()[:32]


def embedded_hash(path: str = LIB):
    """the source hash a built library carries (None: no library, or one built without it)."""
    try:
        with open(path, "rb") as fh:
            blob = fh.read()
    except OSError:
        return None
    i = blob.find(_HASH_MARK)
    if i < 0:
        return None
    return blob[i + len(_HASH_MARK): i + len(_HASH_MARK) + 32].decode("ascii", "replace")


def is_current() -> bool:
    return embedded_hash() == source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    want = source_hash()
    if not force and embedded_hash() == want:
        return LIB
    # one builder at a time (several ranks of one job may find the same stale binary)
    import fcntl

    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and embedded_hash() == want:   # another process built it while we waited
            return LIB
        return _build_locked(want, verbose)


def _build_locked(want: str, verbose: bool) -> str:
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}; libdfx.so cannot be built")
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if src == "dfx_api.cu":
            cmd.insert(1, f'-DDFX_SRC_HASH="{want}"')
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB + ".tmp"
    cmd = [NVCC, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-ccbin", "/usr/bin/g++", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)   # readers see the old or the new library, never a partial one
    if embedded_hash() != want:
        raise RuntimeError("libdfx.so was built but does not carry the expected source hash")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
