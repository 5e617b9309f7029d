"""TIMED CPU BASELINE of the IVF-PQ search path (bench infrastructure -- NOT product code and not
the parity checker; see the header of cpu_ivfpq.c).

`CpuIVFPQ(state)` holds ONE shard exactly as `oracle.OracleIVFPQ.get_state()` / `GpuIndex.get_state()`
lay it out and searches it the way faiss's IndexIVFPQ does on a CPU: BLAS sgemm for the coarse
quantizer (MKL through torch.mm, all host threads), then cpu_ivfpq.c (OpenMP) for probe selection,
per-query tables and the list scan.  Only tests/ and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "cpu_ivfpq.c")
_LIB_PATH = os.path.join(_HERE, "libdfx_cpu_baseline.so")
_FLAGS = ["-O3", "-ffast-math", "-fopenmp", "-fPIC", "-shared", "-Wall"]

_lib = None
_lib_kind = None


def build(force: bool = False) -> str:
    """portable build (x86-64-v3), in-tree: travels to the GPU box with the snapshot"""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(_SRC):
        subprocess.check_call(["/usr/bin/gcc", *_FLAGS, "-march=x86-64-v3", "-o", _LIB_PATH, _SRC, "-lm"])
    return _LIB_PATH


def _build_native():
    """-march=native build on the machine that runs the bench (AVX-512 where the host has it)"""
    if not os.path.exists(_SRC) or not os.path.exists("/usr/bin/gcc"):
        return None
    out = os.path.join(tempfile.gettempdir(), f"libdfx_cpu_baseline_native_{os.getuid()}.so")
    try:
        subprocess.check_call(["/usr/bin/gcc", *_FLAGS, "-march=native", "-o", out, _SRC, "-lm"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return out
    except Exception:
        return None


def lib(native: bool = True):
    global _lib, _lib_kind
    if _lib is None:
        path = _build_native() if native else None
        _lib_kind = "native" if path else "x86-64-v3"
        _lib = C.CDLL(path or build())
    return _lib


def host_threads() -> int:
    return len(os.sched_getaffinity(0))


def set_threads(n: int) -> int:
    """torchrun exports OMP_NUM_THREADS=1 to its workers: set the OpenMP / MKL pools explicitly"""
    import torch

    n = max(1, int(n))
    lib().cpu_set_num_threads(C.c_int(n))
    torch.set_num_threads(n)
    return n


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class CpuIVFPQ:
    def __init__(self, st, threads: int = 0):
        self.d = int(st["d"])
        self.M = int(st["M"])
        self.nlist = int(st["nlist"])
        self.centroids = np.ascontiguousarray(st["centroids"], dtype=np.float32)
        self.cnorm = (self.centroids.astype(np.float32) ** 2).sum(1).astype(np.float32)
        self.codebooks = np.ascontiguousarray(st["codebooks"], dtype=np.float32)
        self.list_off = np.ascontiguousarray(st["list_off"], dtype=np.int64)
        self.codes = np.ascontiguousarray(st["codes"], dtype=np.uint8)
        self.tvals = np.ascontiguousarray(st["tvals"], dtype=np.float32)
        self.ids = np.ascontiguousarray(st["ids"], dtype=np.int64)
        self.ntotal = int(self.ids.shape[0])
        self.threads = set_threads(threads or host_threads())
        import torch

        self._cent_t = torch.from_numpy(self.centroids).t().contiguous()
        self.last_ndis = 0

    def search(self, xq, k: int, nprobe: int, qblock: int = 1024):
        import torch

        L = lib()
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        nq = xq.shape[0]
        nprobe = min(int(nprobe), self.nlist)
        keys = np.empty((nq, nprobe), dtype=np.int32)
        for q0 in range(0, nq, qblock):
            xb = torch.from_numpy(xq[q0:q0 + qblock])
            G = torch.mm(xb, self._cent_t).numpy()              # sgemm: [qb, nlist]
            kb = keys[q0:q0 + qblock]
            L.cpu_select_probes(C.c_int64(G.shape[0]), C.c_int64(self.nlist), C.c_int64(G.shape[1]), _p(G),
                                _p(self.cnorm), C.c_int(nprobe), _p(kb))
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        ndis = C.c_int64(0)
        L.cpu_ivfpq_scan(C.c_int64(nq), C.c_int(self.d), _p(xq), _p(self.centroids), C.c_int(self.M),
                         _p(self.codebooks), _p(self.list_off), _p(self.codes), _p(self.tvals), _p(self.ids),
                         _p(keys), C.c_int(nprobe), C.c_int(k), _p(D), _p(I), C.byref(ndis))
        self.last_ndis = int(ndis.value)
        return D, I
