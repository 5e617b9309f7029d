"""The IVF-PQ block-layout kernels, compiled for the CPU from the SAME device headers nvcc compiles
(distributed_faiss_b200/csrc/*_dev.cuh) and executed on a fiber SIMT runtime (tests/emu/): layout
conversion, K3 pq_prep, K4 v2 (shipping scan, also verified on hardware) and K4 v3 (experimental
scan_variant=2, never run on hardware yet) against the oracle, bit for bit.

What this covers: the kernels' C++ -- indexing, layouts, the order of the floating-point
operations, the warp-collective protocols, shared-memory sizes (reads are bounds-checked,
dynamic shared memory is poisoned).  What it cannot cover: the PTX primitives of dfx_ptx.cuh
(replaced by plain C++ here), timing, and real concurrency between warps.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
NONE = np.uint64(0xFFFFFFFFFFFFFFFF)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def emu():
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("g++ not available")
    out = os.path.join(EMU_DIR, "_build", "libdfx_emu.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    srcs = [os.path.join(EMU_DIR, "emu_pq.cpp"), os.path.join(EMU_DIR, "simt.h"),
            os.path.join(EMU_DIR, "shim", "cuda_runtime.h")]
    csrc = os.path.join(ROOT, "distributed_faiss_b200", "csrc")
    srcs += [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        cmd = [cxx, "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-DDFX_EMU",
               "-Wno-unknown-pragmas", "-Wno-attributes", "-I", os.path.join(EMU_DIR, "shim"), "-I", csrc,
               "-o", out, os.path.join(EMU_DIR, "emu_pq.cpp")]
        subprocess.run(cmd, check=True)
    return C.CDLL(out)


def f2val(comp):
    k = (comp >> np.uint64(32)).astype(np.uint32)
    u = np.where(k & np.uint32(0x80000000), k & np.uint32(0x7FFFFFFF), ~k).astype(np.uint32)
    return u.view(np.float32)


class Pipeline:
    """host side of dfx_search_impl for IVF-PQ M=32 on a block layout, kernels run by the emulator"""

    def __init__(self, lib, ix, layout, seed=0):
        self.lib, self.ix, self.layout = lib, ix, layout
        lib.emu_set_seed(C.c_uint64(seed))
        st = ix.get_state()
        self.st = st
        self.list_off = np.ascontiguousarray(st["list_off"], dtype=np.int64)
        self.nlist = len(self.list_off) - 1
        self.codes = np.ascontiguousarray(st["codes"], dtype=np.uint8)
        self.tvals = np.ascontiguousarray(st["tvals"], dtype=np.float32)
        self.ids = np.ascontiguousarray(st["ids"], dtype=np.int32)
        self.blk_off = np.zeros(self.nlist + 1, dtype=np.int64)
        self.blk_off[1:] = np.cumsum((np.diff(self.list_off) + 31) // 32)
        self.nblk = int(self.blk_off[-1])
        nb = max(self.nblk, 1)
        self.il_codes = np.full((nb, 1024), 0xEE, dtype=np.uint8)
        self.il_tvals = np.full((nb, 32), np.nan, dtype=np.float32)
        self.il_ids = np.full((nb, 32), -7, dtype=np.int32)
        if self.nblk:
            lib.emu_rm_to_il(layout, C.c_int64(self.nlist), _p(self.list_off), _p(self.blk_off), _p(self.codes),
                             _p(self.tvals), _p(self.ids), C.c_int64(self.nblk), _p(self.il_codes),
                             _p(self.il_tvals), _p(self.il_ids))

    def back_to_row_major(self):
        codes = np.zeros_like(self.codes)
        tvals = np.zeros_like(self.tvals)
        ids = np.zeros_like(self.ids)
        if self.nblk:
            self.lib.emu_il_to_rm(self.layout, C.c_int64(self.nlist), _p(self.list_off), _p(self.blk_off),
                                  _p(self.il_codes), _p(self.il_tvals), _p(self.il_ids), C.c_int64(self.nblk),
                                  _p(codes), _p(tvals), _p(ids))
        return codes, tvals, ids

    def prep(self, xq, nprobe):
        st = self.st
        d, M = st["d"], st["M"]
        keys64, _ = O.coarse(st["coarse_metric"], st["centroids"], xq, nprobe)
        keys = np.ascontiguousarray(keys64, dtype=np.int32)
        nq = xq.shape[0]
        width = 64 if self.layout == 2 else 32            # layouts 1 and 3 share the [code][m] table
        lut = np.full((nq, 256, width), np.nan, dtype=np.float32)
        dis0 = np.full((nq, nprobe), np.nan, dtype=np.float32)
        cb = np.ascontiguousarray(st["codebooks"], dtype=np.float32)
        cent = np.ascontiguousarray(st["centroids"], dtype=np.float32)
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        self.lib.emu_pq_prep(_p(xq), C.c_int64(nq), d, M, 256, d // M, _p(cb), _p(cent), _p(keys), nprobe,
                             _p(lut), _p(dis0), 2 if self.layout == 2 else 1)
        return keys, lut, dis0

    def search(self, xq, k, nprobe, G):
        keys, lut, dis0 = self.prep(xq, nprobe)
        nq = xq.shape[0]
        ngroups = (nprobe + G - 1) // G
        KP = 32
        while KP < k:
            KP *= 2
        cap = 2 * KP
        part = np.full((nq, ngroups, k), 0x1234, dtype=np.uint64)
        args = (_p(lut), _p(dis0), _p(keys), C.c_int64(nq), nprobe, G, ngroups, _p(self.blk_off), _p(self.il_codes),
                _p(self.il_tvals), _p(self.il_ids), k, cap, _p(part))
        if self.layout == 2:
            rc = self.lib.emu_scan_v3(*args)
        else:  # layout 1: shipping kernel; layout 3: the same kernel on coalesced halves
            rc = self.lib.emu_scan_v2(*args, 1 if self.layout == 3 else 0)
        assert rc == 0
        D = np.full((nq, k), np.inf, dtype=np.float32)
        I = np.full((nq, k), -1, dtype=np.int64)
        for q in range(nq):
            c = np.sort(part[q].reshape(-1))[:k]          # the downstream selection kernel
            live = c != NONE
            D[q, live] = f2val(c[live])
            I[q, live] = (c[live] & np.uint64(0xFFFFFFFF)).astype(np.int64)
        return D, I, (keys, lut, dis0)


def _index(n=3000, d=128, nlist=8, dup=200, seed=0, n_add=None):
    rs = np.random.RandomState(seed)
    x = rs.randn(n, d).astype(np.float32)
    if dup:
        x[-dup:] = x[:dup]                               # exact duplicates: ties decided by the id
    ix = O.make_index("ivf_pq", d, O.METRIC_L2, nlist=nlist, M=32)
    ix.train_niter = 3
    ix.train(x[:1500])
    ix.add(x if n_add is None else x[:n_add])
    return ix, x, rs


@pytest.fixture(scope="module")
def built():
    return _index()


@pytest.mark.parametrize("layout", [1, 2, 3])
def test_layout_conversion_round_trip(emu, built, layout):
    ix, _, _ = built
    p = Pipeline(emu, ix, layout)
    byte_of = np.array([[emu_byte(layout, v, m) for m in range(32)] for v in range(32)])
    b, l = 3, int(np.searchsorted(p.blk_off, 3, side="right") - 1)
    base = int(p.list_off[l]) + (b - int(p.blk_off[l])) * 32
    for v in (0, 7, 31):
        if base + v < p.list_off[l + 1]:
            assert np.array_equal(p.il_codes[b, byte_of[v]], p.codes[base + v])
    codes, tvals, ids = p.back_to_row_major()
    assert np.array_equal(codes, p.codes) and np.array_equal(tvals, p.tvals) and np.array_equal(ids, p.ids)
    real = p.il_ids.reshape(-1) >= 0                      # padding: id -1, t = +inf
    assert real.sum() == len(p.ids) and np.isinf(p.il_tvals.reshape(-1)[~real]).all()


def emu_byte(layout, v, m):
    t = (m - v) & 31
    if layout == 2:
        return (t >> 4) * 512 + v * 16 + (t & 15)
    u, w, i, j = v >> 3, v & 7, m & 7, m >> 3
    lane, r, tt = 8 * u + i, w ^ i, (j - u) & 3
    if layout == 3:
        return (r >> 2) * 512 + lane * 16 + (r & 3) * 4 + tt
    return lane * 32 + r * 4 + tt


@pytest.mark.parametrize("layout", [1, 2])
def test_pq_prep_tables_and_dis0(emu, built, layout):
    ix, x, rs = built
    p = Pipeline(emu, ix, layout)
    xq = x[:3] + 0.1 * rs.randn(3, 128).astype(np.float32)
    keys, lut, dis0 = p.prep(xq, 5)
    for q in range(3):
        ref = ix.query_lut(xq[q])                         # [m][j]
        if layout == 1:
            assert lut[q].tobytes() == np.ascontiguousarray(ref.T).tobytes()
        else:
            assert lut[q].tobytes() == np.ascontiguousarray(ref.T[:, np.arange(64) & 31]).tobytes()
        for j, l in enumerate(keys[q]):
            assert dis0[q, j] == np.float32(O.warp_dot(xq[q], ix.centroids[l], 1))


@pytest.mark.parametrize("layout", [1, 2])
@pytest.mark.parametrize("nq", [1, 8, 13])
def test_pq_prep_variant_2_equals_variant_1(emu, built, layout, nq):
    """experimental K3 (transposed codebook, 8 queries per CTA) writes the same bits"""
    ix, x, rs = built
    p = Pipeline(emu, ix, layout, seed=nq)
    xq = np.ascontiguousarray(x[100:100 + nq] + 0.1 * rs.randn(nq, 128).astype(np.float32))
    keys, lut, dis0 = p.prep(xq, 5)
    lut2 = np.full_like(lut, np.nan)
    dis2 = np.full_like(dis0, np.nan)
    cb = np.ascontiguousarray(p.st["codebooks"], dtype=np.float32)
    cent = np.ascontiguousarray(p.st["centroids"], dtype=np.float32)
    emu.emu_pq_prep2(_p(xq), C.c_int64(nq), 128, _p(cb), _p(cent), _p(keys), 5, _p(lut2), _p(dis2),
                     1 if layout == 2 else 0)
    assert lut2.tobytes() == lut.tobytes() and dis2.tobytes() == dis0.tobytes()


CASES = [  # layout, k, nprobe, G, scheduler seed
    (1, 10, 4, 4, 0), (1, 10, 4, 1, 7),                   # shipping kernel: validates the emulator itself
    (2, 10, 4, 4, 0), (2, 10, 4, 1, 7), (2, 1, 3, 2, 3), (2, 32, 8, 8, 11),   # v3, register top-k
    (2, 40, 4, 4, 0), (2, 100, 8, 3, 5),                  # v3, shared-memory top-k (k > 32)
    (3, 10, 4, 4, 0), (3, 50, 5, 2, 9),                   # shipping kernel on block layout 3
]


@pytest.mark.parametrize("layout,k,nprobe,G,seed", CASES)
def test_scan_kernels_equal_oracle(emu, built, layout, k, nprobe, G, seed):
    ix, x, rs = built
    rs = np.random.RandomState(100 + seed)
    xq = x[rs.randint(0, len(x), 3)] + 0.05 * rs.randn(3, 128).astype(np.float32)
    xq[0] = ix.reconstruct_rows([5])[0]                  # sits on a duplicated point: exact ties
    ix.nprobe = nprobe
    Dref, Iref = ix.search(xq, k)
    D, I, _ = Pipeline(emu, ix, layout, seed).search(xq, k, nprobe, G)
    assert np.array_equal(I, Iref)
    assert D.tobytes() == Dref.tobytes()


@pytest.mark.parametrize("layout", [1, 2, 3])
def test_scan_kernels_short_lists(emu, layout):
    """lists shorter than a block, empty lists, fewer than k results"""
    ix, x, rs = _index(n=2000, nlist=8, dup=0, seed=4, n_add=21)
    xq = rs.randn(2, 128).astype(np.float32)
    ix.nprobe = 8
    Dref, Iref = ix.search(xq, 32)
    D, I, _ = Pipeline(emu, ix, layout, 3).search(xq, 32, 8, 8)
    assert np.array_equal(I, Iref) and (I[:, 21:] == -1).all()
    assert D[:, :21].tobytes() == Dref[:, :21].tobytes()


def test_emulated_kernels_under_address_sanitizer(tmp_path):
    """CPU stand-in for compute-sanitizer memcheck: every emulated kernel on random ragged shards
    with exact-size allocations, under ASan + UBSan (tests/emu/emu_memcheck.cpp)"""
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "emu_memcheck")
    csrc = os.path.join(ROOT, "distributed_faiss_b200", "csrc")
    cmd = [cxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-fno-omit-frame-pointer", "-ffp-contract=off", "-DDFX_EMU", "-Wno-unknown-pragmas", "-Wno-attributes",
           "-I", os.path.join(EMU_DIR, "shim"), "-I", csrc, "-I", EMU_DIR, "-o", exe,
           os.path.join(EMU_DIR, "emu_memcheck.cpp")]
    probe = subprocess.run(cmd, capture_output=True, text=True)
    if probe.returncode != 0 and "sanitize" in probe.stderr and "cannot find" in probe.stderr:
        pytest.skip("sanitizer runtime not installed")
    assert probe.returncode == 0, probe.stderr[-2000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_stack_use_after_return=0")
    run = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=900)
    assert run.returncode == 0 and "memcheck ok" in run.stdout, (run.stdout + run.stderr)[-3000:]
