"""CPU checks of the experimental lane-per-vector scan (dfx_scan_il2.cu, scan_variant = 2):
the block layout exported by the library and a lane-level emulation of the kernel's algorithm
(tests/emu_scan_il2.py) against the oracle.  The CUDA transcription itself is covered by the
gpu tests in test_gpu_parity.py (scan_variant parametrisation)."""
import ctypes

import numpy as np
import pytest

from oracle import oracle as O
from tests import emu_scan_il2 as emu


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as entry

    entry.build()
    from distributed_faiss_b200 import engine

    return ctypes.CDLL(engine.LIB_PATH)


@pytest.mark.parametrize("layout", [1, 2, 3])
def test_block_layout_is_a_bijection(lib, layout):
    t = emu.il_byte_table(lib, layout)
    assert sorted(t.reshape(-1).tolist()) == list(range(1024))


def test_layout2_matches_its_definition(lib):
    """lane v reads its 32 code bytes as two 16-byte halves; byte t holds subquantizer (t + v) & 31,
    and each half of the warp is one contiguous 512-byte run"""
    t = emu.il_byte_table(lib, 2)
    for v in range(32):
        for step in range(32):
            assert t[v, (step + v) & 31] == (step >> 4) * 512 + v * 16 + (step & 15)


def test_rotated_tree_equals_canonical_tree():
    """the halving tree over the rotated index joins the same operands as oracle pq_sum"""
    rs = np.random.RandomState(0)
    lut = (rs.randn(32, 256) * rs.choice([1e-3, 1.0, 1e3], size=(32, 1))).astype(np.float32)
    codes = rs.randint(0, 256, size=(32, 32)).astype(np.uint8)  # [v][m]
    lutW = emu.wide_table(lut)
    block = np.zeros(1024, dtype=np.uint8)
    for v in range(32):
        for m in range(32):
            t = (m - v) & 31
            block[(t >> 4) * 512 + v * 16 + (t & 15)] = codes[v, m]
    got = emu.block_values(lutW, block, np.zeros(32, np.float32), np.float32(0))
    for v in range(32):
        s = lut[np.arange(32), codes[v]].astype(np.float32)
        off = 16
        while off >= 1:
            s[:off] = (s[:off] + s[off:2 * off]).astype(np.float32)
            off >>= 1
        want = np.float32(0) + (np.float32(0) + s[0])
        assert got[v].tobytes() == np.float32(want).tobytes()


def test_shuffle_networks():
    rs = np.random.RandomState(1)
    for _ in range(50):
        x = rs.randint(0, 1 << 62, size=32).astype(np.uint64)
        x[rs.rand(32) < 0.2] = emu.NONE
        s = emu.sort32_asc(x.copy())
        assert np.array_equal(s, np.sort(x))
        kept = np.sort(rs.randint(0, 1 << 62, size=32).astype(np.uint64))
        kept[rs.randint(0, 33):] = emu.NONE
        m = emu.merge_sorted(kept, s)
        assert np.array_equal(m, np.sort(np.concatenate([kept, x]))[:32])


def _index(n, d, nlist, dup=0, seed=0):
    rs = np.random.RandomState(seed)
    x = rs.randn(n, d).astype(np.float32)
    if dup:  # exact duplicates: equal distances, order decided by the id
        x[-dup:] = x[:dup]
    ix = O.make_index("ivf_pq", d, O.METRIC_L2, nlist=nlist, M=32)
    ix.train_niter = 4
    ix.train(x[:2000])
    ix.add(x)
    return ix, rs


@pytest.mark.parametrize("k,nprobe,G,seed", [(10, 8, 8, 0), (1, 4, 2, 1), (32, 5, 16, 2), (7, 3, 1, 3)])
def test_emulated_kernel_equals_oracle(lib, k, nprobe, G, seed):
    ix, rs = _index(5000, 64, 12, dup=300, seed=seed)
    xq = rs.randn(4, 64).astype(np.float32)
    xq[0] = ix.reconstruct_rows([5])[0]  # a query sitting on a duplicated database point
    ix.nprobe = nprobe
    Dref, Iref = ix.search(xq, k)
    D, I, stats = emu.search(ix, lib, xq, k, nprobe, G=G, rng=np.random.RandomState(seed))
    assert np.array_equal(I, Iref)
    assert D.tobytes() == Dref.tobytes()
    assert stats["flushes"] > 0


def test_emulated_kernel_short_lists_and_few_hits(lib):
    """lists shorter than a block, empty lists, fewer than k hits"""
    rs = np.random.RandomState(5)
    x = rs.randn(2000, 64).astype(np.float32)
    ix = O.make_index("ivf_pq", 64, O.METRIC_L2, nlist=12, M=32)
    ix.train_niter = 4
    ix.train(x)
    ix.add(x[:23])
    xq = rs.randn(3, 64).astype(np.float32)
    ix.nprobe = 12
    Dref, Iref = ix.search(xq, 32)
    D, I, _ = emu.search(ix, lib, xq, 32, 12, G=12)
    assert np.array_equal(I, Iref) and (I[:, 23:] == -1).all()
    assert D[:, :23].tobytes() == Dref[:, :23].tobytes()
