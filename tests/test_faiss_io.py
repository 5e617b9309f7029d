"""faiss on-disk format (distributed_faiss_b200/faiss_io.py).

faiss is not available here, so these are self-consistency checks: every index kind round-trips
through the file bit for bit, the byte layout matches the documented one at fixed offsets, and a
shard saved in that format answers identically after `Index.from_storage_dir`."""
import os
import struct

import numpy as np
import pytest

from distributed_faiss_b200 import faiss_io
from distributed_faiss_b200.index import Index
from distributed_faiss_b200.index_cfg import IndexCfg
from distributed_faiss_b200.index_state import IndexState
from oracle import oracle as O
from tests.oracle_engine import oracle_engine_factory


def _built(kind, d=32, n=700, nlist=9, M=8, metric=O.METRIC_L2, seed=0):
    rs = np.random.RandomState(seed)
    x = rs.randn(n, d).astype(np.float32)
    ix = O.make_index(kind, d, metric=metric, nlist=nlist, M=M)
    if kind != "flat":
        ix.train_niter = 3
        ix.train(x)
    ix.add(x)
    return ix, rs


@pytest.mark.parametrize("kind,metric", [("flat", O.METRIC_IP), ("ivf_flat", O.METRIC_L2), ("ivf_flat", O.METRIC_IP),
                                         ("ivf_pq", O.METRIC_L2), ("ivf_pq", O.METRIC_IP), ("ivf_sq", O.METRIC_L2)])
def test_round_trip_is_bit_exact(tmp_path, kind, metric):
    ix, rs = _built(kind, metric=metric)
    st = ix.get_state()
    path = str(tmp_path / "index.faiss")
    faiss_io.write_index(st, path, nprobe=5)
    back, nprobe = faiss_io.read_index(path)
    assert back["kind"] == kind
    if kind != "flat":
        assert nprobe == 5
    for key, val in st.items():
        if key == "tvals":  # not part of the faiss format: recomputed on import
            continue
        if isinstance(val, np.ndarray):
            assert np.asarray(back[key]).tobytes() == np.ascontiguousarray(val).tobytes(), key
        else:
            assert back[key] == val, key
    ix2 = O.make_index(kind, 32, metric=metric, nlist=9, M=8)
    ix2.set_state(back)
    xq = rs.randn(6, 32).astype(np.float32)
    if kind != "flat":
        ix.nprobe = ix2.nprobe = 4
    D1, I1 = ix.search(xq, 7)
    D2, I2 = ix2.search(xq, 7)
    assert np.array_equal(I1, I2) and D1.tobytes() == D2.tobytes()


def test_byte_layout_of_an_ivfpq_file(tmp_path):
    ix, _ = _built("ivf_pq", d=32, n=300, nlist=4, M=8)
    path = str(tmp_path / "i.faiss")
    faiss_io.write_index(ix.get_state(), path, nprobe=3)
    raw = open(path, "rb").read()
    assert raw[:4] == b"IwPQ"
    d, ntotal, dm1, dm2, trained, metric = struct.unpack_from("<iqqqBi", raw, 4)
    assert (d, ntotal, dm1, dm2, trained, metric) == (32, 300, 1 << 20, 1 << 20, 1, 1)
    off = 4 + 33
    assert struct.unpack_from("<QQ", raw, off) == (4, 3)                     # nlist, nprobe
    off += 16
    assert raw[off:off + 4] == b"IxF2"                                        # quantizer
    qd, qn = struct.unpack_from("<iq", raw, off + 4)
    assert (qd, qn) == (32, 4)
    off += 4 + 33
    assert struct.unpack_from("<Q", raw, off)[0] == 4 * 32                    # centroid floats
    off += 8 + 4 * 32 * 4
    assert raw[off] == 0 and struct.unpack_from("<Q", raw, off + 1)[0] == 0   # no direct map
    off += 9
    assert raw[off] == 1 and struct.unpack_from("<Q", raw, off + 1)[0] == 8   # by_residual, code_size
    off += 9
    assert struct.unpack_from("<QQQ", raw, off) == (32, 8, 8)                 # pq.d, pq.M, pq.nbits
    off += 24
    assert struct.unpack_from("<Q", raw, off)[0] == 8 * 256 * 4
    off += 8 + 8 * 256 * 4 * 4
    assert raw[off:off + 4] == b"ilar" and struct.unpack_from("<QQ", raw, off + 4) == (4, 8)
    assert raw[off + 20:off + 24] == b"full"
    sizes = struct.unpack_from("<Q4Q", raw, off + 24)
    assert sizes[0] == 4 and sum(sizes[1:]) == 300
    assert len(raw) == off + 24 + 40 + 300 * (8 + 8)


def test_sparse_list_table_and_empty_index(tmp_path):
    ix, rs = _built("ivf_flat", n=600, nlist=9)
    st = ix.get_state()
    keep = np.zeros(len(st["ids"]), dtype=bool)
    keep[st["list_off"][2]:st["list_off"][3]] = True                           # only list 2 survives
    lo = np.zeros(10, dtype=np.int64)
    lo[3:] = keep.sum()
    st2 = dict(st, list_off=lo, ids=st["ids"][keep], vecs=st["vecs"][keep])
    path = str(tmp_path / "s.faiss")
    faiss_io.write_index(st2, path)
    assert b"sprs" in open(path, "rb").read()
    back, _ = faiss_io.read_index(path)
    assert np.array_equal(back["list_off"], lo) and np.array_equal(back["vecs"], st2["vecs"])
    st3 = dict(st, list_off=np.zeros(10, np.int64), ids=st["ids"][:0], vecs=st["vecs"][:0])
    faiss_io.write_index(st3, path)
    back, _ = faiss_io.read_index(path)
    assert len(back["ids"]) == 0 and back["vecs"].shape == (0, 32)


def test_rejects_other_index_types(tmp_path):
    path = str(tmp_path / "h.faiss")
    with open(path, "wb") as f:
        f.write(b"IHNf" + b"\0" * 64)
    with pytest.raises(faiss_io.FaissFormatError, match="IHNf"):
        faiss_io.read_index(path)
    with open(path, "wb") as f:
        f.write(b"IwPQ\x20\0")
    with pytest.raises(faiss_io.FaissFormatError, match="end of file"):
        faiss_io.read_index(path)


@pytest.mark.parametrize("builder", ["flat", "ivf_simple", "knnlm", "ivfsq"])
def test_shard_saved_as_index_faiss_reloads(tmp_path, builder):
    """Index.save with index_format="faiss" writes the reference's file name; from_storage_dir
    picks it up (reference index.py:284-344) and the reloaded shard answers identically."""
    rs = np.random.RandomState(3)
    d = 32
    cfg = IndexCfg(index_builder_type=builder, dim=d, centroids=6, metric="l2", nprobe=3, train_num=400,
                   index_storage_dir=str(tmp_path), code_size=8, index_format="faiss")
    ix = Index(cfg, engine_factory=oracle_engine_factory)
    x = rs.randn(500, d).astype(np.float32)
    ix.add_batch(x, [("m", i) for i in range(500)], train_async_if_triggered=False)
    ix.train()
    deadline = __import__("time").time() + 30
    while ix.get_state() != IndexState.TRAINED and __import__("time").time() < deadline:
        __import__("time").sleep(0.01)
    ix.set_nprobe(3) if builder != "flat" else None
    assert ix.save()
    assert os.path.exists(tmp_path / "index.faiss") and not os.path.exists(tmp_path / "index.dfx.npz")
    q = rs.randn(5, d).astype(np.float32)
    D1, m1, _ = ix.search(q, 4)
    cfg2 = IndexCfg.from_json(str(tmp_path / "cfg.json"))
    ix2 = Index.from_storage_dir(str(tmp_path), cfg2, engine_factory=oracle_engine_factory)
    assert ix2 is not None and ix2.get_state() == IndexState.TRAINED
    ix2.set_nprobe(3) if builder != "flat" else None
    D2, m2, _ = ix2.search(q, 4)
    assert D1.tobytes() == D2.tobytes() and m1 == m2
