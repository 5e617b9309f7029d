"""CPU tests of the oracle (oracle/dfx_oracle.c): pinned by the reference's golden vectors where
the reference has any (the cross-shard merge), cross-checked against the independent float64
restatement everywhere else, and frozen by tests/golden/oracle_small.npz."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_numpy as R
from tests.conftest import clustered

HERE = os.path.dirname(os.path.abspath(__file__))
FLT_MAX = np.finfo(np.float32).max


def test_merge_reference_golden_vectors():
    """reference tests/test_integration.py:181-203 (test_result_aggregation)"""
    gold = json.load(open(os.path.join(HERE, "golden", "merge_golden.json")))
    D = np.array(gold["shard_D"], dtype=np.float32)
    meta = [m for s in gold["shard_meta"] for row in s for m in row]
    S, nq, k = D.shape
    pos = np.arange(S * nq * k, dtype=np.int64).reshape(S, nq, k)
    for name, Din in (("minimize", D), ("maximize", -D)):
        for impl in (O.merge, R.merge):
            outD, outP = impl(Din, pos)
            assert [meta[p] for p in outP[0]] == gold[name]["meta"][0]
            assert np.array_equal(outD[0], np.array(gold[name]["D"][0], dtype=np.float32))
    # the assertions of the reference test itself
    _, pmin = O.merge(D, pos)
    _, pmax = O.merge(-D, pos)
    imin, imax = [meta[p] for p in pmin[0]], [meta[p] for p in pmax[0]]
    assert imax != imin and imin[0] == 0 and imax[0] == 1 and 0 in imin


def test_merge_semantics():
    rs = np.random.RandomState(0)
    S, nq, k = 5, 40, 7
    D = np.sort(rs.rand(S, nq, k).astype(np.float32), axis=2)
    D[rs.rand(S, nq, k) < 0.2] = FLT_MAX           # missing hits are never admitted
    D[1, :, :3] = D[0, :, :3]                        # ties: earlier shard (smaller position) wins
    pos = np.arange(S * nq * k, dtype=np.int64).reshape(S, nq, k)
    Do, Po = O.merge(D, pos)
    Dr, Pr = R.merge(D, pos)
    assert np.array_equal(Do, Dr) and np.array_equal(Po, Pr)
    assert (np.diff(Do, axis=1) >= 0).all()
    assert ((Po == -1) == (Do == FLT_MAX)).all()
    # fewer than k real entries in total -> (FLT_MAX, -1) padding
    D2 = np.full((2, 1, 4), FLT_MAX, dtype=np.float32)
    D2[0, 0, 0] = 1.5
    Dm, Pm = O.merge(D2, np.arange(8, dtype=np.int64).reshape(2, 1, 4))
    assert Dm.tolist() == [[1.5, FLT_MAX, FLT_MAX, FLT_MAX]] and Pm.tolist() == [[0, -1, -1, -1]]


def test_half_conversions_match_ieee():
    h = np.arange(65536, dtype=np.uint16)
    f = O.half_to_float(h)
    g = h.view(np.float16).astype(np.float32)
    ok = ~np.isnan(g)
    assert np.array_equal(f.view(np.uint32)[ok], g.view(np.uint32)[ok])
    rs = np.random.RandomState(1)
    x = (rs.randn(200000) * rs.choice([1e-8, 1e-6, 1e-3, 1, 100, 6e4, 1e6], 200000)).astype(np.float32)
    with np.errstate(over="ignore"):
        ref = x.astype(np.float16).view(np.uint16)
    assert np.array_equal(O.float_to_half(x), ref)


@pytest.mark.parametrize("metric", [O.METRIC_IP, O.METRIC_L2])
def test_flat_matches_float64(metric):
    rs = np.random.RandomState(2)
    xb = rs.rand(3000, 128).astype(np.float32)
    xq = rs.rand(9, 128).astype(np.float32)
    D, I = O.flat_search(metric, xb, xq, 10)
    D64, I64 = R.flat_search(metric, xb, xq, 10)
    assert np.array_equal(I, I64)
    assert np.abs(D - D64).max() <= 1e-4 * np.abs(D64).max()
    # sharded == unsharded, exactly: the metamorphic test of reference test_integration.py:205-265
    cut = 1234
    D1, I1 = O.flat_search(metric, xb[:cut], xq, 10)
    D2, I2 = O.flat_search(metric, xb[cut:], xq, 10)
    sign = -1.0 if metric == O.METRIC_IP else 1.0
    Dm, Pm = O.merge(np.stack([sign * D1, sign * D2]), np.stack([I1, I2 + cut]))
    assert np.array_equal(sign * Dm, D) and np.array_equal(Pm, I)


CASES = [("ivf_flat", O.METRIC_L2, 0), ("ivf_flat", O.METRIC_IP, 0), ("ivf_pq", O.METRIC_L2, 16),
         ("ivf_pq", O.METRIC_IP, 16), ("ivf_sq", O.METRIC_L2, 0)]


@pytest.mark.parametrize("kind,metric,M", CASES)
def test_ivf_matches_float64(kind, metric, M):
    rs = np.random.RandomState(3)
    d, nlist, n = 64, 16, 4000
    xb = clustered(rs, n, d, ncl=24)
    ix = O.make_index(kind, d, metric=metric, nlist=nlist, M=M)
    ix.train_niter = 8
    ix.train(xb)
    ix.add(xb[:1500]); ix.add(xb[1500:])           # two adds: ids stay ascending inside each list
    assert ix.ntotal == n
    st = ix.get_state()
    lo = st["list_off"]
    assert all((np.diff(st["ids"][lo[l]:lo[l + 1]]) > 0).all() for l in range(nlist))
    xq = clustered(rs, 12, d, ncl=24)
    scale = (xq.astype(np.float64) ** 2).sum(1, keepdims=True) + (xb.astype(np.float64) ** 2).sum(1).max()
    for nprobe in (1, 4, nlist):
        ix.nprobe = nprobe
        D, I = ix.search(xq, 10)
        keys, _ = O.coarse(ix.coarse_metric, ix.centroids, xq, nprobe)
        D64, I64 = R.ivf_search(st, xq, nprobe, 10, keys=keys)
        assert (np.abs(D - D64) <= 1e-4 * scale).all()
        assert (I == I64).mean() > 0.98
        if nprobe == nlist:                         # full nprobe: the coarse stage cannot matter
            D64b, I64b = R.ivf_search(st, xq, nprobe, 10)
            assert np.array_equal(I64, I64b)


def test_ivf_edge_cases():
    rs = np.random.RandomState(4)
    d, nlist = 32, 8
    xb = clustered(rs, 500, d, ncl=3, sigma=0.05)
    xb[40:50] = xb[7]                               # exact duplicates -> exact ties, broken by id
    for kind, M in (("ivf_flat", 0), ("ivf_pq", 8), ("ivf_sq", 0)):
        ix = O.make_index(kind, d, metric=O.METRIC_L2, nlist=nlist, M=M)
        ix.train(xb)
        ix.add(xb[:60])
        ix.nprobe = 100                             # nprobe > nlist is clamped
        D, I = ix.search(xb[:5], 80)                # k > ntotal -> (-1, FLT_MAX) padding at the tail
        assert (I[:, 60:] == -1).all() and (D[:, 60:] == FLT_MAX).all()
        assert (I[:, :60] >= 0).all() and (np.diff(D[:, :60], axis=1) >= 0).all()
        row = I[0, :60].tolist()
        dup = [i for i in row if i in (7, *range(40, 50))]
        assert dup == sorted(dup)                   # ties appear in ascending id order
        ix.nprobe = 1
        D1, I1 = ix.search(xb[:5], 4)
        assert D1.shape == (5, 4)


def test_golden_fixture_is_reproduced():
    z = np.load(os.path.join(HERE, "golden", "oracle_small.npz"))
    xb, xq, k = z["xb"], z["xq"], int(z["k"])
    for metric, name in ((O.METRIC_IP, "ip"), (O.METRIC_L2, "l2")):
        D, I = O.flat_search(metric, xb, xq, k)
        assert np.array_equal(D, z[f"flat_{name}_D"]) and np.array_equal(I, z[f"flat_{name}_I"])
    for kind, kw in (("ivf_flat", {}), ("ivf_pq", {"M": 8}), ("ivf_sq", {})):
        ix = O.make_index(kind, xb.shape[1], metric=O.METRIC_L2, nlist=8, **kw)
        st = {key[len(kind) + 7:]: z[key] for key in z.files if key.startswith(kind + "_state_")}
        ix.set_state(st)
        for nprobe in (2, 8):
            ix.nprobe = nprobe
            D, I = ix.search(xq, k)
            assert np.array_equal(D, z[f"{kind}_np{nprobe}_D"]) and np.array_equal(I, z[f"{kind}_np{nprobe}_I"])


def test_pq_decomposition_terms():
    """d(q, c+p) = |q-c|^2 + sum(|p|^2 + 2<c,p>) - 2<q,p>  (faiss precomputed-table form)"""
    rs = np.random.RandomState(5)
    d, M = 32, 8
    ix = O.make_index("ivf_pq", d, metric=O.METRIC_L2, nlist=4, M=M)
    xb = clustered(rs, 800, d, ncl=6)
    ix.train(xb); ix.add(xb)
    q = xb[3] + 0.1
    lut = ix.query_lut(q)
    lo = ix.list_of_rows()
    recon = R.reconstruct_all(ix.get_state())
    for pos in (0, 17, 555):
        l = lo[pos]
        dis0 = O.warp_dot(q, ix.centroids[l], 1)
        s = sum(lut[m, ix.codes[pos, m]] for m in range(M))
        approx = dis0 + ix.tvals[pos] + s
        exact = ((q.astype(np.float64) - recon[pos]) ** 2).sum()
        assert abs(approx - exact) <= 1e-4 * (1 + (q ** 2).sum() + (recon[pos] ** 2).sum())


def test_cpu_baseline_matches_oracle(oracle_lib):
    """the TIMED CPU baseline (oracle/cpu_ivfpq.c: sgemm coarse + threaded scan, free summation
    order) against the bit-exact checker: same probe lists, distances within 1e-4 relative, ids
    equal wherever the oracle's distances are separated by more than that"""
    from oracle import cpu_baseline as CB
    from oracle import oracle as O
    from tests.conftest import clustered

    rs = np.random.RandomState(5)
    d, nlist, M, n = 128, 64, 32, 20000
    xb = clustered(rs, n, d, ncl=80)
    o = O.OracleIVFPQ(d, nlist, M, 8, coarse_metric=O.METRIC_L2)
    o.train_niter = 4
    o.train(xb[:6000])
    o.add(xb)
    xq = xb[:300] + 0.01 * rs.randn(300, d).astype(np.float32)
    cb = CB.CpuIVFPQ(o.get_state(), threads=4)
    for nprobe, k in ((8, 10), (1, 1), (nlist, 25)):
        o.nprobe = nprobe
        Do, Io = o.search(xq, k)
        Dc, Ic = cb.search(xq, k, nprobe)
        assert cb.last_ndis == o.last_ndis                      # same probe lists
        fin = Do < 3e38
        assert np.allclose(Dc[fin], Do[fin], rtol=1e-4, atol=1e-4) and ((Ic >= 0) == fin).all()
        gap = np.abs(np.diff(Do, axis=1, append=3.4e38)) > 1e-3 * np.maximum(np.abs(Do), 1.0)
        gap[:, 1:] &= gap[:, :-1]                               # separated from both neighbours
        assert (Ic[gap & fin] == Io[gap & fin]).mean() > 0.999
        assert (Ic == Io).mean() > 0.98
