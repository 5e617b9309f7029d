"""Independent float64 numpy restatement (test infrastructure -- NOT product code).

A second, deliberately naive restatement of the same faiss semantics as
oracle/dfx_oracle.c, written without any of its decompositions: every distance
is evaluated directly in float64 from the *reconstructed* vector
(flat: the row; IVF-Flat: the row; IVF-PQ: centroid + PQ codewords;
IVF-SQ: centroid + fp16 residual).  faiss itself is unavailable (see the header
of dfx_oracle.c), so the two restatements cross-check each other: ids must agree
wherever float64 distances are separated by more than the float32 rounding of the
canonical arithmetic, and distances must agree to the stated tolerance.

Reference anchors: distributed_faiss/index.py:25-48,63-68,94 (builders),
index.py:241-270 (search), client.py:29-54,265-310 (merge).
"""
from __future__ import annotations

import numpy as np

METRIC_IP = 0
METRIC_L2 = 1
FLT_MAX = float(np.finfo(np.float32).max)


def _values(metric, q, X):
    """ranking values (smaller is better) of rows X for one query q, float64."""
    q = q.astype(np.float64)
    X = X.astype(np.float64)
    if metric == METRIC_IP:
        return -(X @ q)
    diff = X - q[None, :]
    return np.einsum("ij,ij->i", diff, diff)


def _topk(vals, ids, k, metric):
    order = np.lexsort((ids, vals))[:k]
    D = np.full(k, -FLT_MAX if metric == METRIC_IP else FLT_MAX, dtype=np.float64)
    I = np.full(k, -1, dtype=np.int64)
    n = len(order)
    D[:n] = -vals[order] if metric == METRIC_IP else vals[order]
    I[:n] = ids[order]
    return D, I


def flat_search(metric, xb, xq, k):
    nq = xq.shape[0]
    D = np.empty((nq, k))
    I = np.empty((nq, k), dtype=np.int64)
    ids = np.arange(xb.shape[0], dtype=np.int64)
    for i in range(nq):
        D[i], I[i] = _topk(_values(metric, xq[i], xb), ids, k, metric)
    return D, I


def coarse(metric, cent, xq, nprobe):
    keys = np.empty((xq.shape[0], nprobe), dtype=np.int64)
    ids = np.arange(cent.shape[0], dtype=np.int64)
    for i in range(xq.shape[0]):
        v = _values(metric, xq[i], cent)
        keys[i] = np.lexsort((ids, v))[:nprobe]
    return keys


def half_to_float64(h):
    return np.ascontiguousarray(h, dtype=np.uint16).view(np.float16).astype(np.float64)


def reconstruct_all(state):
    """float64 reconstruction of every stored vector, in list-sorted storage order."""
    kind = state["kind"]
    if kind == "flat":
        return state["xb"].astype(np.float64)
    if kind == "ivf_flat":
        return state["vecs"].astype(np.float64)
    list_of = np.repeat(np.arange(state["nlist"]), np.diff(state["list_off"]))
    base = state["centroids"].astype(np.float64)[list_of]
    if kind == "ivf_sq":
        return base + half_to_float64(state["codes16"])
    if kind == "ivf_pq":
        M, ksub = state["M"], state["ksub"]
        cb = state["codebooks"].astype(np.float64).reshape(M, ksub, -1)
        dsub = cb.shape[2]
        out = base.copy()
        for m in range(M):
            out[:, m * dsub:(m + 1) * dsub] += cb[m][state["codes"][:, m]]
        return out
    raise ValueError(kind)


def ivf_search(state, xq, nprobe, k, keys=None):
    """IVF search over a state dict (oracle.get_state() layout). If `keys` is given the
    probe lists are taken from it (so scan arithmetic can be checked in isolation)."""
    kind = state["kind"]
    metric = state["metric"] if kind == "ivf_flat" else METRIC_L2
    cmetric = state["metric"] if kind == "ivf_flat" else state["coarse_metric"]
    if keys is None:
        keys = coarse(cmetric, state["centroids"], xq, nprobe)
    recon = reconstruct_all(state)
    off = state["list_off"]
    nq = xq.shape[0]
    D = np.empty((nq, k))
    I = np.empty((nq, k), dtype=np.int64)
    for i in range(nq):
        rows = np.concatenate([np.arange(off[l], off[l + 1]) for l in keys[i] if l >= 0] + [np.zeros(0, dtype=np.int64)]).astype(np.int64)
        D[i], I[i] = _topk(_values(metric, xq[i], recon[rows]), state["ids"][rows], k, metric)
    return D, I


def merge(Dall, Iall):
    """ResultHeap semantics (client.py:29-54): keep the k smallest, ascending, pad (FLT_MAX,-1)."""
    S, nq, k = Dall.shape
    outD = np.full((nq, k), FLT_MAX, dtype=np.float32)
    outI = np.full((nq, k), -1, dtype=np.int64)
    for i in range(nq):
        v = Dall[:, i, :].reshape(-1)
        ids = Iall[:, i, :].reshape(-1)
        keep = v < np.float32(FLT_MAX)
        v, ids = v[keep], ids[keep]
        order = np.lexsort((ids, v))[:k]
        outD[i, :len(order)] = v[order]
        outI[i, :len(order)] = ids[order]
    return outD, outI
