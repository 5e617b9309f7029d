"""world_size-2 test of the NCCL data-plane LOGIC on CPU (gloo): broadcast of the query,
per-rank shard search, all-gather, merge -- with the oracle as the numeric backend
(tests/oracle_engine.py).  The CUDA/NCCL form of the same flow is tested in test_gpu_api.py."""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_data():
    rs = np.random.RandomState(7)
    d, n_per, nshards = 32, 500, 4
    xs = [rs.rand(n_per, d).astype(np.float32) for _ in range(nshards)]
    xq = rs.rand(9, d).astype(np.float32)
    return d, xs, xq


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from distributed_faiss_b200 import spmd
    from oracle import oracle as O
    from tests.oracle_engine import OracleBackend

    r, _, w = spmd.init_process_group_from_env(backend="gloo")
    d, xs, xq = _make_data()
    per_rank = len(xs) // w
    shards, tables = [], []
    for j in range(per_rank):
        s = r * per_rank + j                       # rank-major global shard order
        ix = O.make_index("flat", d, metric=O.METRIC_IP)
        ix.add(xs[s])
        shards.append(ix)
        tables.append(torch.arange(s * 1000, s * 1000 + xs[s].shape[0], dtype=torch.int64))
    group = spmd.ShardGroup(shards, tables, backend=OracleBackend(), device=torch.device("cpu"))
    assert group.num_shards == len(xs) and group.get_ntotal() == sum(x.shape[0] for x in xs)
    # only rank 0 holds the real query: the collective broadcasts it
    x_t = torch.from_numpy(xq.copy()) if r == 0 else torch.zeros(xq.shape, dtype=torch.float32)
    D, I = group.search(x_t, 5, maximize=True, src=0)
    Dh, Ih = group.search_host(xq if r == 0 else np.zeros_like(xq), 5, maximize=True, src=0)
    assert np.array_equal(D.numpy(), Dh) and np.array_equal(I.numpy(), Ih)
    np.savez(os.path.join(outdir, f"rank{r}.npz"), D=D.numpy(), I=I.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_process():
    from oracle import oracle as O

    world, port = 2, _free_port()
    with tempfile.TemporaryDirectory() as outdir:
        mp.spawn(_worker, args=(world, port, outdir), nprocs=world, join=True)
        res = [np.load(os.path.join(outdir, f"rank{r}.npz")) for r in range(world)]
    # every rank ends with the same merged answer
    assert np.array_equal(res[0]["D"], res[1]["D"]) and np.array_equal(res[0]["I"], res[1]["I"])
    # and it is what ONE process would compute over all four shards (reference client semantics:
    # negated scores for "dot", ascending)
    d, xs, xq = _make_data()
    Ds, Is = [], []
    for s, x in enumerate(xs):
        D, I = O.flat_search(O.METRIC_IP, x, xq, 5)
        Ds.append(-D)
        Is.append(I + s * 1000)
    Dref, Pref = O.merge(np.stack(Ds), np.arange(len(xs) * xq.shape[0] * 5, dtype=np.int64).reshape(len(xs), -1, 5))
    Iref = np.stack(Is).reshape(-1)[Pref]
    assert np.array_equal(res[0]["D"], Dref) and np.array_equal(res[0]["I"], Iref)
    # ... which equals the unsharded search
    Dall, Iall = O.flat_search(O.METRIC_IP, np.concatenate(xs), xq, 5)
    assert np.array_equal(-Dall, Dref)
    assert np.array_equal((Iref // 1000) * 500 + Iref % 1000, Iall)
