"""IndexClient.search over the collective data plane (spmd.SearchPlane), world_size 2 on CPU (gloo).

Two processes, two IndexServers each (oracle-backed engines, tests/oracle_engine.py), the
client in the process of plane rank 0.  The same client object answers every query twice --
through the plane (header + query broadcast, per-rank shard search, ONE all-gather, merge)
and, after `detach_plane()`, through the reference's socket fan-out -- and the answers must be
identical: scores bit for bit, metadata equal (reference contract: client.py:200-210,
213-263, 265-310).  The CUDA/NCCL form of the same flow is tested in test_gpu_api.py."""
import os
import socket
import sys
import tempfile
import threading
import time

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_ports(n):
    socks = [socket.socket() for _ in range(n)]
    try:
        for s in socks:
            s.bind(("127.0.0.1", 0))
        return [s.getsockname()[1] for s in socks]
    finally:
        for s in socks:
            s.close()


def _wait_listening(ports, timeout=30.0):
    t0 = time.time()
    for p in ports:
        while True:
            try:
                socket.create_connection(("127.0.0.1", p), timeout=1.0).close()
                break
            except OSError:
                assert time.time() - t0 < timeout
                time.sleep(0.05)


def _worker(rank, world, port, ports, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from distributed_faiss_b200 import rpc, spmd
    from distributed_faiss_b200.client import IndexClient, MetaRows
    from distributed_faiss_b200.index_cfg import IndexCfg
    from distributed_faiss_b200.index_state import IndexState
    from distributed_faiss_b200.server import IndexServer
    from tests.oracle_engine import OracleBackend, oracle_engine_factory, oracle_merge
    from distributed_faiss_b200 import client as client_mod

    client_mod.ResultHeap.merge_backend = staticmethod(oracle_merge)   # socket path's merge, CPU double
    spmd.init_process_group_from_env(backend="gloo")
    s_loc = len(ports) // world
    servers = []
    for j in range(s_loc):
        sr = rank * s_loc + j
        srv = IndexServer(sr, os.path.join(outdir, "store"), engine_factory=oracle_engine_factory)
        threading.Thread(target=srv.start_blocking, args=(ports[sr],), daemon=True).start()
        servers.append(srv)
    plane = spmd.SearchPlane(servers, backend=OracleBackend(), device=torch.device("cpu"))
    assert plane.num_servers == len(ports)
    dist.barrier()
    if rank != 0:
        plane.serve_forever()
        dist.barrier()
        dist.destroy_process_group()
        return

    _wait_listening(ports)
    disc = os.path.join(outdir, "servers.txt")
    with open(disc, "w") as fh:
        fh.write(f"{len(ports)}\n" + "".join(f"127.0.0.1,{p}\n" for p in ports))
    client = IndexClient(disc)
    assert client.plane is plane
    rs = np.random.RandomState(3)
    d = 16

    def wait_trained(index_id):
        t0 = time.time()
        while client.get_state(index_id) != IndexState.TRAINED:
            assert time.time() - t0 < 60
            time.sleep(0.02)

    def both_ways(fn):
        client.plane = plane
        a = fn()
        client.detach_plane()
        b = fn()
        client.plane = plane
        return a, b

    # ---- integer metadata (scripts/load_data.py convention): the device fast path, lazy rows
    client.create_index("ints", IndexCfg(index_builder_type="flat", dim=d, metric="dot", train_num=40))
    nxt = 0
    for _ in range(12):
        n = int(rs.randint(20, 60))
        client.add_index_data("ints", rs.rand(n, d).astype(np.float32), list(range(nxt, nxt + n)))
        nxt += n
    client.sync_train("ints")
    wait_trained("ints")
    assert client.get_ntotal("ints") == nxt
    xq = rs.rand(7, d).astype(np.float32)
    (Dp, Mp), (Ds, Ms) = both_ways(lambda: client.search(xq, 5, "ints"))
    assert isinstance(Mp, MetaRows) and isinstance(Ms, list)
    assert np.array_equal(Dp, Ds) and Mp == Ms and Mp[3] == Ms[3] and len(Mp) == 7
    assert (np.diff(Dp, axis=1) >= 0).all() and (Dp < 0).all()      # negated scores, ascending (quirk B5)
    # with embeddings: winners are decoded by their owners and summed to the client rank
    (Dp, Mp, Ep), (Ds, Ms, Es) = both_ways(lambda: client.search(xq, 5, "ints", return_embeddings=True))
    assert np.array_equal(Dp, Ds) and Mp == Ms
    assert all(np.array_equal(Ep[q][j], Es[q][j]) for q in range(7) for j in range(5))
    # k larger than the number of stored vectors of some shards: padded results, None metadata
    (Dp, Mp), (Ds, Ms) = both_ways(lambda: client.search(xq[:2], 300, "ints"))
    assert np.array_equal(Dp, Ds) and Mp == Ms

    # ---- object metadata (tuples): exchange ids through the collective, objects fetched from owners
    client.create_index("objs", IndexCfg(index_builder_type="flat", dim=d, metric="l2", train_num=30))
    nxt = 0
    for _ in range(9):
        n = int(rs.randint(20, 50))
        meta = [(i, "even" if i % 2 == 0 else "odd", f"doc{i}") for i in range(nxt, nxt + n)]
        if nxt == 0:
            meta[3] = (3,)           # too short for filter_pos 1: never kept by the filter
            meta[5] = None           # no metadata at all
        client.add_index_data("objs", rs.rand(n, d).astype(np.float32), meta)
        nxt += n
    client.sync_train("objs")
    wait_trained("objs")
    (Dp, Mp), (Ds, Ms) = both_ways(lambda: client.search(xq, 6, "objs"))
    assert np.array_equal(Dp, Ds) and Mp == Ms and isinstance(Mp[0][0], tuple)
    for fv in ("even", "odd", "absent"):
        (Sp, Mp), (Ss, Ms) = both_ways(lambda: client.search_with_filter(xq, 4, "objs", filter_pos=1, filter_value=fv))
        assert Mp == Ms and len(Sp) == len(Ss) == 7
        assert all(np.array_equal(a, b) for a, b in zip(Sp, Ss))
        assert all(m[1] != fv for row in Mp for m in row)
    # data added after the first search: metadata kind / filter columns are refreshed
    client.add_index_data("objs", rs.rand(30, d).astype(np.float32), [(10_000 + i, "even", "late") for i in range(30)])
    wait_trained("objs")
    (Sp, Mp), (Ss, Ms) = both_ways(lambda: client.search_with_filter(xq, 4, "objs", filter_pos=1, filter_value="odd"))
    assert Mp == Ms and all(np.array_equal(a, b) for a, b in zip(Sp, Ss))

    # ---- error contract: an untrained / unknown index raises ServerException on the client
    client.create_index("cold", IndexCfg(index_builder_type="flat", dim=d, train_num=10_000))
    client.add_index_data("cold", rs.rand(10, d).astype(np.float32), list(range(10)))
    with pytest.raises(rpc.ServerException, match="not trained"):
        client.search(xq, 3, "cold")
    client.cfg = IndexCfg(metric="l2")
    with pytest.raises(rpc.ServerException, match="no index"):
        client.search(xq, 3, "never-created")
    # the plane survives errors
    client.cfg = IndexCfg(metric="dot")
    D2, M2 = client.search(xq, 5, "ints")
    assert D2.shape == (7, 5)

    # ---- device timing helper used by bench.py (max over ranks)
    plane.timer_start()
    client.search(xq, 5, "ints")
    assert plane.timer_stop() > 0.0
    np.savez(os.path.join(outdir, "done.npz"), ok=np.array([1]))
    client.close()
    plane.stop()
    dist.barrier()
    dist.destroy_process_group()


def test_client_search_over_collective_equals_socket_fanout():
    world = 2
    ports = _free_ports(5)
    with tempfile.TemporaryDirectory() as outdir:
        mp.spawn(_worker, args=(world, ports[0], ports[1:], outdir), nprocs=world, join=True)
        assert os.path.exists(os.path.join(outdir, "done.npz"))
