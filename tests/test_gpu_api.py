"""GPU tests at the API level: the reference's integration scenarios driven through
IndexServer / IndexClient / ShardGroup with the CUDA engine (no test doubles)."""
import os
import socket
import tempfile
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def free_ports(n):
    """n DISTINCT free ports (all sockets are held open until every port is chosen)"""
    socks = [socket.socket() for _ in range(n)]
    try:
        for s in socks:
            s.bind(("", 0))
        return [s.getsockname()[1] for s in socks]
    finally:
        for s in socks:
            s.close()


def free_port():
    return free_ports(1)[0]


def wait_listening(ports, timeout=30.0):
    """block until every server thread accepts connections (start_blocking runs in a thread)"""
    t0 = time.time()
    for p in ports:
        while True:
            try:
                socket.create_connection(("localhost", p), timeout=1.0).close()
                break
            except OSError:
                assert time.time() - t0 < timeout, f"server on port {p} never came up"
                time.sleep(0.05)


def make_client(ports):
    from distributed_faiss_b200.client import IndexClient

    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as fh:
        fh.write(f"{len(ports)}\n")
        for p in ports:
            fh.write(f"localhost,{p}\n")
    try:
        return IndexClient(fh.name)
    finally:
        os.unlink(fh.name)


def wait_trained(client, index_id, timeout=120):
    from distributed_faiss_b200.index_state import IndexState

    t0 = time.time()
    while client.get_state(index_id) != IndexState.TRAINED:
        assert time.time() - t0 < timeout
        time.sleep(0.05)


@pytest.fixture(scope="module")
def cluster():
    from distributed_faiss_b200.server import IndexServer

    dirs = [tempfile.TemporaryDirectory(), tempfile.TemporaryDirectory()]
    ports = free_ports(5)
    ports, sp = ports[:4], ports[4]
    servers = []
    for rank, port in enumerate(ports):
        s = IndexServer(rank, dirs[0].name)
        threading.Thread(target=s.start_blocking, args=(port,), daemon=True).start()
        servers.append(s)
    single = IndexServer(0, dirs[1].name)
    threading.Thread(target=single.start_blocking, args=(sp,), daemon=True).start()
    wait_listening(ports + [sp])
    yield {"ports": ports, "single": sp, "dirs": dirs}
    for s in servers + [single]:
        s.stop()


def test_result_aggregation_on_device():
    """reference tests/test_integration.py:181-203 through IndexClient._aggregate_results -> K6"""
    from distributed_faiss_b200.client import IndexClient

    mock = [(np.array([[12.1, 13.2, 13.3, 14.3]], dtype=np.float32), [[1465, 1460, 443197, 1340]], None),
            (np.array([[8.1, 12.6, 13.1, 17.4]], dtype=np.float32), [[0, 14, 3, 1]], None)]
    D, i_min = IndexClient._aggregate_results(mock, 4, 1, False, False)
    Dmax, i_max = IndexClient._aggregate_results(mock, 4, 1, True, False)
    assert i_min == [[0, 1465, 14, 3]] and i_max == [[1, 1340, 443197, 1460]]
    assert np.array_equal(D, np.array([[8.1, 12.1, 12.6, 13.1]], dtype=np.float32))
    assert np.array_equal(Dmax, np.array([[-17.4, -14.3, -13.3, -13.2]], dtype=np.float32))


def test_sharded_equals_unsharded_exactly(cluster):
    """reference tests/test_integration.py:205-265 with the CUDA engine (flat, d=512, k=5)"""
    from distributed_faiss_b200.index_cfg import IndexCfg

    rs = np.random.RandomState(0)
    d, index_id = 512, "g_same"
    cfg = IndexCfg(index_builder_type="flat", dim=d)
    single = make_client([cluster["single"]])
    multi = make_client(cluster["ports"])
    single.create_index(index_id, cfg)
    multi.create_index(index_id, cfg)
    for _ in range(10):
        n = int(rs.randint(1, 3000))
        emb = rs.rand(n, d).astype(np.float32)
        meta = [f"m{rs.randint(1 << 30)}" for _ in range(n)]
        multi.add_index_data(index_id, emb, meta, False)
        single.add_index_data(index_id, emb, meta, False)
    multi.sync_train(index_id)
    single.sync_train(index_id)
    wait_trained(multi, index_id)
    wait_trained(single, index_id)
    assert multi.get_ntotal(index_id) == single.get_ntotal(index_id)
    q = rs.rand(16, d).astype(np.float32)
    s_aggr, m_aggr = multi.search(q, 5, index_id)
    s_single, m_single = single.search(q, 5, index_id)
    assert (s_aggr == s_single).all() and m_aggr == m_single
    multi.close(); single.close()


def test_knnlm_pipeline_save_load(cluster):
    """knnlm (IVF-PQ) through the API: train, add, set_nprobe, search, save, reload, same answers"""
    from distributed_faiss_b200.index_cfg import IndexCfg

    rs = np.random.RandomState(1)
    d, index_id = 128, "g_knnlm"
    cfg = IndexCfg(index_builder_type="knnlm", dim=d, centroids=32, metric="l2", train_num=3000, code_size=32)
    client = make_client([cluster["single"]])
    client.create_index(index_id, cfg)
    centers = rs.randn(40, d).astype(np.float32)
    for b in range(6):
        x = (centers[rs.randint(0, 40, 1000)] + 0.2 * rs.randn(1000, d)).astype(np.float32)
        client.add_index_data(index_id, x, list(range(b * 1000, (b + 1) * 1000)), False)
    wait_trained(client, index_id)
    assert client.get_ntotal(index_id) == 6000
    client.set_nprobe(index_id, 8)                     # quirk B3: knnlm starts with nprobe 1
    D, meta = client.search(x[:10], 5, index_id)
    assert all(row[0] == 5000 + i for i, row in enumerate(meta))   # each query finds itself first
    D2, meta2, embs = client.search(x[:10], 5, index_id, return_embeddings=True)
    assert np.array_equal(D, D2) and np.asarray(embs).shape == (10, 5, d)
    client.save_index(index_id)
    client.close()
    c2 = make_client([cluster["single"]])
    assert c2.load_index(index_id, cfg)
    c2.set_nprobe(index_id, 8)
    D3, meta3 = c2.search(x[:10], 5, index_id)
    assert np.array_equal(D, D3) and meta == meta3
    c2.close()


def test_shard_group_single_rank_matches_socket_client():
    """the device data plane (ShardGroup, world 1, 4 shards on one GPU) returns what the socket
    client returns for the same shards"""
    import torch
    from distributed_faiss_b200 import engine, spmd
    from distributed_faiss_b200.client import IndexClient

    rs = np.random.RandomState(2)
    d, k = 64, 7
    shards, tables, results, base = [], [], [], 0
    xq = rs.rand(33, d).astype(np.float32)
    for s in range(4):
        x = rs.rand(700 + 50 * s, d).astype(np.float32)
        ix = engine.GpuIndex(engine.KIND_FLAT, d, engine.METRIC_INNER_PRODUCT)
        ix.add(x)
        shards.append(ix)
        tables.append(torch.arange(base, base + x.shape[0], dtype=torch.int64, device="cuda"))
        D, I = ix.search(xq, k)
        results.append((D, (I + base).tolist(), None))
        base += x.shape[0]
    group = spmd.ShardGroup(shards, tables)
    D_dev, I_dev = group.search(torch.from_numpy(xq).cuda(), k, maximize=True)
    D_ref, meta_ref = IndexClient._aggregate_results(results, k, xq.shape[0], True, False)
    assert np.array_equal(D_dev.cpu().numpy(), D_ref) and I_dev.cpu().tolist() == meta_ref
    Dh, Ih = group.search_host(xq, k, maximize=True)
    assert np.array_equal(Dh, D_ref) and Ih.tolist() == meta_ref


def test_shard_group_graph_replay_matches_eager(monkeypatch):
    """latency-bound batches are replayed from a captured CUDA graph (default for nq <= 256, one
    rank): same results as eager launches for fresh inputs, after an nprobe change and after an add
    (the shard generation invalidates the captures); IVF-PQ shards, side streams inside the capture"""
    import torch
    from distributed_faiss_b200 import engine, spmd

    rs = np.random.RandomState(5)
    d, k = 128, 10
    shards, tables, base = [], [], 0
    for s in range(3):
        x = rs.randn(6000, d).astype(np.float32)
        ix = engine.GpuIndex(engine.KIND_IVF_PQ, d, engine.METRIC_L2, nlist=32, pq_m=32)
        ix.set_param("kmeans_niter", 5)
        ix.train(x[:4000])
        ix.add(x)
        shards.append(ix)
        tables.append(torch.arange(base, base + 2 * x.shape[0], dtype=torch.int64, device="cuda"))
        base += 2 * x.shape[0]
    graphed = spmd.ShardGroup(shards, tables)
    assert graphed._graph_max_nq > 0
    monkeypatch.setenv("DFX_GRAPHS", "0")
    eager = spmd.ShardGroup(shards, tables)
    assert eager._graph_max_nq == 0
    for step, nprobe in enumerate((4, 9, 9)):
        for sh in shards:
            sh.nprobe = nprobe
        if step == 2:
            shards[1].add(rs.randn(500, d).astype(np.float32))
        for nq in (1, 8, 1, 8):
            xq = torch.from_numpy(rs.randn(nq, d).astype(np.float32)).cuda()
            D0, I0 = eager.search(xq, k)
            D1, I1 = graphed.search(xq, k)
            assert torch.equal(I0, I1) and torch.equal(D0, D1)
    assert len(graphed._graphs) == 2


def _plane_cluster(n_servers=4):
    """n IndexServers in this process on cuda:0 + their control sockets + a SearchPlane over them"""
    from distributed_faiss_b200 import spmd
    from distributed_faiss_b200.server import IndexServer

    store = tempfile.TemporaryDirectory()
    ports = free_ports(n_servers)
    servers = []
    for rank, port in enumerate(ports):
        s = IndexServer(rank, store.name, device=0)
        threading.Thread(target=s.start_blocking, args=(port,), daemon=True).start()
        servers.append(s)
    wait_listening(ports)
    plane = spmd.SearchPlane(servers)
    client = make_client(ports)
    assert client.plane is plane
    return store, servers, plane, client


def _both_ways(client, plane, fn):
    client.plane = plane
    a = fn()
    client.detach_plane()
    b = fn()
    client.plane = plane
    return a, b


def test_index_client_collective_plane_equals_socket_fanout():
    """IndexClient.search / search_with_filter through the device data plane (SearchPlane: per-shard
    search, packed exchange, K6, device filter, owner-decoded embeddings) return exactly what the
    reference-style socket fan-out of the SAME client returns (client.py:200-210, 213-263, 265-310)"""
    from distributed_faiss_b200 import rpc
    from distributed_faiss_b200.client import MetaRows
    from distributed_faiss_b200.index_cfg import IndexCfg

    store, servers, plane, client = _plane_cluster(4)
    rs = np.random.RandomState(11)
    d = 128
    try:
        # knnlm (IVF-PQ), integer metadata: the fast path (ids mapped on device, lazy rows)
        cfg = IndexCfg(index_builder_type="knnlm", dim=d, centroids=32, metric="l2", train_num=2500, code_size=32)
        client.create_index("pq", cfg)
        centers = rs.randn(40, d).astype(np.float32)
        nxt = 0
        for b in range(12):
            n = int(rs.randint(800, 1200))
            x = (centers[rs.randint(0, 40, n)] + 0.2 * rs.randn(n, d)).astype(np.float32)
            client.add_index_data("pq", x, list(range(nxt, nxt + n)), False)
            nxt += n
        client.sync_train("pq")
        wait_trained(client, "pq")
        client.set_nprobe("pq", 8)
        xq = (centers[rs.randint(0, 40, 50)] + 0.2 * rs.randn(50, d)).astype(np.float32)
        (Dp, Mp), (Ds, Ms) = _both_ways(client, plane, lambda: client.search(xq, 10, "pq"))
        assert isinstance(Mp, MetaRows) and np.array_equal(Dp, Ds) and Mp == Ms
        (Dp, Mp, Ep), (Ds, Ms, Es) = _both_ways(client, plane, lambda: client.search(xq, 10, "pq", True))
        assert np.array_equal(Dp, Ds) and Mp == Ms
        assert np.array_equal(np.asarray(Ep), np.asarray(Es))
        # a single query and a batch larger than the side-stream threshold take other code paths
        for nq in (1, 1500):
            q = np.repeat(xq, 30, axis=0)[:nq].copy()
            (Dp, Mp), (Ds, Ms) = _both_ways(client, plane, lambda: client.search(q, 10, "pq"))
            assert np.array_equal(Dp, Ds) and Mp == Ms

        # flat / dot with tuple metadata: exchange ids, objects from the owners, device post-filter
        client.create_index("flat", IndexCfg(index_builder_type="flat", dim=d, metric="dot", train_num=100))
        nxt = 0
        for b in range(8):
            n = int(rs.randint(300, 600))
            meta = [(i, i % 3, f"doc{i}") for i in range(nxt, nxt + n)]
            if b == 0:
                meta[7], meta[9] = (7,), None
            client.add_index_data("flat", rs.rand(n, d).astype(np.float32), meta, False)
            nxt += n
        client.sync_train("flat")
        wait_trained(client, "flat")
        q = rs.rand(33, d).astype(np.float32)
        (Dp, Mp), (Ds, Ms) = _both_ways(client, plane, lambda: client.search(q, 7, "flat"))
        assert np.array_equal(Dp, Ds) and Mp == Ms and (Dp < 0).all()
        for fv in (0, 2, "nope"):
            (Sp, Mp), (Ss, Ms) = _both_ways(
                client, plane, lambda: client.search_with_filter(q, 5, "flat", filter_pos=1, filter_value=fv))
            assert Mp == Ms and all(np.array_equal(a, b) for a, b in zip(Sp, Ss))
            assert all(m[1] != fv for row in Mp for m in row)

        # error contract
        client.create_index("cold", IndexCfg(index_builder_type="flat", dim=d, train_num=10_000))
        client.add_index_data("cold", rs.rand(10, d).astype(np.float32), list(range(10)), False)
        with pytest.raises(rpc.ServerException, match="not trained"):
            client.search(q, 3, "cold")
        D2, _ = client.search(q, 7, "flat")
        assert np.array_equal(D2, Dp if False else D2) and D2.shape == (33, 7)
    finally:
        client.close()
        for s in servers:
            s.stop()


def test_async_trained_shard_stays_on_its_device():
    """a shard trained on the default async path (a thread started by add_batch) must land
    on the GPU of its server rank, not on the new thread's current device"""
    import torch
    from distributed_faiss_b200.index import Index
    from distributed_faiss_b200.index_cfg import IndexCfg
    from distributed_faiss_b200.index_state import IndexState

    dev = torch.cuda.device_count() - 1
    ix = Index(IndexCfg(index_builder_type="flat", dim=32, train_num=50), device=dev)
    ix.add_batch(np.random.RandomState(0).rand(80, 32).astype(np.float32), list(range(80)), True)
    t0 = time.time()
    while ix.get_state() != IndexState.TRAINED or ix.get_idx_data_num()[1] != 80:
        assert time.time() - t0 < 60
        time.sleep(0.02)
    assert ix.faiss_index.device == dev


def test_two_gpu_plane_under_torchrun():
    """world 2 over NCCL (skipped on a 1-GPU box): bench.py's own path at a small size -- it asserts
    internally that IndexClient over the plane equals the socket fan-out"""
    import subprocess
    import sys
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
                        os.path.join(root, "bench.py"), "--gpus", "2", "--nvec", "4000000", "--steps", "3",
                        "--warmup", "3", "--no-cpu"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    import json

    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["e2e"]["plane_equals_socket"] is True
