"""Host-side behaviour of the API mirror (IndexCfg / IndexState / IndexServer / IndexClient /
rpc), written after the reference's own tests (tests/test_integration.py, test_client.py,
test_index_state.py, test_rpc.py): real sockets, real threads, no mocks of the transport.
The numeric engine is the ORACLE (injected test double, tests/oracle_engine.py) because this
container has no GPU; the same scenarios run against the CUDA engine in tests/test_gpu_api.py."""
import os
import random
import socket
import string
import tempfile
import threading
import time

import numpy as np
import pytest

from distributed_faiss_b200.client import IndexClient, ResultHeap
from distributed_faiss_b200.index import Index
from distributed_faiss_b200.index_cfg import IndexCfg
from distributed_faiss_b200.index_state import IndexState
from distributed_faiss_b200.rpc import ServerException
from distributed_faiss_b200.server import IndexServer
from tests.oracle_engine import oracle_engine_factory, oracle_merge

HERE = os.path.dirname(os.path.abspath(__file__))


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


def rand_meta(n, nchars=5):
    return ["".join(random.choices(string.ascii_uppercase + string.digits, k=nchars)) for _ in range(n)]


@pytest.fixture(scope="module")
def cluster():
    """4 shard servers + 1 single server on localhost, oracle engines, device merge replaced"""
    ResultHeap.merge_backend = staticmethod(oracle_merge)
    dirs = [tempfile.TemporaryDirectory(), tempfile.TemporaryDirectory()]
    multi_ports = free_ports(5)
    multi_ports, single_port = multi_ports[:4], multi_ports[4]
    servers = []
    for rank, port in enumerate(multi_ports):
        s = IndexServer(rank, dirs[0].name, engine_factory=oracle_engine_factory)
        threading.Thread(target=s.start_blocking, args=(port,), daemon=True).start()
        servers.append(s)
    single = IndexServer(0, dirs[1].name, engine_factory=oracle_engine_factory)
    threading.Thread(target=single.start_blocking, args=(single_port,), daemon=True).start()
    wait_listening(multi_ports + [single_port])
    yield {"multi_ports": multi_ports, "single_port": single_port, "servers": servers, "single": single,
           "dirs": dirs}
    for s in servers + [single]:
        s.stop()


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
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as fh:
        fh.write(f"{len(ports)}\n")
        for p in ports:
            fh.write(f"localhost,{p}\n")
        path = fh.name
    try:
        return IndexClient(path)
    finally:
        os.unlink(path)


def wait_trained(client, index_id, timeout=60):
    t0 = time.time()
    while client.get_state(index_id) != IndexState.TRAINED:
        assert time.time() - t0 < timeout, "index never reached TRAINED"
        time.sleep(0.05)


# ------------------------------------------------------------------ unit level
def test_index_cfg_json_roundtrip():
    cfg = IndexCfg.from_json(os.path.join(HERE, "golden", "index_config.json"))
    assert cfg.dim == 1024 and isinstance(cfg.dim, int)          # string -> int (index_cfg.py:31)
    assert cfg.extra == {"factory_type": "IVF{centroids},SQ8"}   # unknown keys land in .extra
    assert cfg.metric == "dot" and cfg.nprobe == 1 and cfg.buffer_bsz == 50000
    assert cfg.get_metric() == 0 and IndexCfg(metric="l2").get_metric() == 1
    with pytest.raises(RuntimeError, match="Only dot and l2"):
        IndexCfg(metric="cosine").get_metric()
    knn = IndexCfg(index_builder_type="knnlm", dim=128, centroids=64, code_size=32, bits_per_vector=8)
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as fh:
        fh.write(knn.to_json_string())
    back = IndexCfg.from_json(fh.name)
    os.unlink(fh.name)
    assert back.extra["code_size"] == 32 and back.index_builder_type == "knnlm"


def test_index_state_aggregation():
    S = IndexState                                               # reference tests/test_index_state.py:14-22
    assert S.get_aggregated_states([S.TRAINED, S.TRAINED]) == S.TRAINED
    assert S.get_aggregated_states([S.TRAINED, S.NOT_TRAINED]) == S.NOT_TRAINED
    assert S.get_aggregated_states([S.TRAINED, S.NOT_TRAINED, S.TRAINING]) == S.TRAINING
    assert S.get_aggregated_states([S.TRAINED, S.ADD]) == S.ADD
    assert [s.value for s in S] == [1, 2, 3, 4]


def test_read_server_list():
    res = IndexClient.read_server_list(os.path.join(HERE, "golden", "server_list.txt"))
    assert res == [("hostA", 8080), ("hostA", 8081), ("hostB", 8080), ("hostB", 8081)]
    path = os.path.join(HERE, "golden", "server_list_short.txt")
    with pytest.raises(AssertionError) as e:
        IndexClient.read_server_list(path, total_max_timeout=0.0)
    assert str(e.value) == f"4 != 3 in server list {path}. Timed out after waiting 0.0 seconds"


def test_result_aggregation_reference_vectors():
    """reference tests/test_integration.py:181-203, run through IndexClient._aggregate_results"""
    ResultHeap.merge_backend = staticmethod(oracle_merge)
    mock = [(np.array([[12.1, 13.2, 13.3, 14.3]], dtype=np.float32), [[1465, 1460, 443197, 1340]], None),
            (np.array([[8.1, 12.6, 13.1, 17.4]], dtype=np.float32), [[0, 14, 3, 1]], None)]
    D, i_min = IndexClient._aggregate_results(mock, 4, 1, False, False)
    Dmax, i_max = IndexClient._aggregate_results(mock, 4, 1, True, False)
    assert i_max != i_min and i_min[0][0] == 0 and D[0][0] < D[0][1] and i_max[0][0] == 1 and 0 in i_min[0]
    assert i_min == [[0, 1465, 14, 3]] and i_max == [[1, 1340, 443197, 1460]]
    assert np.array_equal(Dmax, np.array([[-17.4, -14.3, -13.3, -13.2]], dtype=np.float32))
    # with embeddings, and with a shard that has fewer than k hits
    mock2 = [(np.array([[1.0, np.finfo(np.float32).max]], dtype=np.float32), [["a", None]], [[np.ones(2), None]]),
             (np.array([[0.5, 2.0]], dtype=np.float32), [["b", "c"]], [[np.zeros(2), np.full(2, 2.0)]])]
    D2, m2, e2 = IndexClient._aggregate_results(mock2, 2, 1, False, True)
    assert m2 == [["b", "a"]] and np.array_equal(e2[0][0], np.zeros(2))
    rh = ResultHeap(1, 2)
    rh.add_result(np.array([[3.0, 4.0]], np.float32), np.array([[10, 11]]))
    rh.add_result(np.array([[1.0, 5.0]], np.float32), np.array([[20, 21]]))
    rh.finalize()
    assert rh.D.tolist() == [[1.0, 3.0]] and rh.I.tolist() == [[20, 10]]


def test_index_state_machine_and_buffering():
    cfg = IndexCfg(index_builder_type="flat", dim=16, train_num=10, buffer_bsz=7)
    ix = Index(cfg, engine_factory=oracle_engine_factory)
    rs = np.random.RandomState(0)
    ix.add_batch(rs.rand(9, 16).astype(np.float32), list(range(9)), train_async_if_triggered=False)
    assert ix.get_state() == IndexState.NOT_TRAINED
    with pytest.raises(RuntimeError, match="Server index is not trained"):
        ix.search(rs.rand(2, 16).astype(np.float32), 3)
    with pytest.raises(RuntimeError, match="metadata length"):
        ix.add_batch(rs.rand(3, 16).astype(np.float32), [1, 2])
    ix.add_batch(rs.rand(1, 16).astype(np.float32), [9], train_async_if_triggered=False)  # hits train_num
    t0 = time.time()
    while ix.get_state() != IndexState.TRAINED:
        assert time.time() - t0 < 20
        time.sleep(0.01)
    assert ix.get_idx_data_num() == (0, 10)
    ix.add_batch(rs.rand(20, 16).astype(np.float32), list(range(10, 30)))  # TRAINED -> ADD -> TRAINED
    t0 = time.time()
    while ix.get_idx_data_num() != (0, 30) or ix.get_state() != IndexState.TRAINED:
        assert time.time() - t0 < 20
        time.sleep(0.01)
    D, meta, embs = ix.search(rs.rand(3, 16).astype(np.float32), 40)
    assert D.shape == (3, 40) and embs is None
    assert all(m is None for m in meta[0][30:]) and sorted(meta[0][:30]) == list(range(30))  # -1 -> None
    D, meta, embs = ix.search(rs.rand(3, 16).astype(np.float32), 2, return_embeddings=True)
    assert embs.shape == (3, 2, 16)
    assert Index.infer_n_centroids(10_000) == 200 and Index.infer_n_centroids(5e6) == 65536


# ------------------------------------------------------------------ client / server over sockets
def test_train_num_honored_and_save_load(cluster):
    index_id = "t_train_num"
    cfg = IndexCfg(index_builder_type="flat", dim=32, train_num=10)
    client = make_client([cluster["single_port"]])
    client.create_index(index_id, cfg)
    rs = np.random.RandomState(1)

    def add(n):
        client.add_index_data(index_id, rs.rand(n, 32).astype(np.float32), rand_meta(n), False)
        return client.get_state(index_id)

    assert add(9) == IndexState.NOT_TRAINED
    assert add(1) != IndexState.NOT_TRAINED
    wait_trained(client, index_id)
    res = client.search(rs.rand(4, 32).astype(np.float32), 4, index_id)
    assert res[0].shape == (4, 4)
    client.save_index(index_id)
    assert os.path.isfile(os.path.join(cluster["dirs"][1].name, index_id, "0", "cfg.json"))
    client.close()
    client2 = make_client([cluster["single_port"]])
    assert client2.load_index(index_id, cfg)
    assert client2.get_state(index_id) == IndexState.TRAINED and client2.get_ntotal(index_id) == 10
    assert client2.search(rs.rand(4, 32).astype(np.float32), 4, index_id)[0].shape == (4, 4)
    # cfg read back from disk when none is given; a cfg given at load time overrides nprobe
    assert client2.load_index(index_id) and client2.cfg.dim == 32
    client2.drop_index(index_id)
    assert client2.get_ntotal(index_id) == 0
    client2.close()


def test_search_quality_same_for_multiple_clients(cluster):
    """sharded search + merge is bit-identical to unsharded search (reference :205-265)"""
    index_id = "t_same"
    d = 64
    cfg = IndexCfg(index_builder_type="flat", dim=d)
    single = make_client([cluster["single_port"]])
    single.create_index(index_id, cfg)
    clients = [make_client(cluster["multi_ports"]) for _ in range(4)]
    rs = np.random.RandomState(2)
    for c in clients:
        c.create_index(index_id, cfg)
        for _ in range(random.randint(1, 4)):
            n = random.randint(1, 800)
            emb, meta = rs.rand(n, d).astype(np.float32), rand_meta(n)
            c.add_index_data(index_id, emb, meta, False)
            single.add_index_data(index_id, emb, meta, False)
            assert c.get_state(index_id) == IndexState.NOT_TRAINED
    # the reference test stops here and is flaky by construction: with 1-4 batches per client a
    # shard can stay empty, and training an empty buffer raises (there as here).  Four more
    # round-robin batches from one client reach every shard.
    for _ in range(4):
        emb, meta = rs.rand(50, d).astype(np.float32), rand_meta(50)
        clients[0].add_index_data(index_id, emb, meta, False)
        single.add_index_data(index_id, emb, meta, False)
    clients[0].sync_train(index_id)
    single.sync_train(index_id)
    wait_trained(clients[0], index_id)
    wait_trained(single, index_id)
    assert clients[0].get_ntotal(index_id) == single.get_ntotal(index_id)
    q = rs.rand(16, d).astype(np.float32)
    s_aggr, m_aggr = clients[0].search(q, 5, index_id)
    s_single, m_single = single.search(q, 5, index_id)
    assert (s_aggr == s_single).all() and m_aggr == m_single
    assert (s_aggr <= 0).all() and (np.diff(s_aggr, axis=1) >= 0).all()   # "dot": negated, ascending (B5)
    for c in clients + [single]:
        c.close()


def test_round_robin_balance_and_misc(cluster):
    index_id = "t_rr"
    d, per = 32, 50
    cfg = IndexCfg(index_builder_type="flat", dim=d)
    clients = [make_client(cluster["multi_ports"]) for _ in range(4)]
    rs = np.random.RandomState(3)
    for c in clients:
        c.create_index(index_id, cfg)
        for _ in range(8):                                            # 8 batches over 4 servers: 2 each
            c.add_index_data(index_id, rs.rand(per, d).astype(np.float32), rand_meta(per), False)
    clients[0].sync_train(index_id)
    wait_trained(clients[0], index_id)
    for srv in cluster["servers"]:
        assert srv.get_ntotal(index_id) == 2 * 4 * per
    assert clients[0].get_ntotal(index_id) == 4 * 8 * per
    assert clients[0].get_ntotal("wrong_id") == 0
    D, meta = clients[1].search(rs.rand(16, d).astype(np.float32), 5, index_id)
    assert D.shape == (16, 5) and len(meta) == 16 and len(meta[0]) == 5
    with pytest.raises(ServerException, match="Server has no index"):
        clients[0].sub_indexes[0].search("nope", rs.rand(1, d).astype(np.float32), 1, False)
    with pytest.raises(ServerException):
        clients[0].set_omp_num_threads(4)                              # no server implements it
    # the engine's per-query result limit is reported by the client, not as 4 ServerExceptions
    with pytest.raises(ValueError, match="topk"):
        clients[1].search(rs.rand(2, d).astype(np.float32), 5000, index_id)
    with pytest.raises(ValueError, match="topk"):
        clients[1].search_with_filter(rs.rand(2, d).astype(np.float32), 2000, index_id, filter_pos=0, filter_value=1)
    assert clients[0].get_num_servers() == 4
    clients[0].save_index(index_id)
    clients[0].drop_index(index_id)
    assert clients[0].get_ntotal(index_id) == 0
    for c in clients:
        c.close()


def test_ivf_builders_through_the_api(cluster):
    """ivf_simple / knnlm / ivfsq: train, centroids shape (reference :387-416), set_nprobe, search"""
    rs = np.random.RandomState(4)
    d = 32
    client = make_client(cluster["multi_ports"])
    for builder, extra in (("ivf_simple", {}), ("knnlm", {"code_size": 8}), ("ivfsq", {})):
        index_id = "t_" + builder
        cfg = IndexCfg(index_builder_type=builder, dim=d, centroids=4, metric="l2", nprobe=2, train_num=300, **extra)
        client.create_index(index_id, cfg)
        for _ in range(4):                                            # one batch of 400 per shard
            x = (rs.randn(6, d)[rs.randint(0, 6, 400)] + 0.1 * rs.randn(400, d)).astype(np.float32)
            client.add_index_data(index_id, x, [(i,) for i in range(400)], False)
        wait_trained(client, index_id)
        cents = client.get_centroids(index_id)
        assert len(cents) == 4 and cents[0].shape == (4, d)
        client.set_nprobe(index_id, 4)
        D, meta = client.search(x[:7], 3, index_id)
        assert D.shape == (7, 3) and (np.diff(D, axis=1) >= 0).all()
        assert client.get_ntotal(index_id) == 1600
        scores, fmeta = client.search_with_filter(x[:3], 2, index_id, filter_pos=0, filter_value=meta[0][0][0])
        assert len(fmeta) == 3 and all(m[0] != meta[0][0][0] for row in fmeta for m in row)
    client.close()


def test_many_clients_one_server(cluster):
    """reference tests/test_rpc.py:29-64: concurrent clients against one server"""
    port = cluster["single_port"]
    errors = []

    def worker(i):
        try:
            c = make_client([port])
            index_id = f"t_rpc_{i}"
            c.create_index(index_id, IndexCfg(index_builder_type="flat", dim=16))
            rs = np.random.RandomState(i)
            for _ in range(5):
                c.add_index_data(index_id, rs.rand(20, 16).astype(np.float32), rand_meta(20))
            c.async_train(index_id)
            wait_trained(c, index_id)
            for _ in range(5):
                D, meta, embs = c.sub_indexes[0].search(index_id, rs.rand(5, 16).astype(np.float32), 5, True)
                assert D.shape == (5, 5) and embs.shape == (5, 5, 16)
            c.close()
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(6)]
    [t.start() for t in threads]
    [t.join(60) for t in threads]
    assert not errors, errors


def test_config_to_file_and_nprobe_override(cluster):
    """reference tests/test_integration.py:332-385: cfg.json is written next to the shard, a load
    without cfg reads it back, a cfg given at load time overrides nprobe on the live index."""
    index_id = "t_cfgfile"
    rs = np.random.RandomState(5)
    d = 16
    cfg = IndexCfg(index_builder_type="ivf_simple", dim=d, centroids=4, metric="l2", nprobe=2, train_num=200)
    client = make_client([cluster["single_port"]])
    client.create_index(index_id, cfg)
    x = rs.rand(300, d).astype(np.float32)
    client.add_index_data(index_id, x, list(range(300)), False)
    wait_trained(client, index_id)
    client.save_index(index_id)
    cfg_path = os.path.join(cluster["dirs"][1].name, index_id, "0", "cfg.json")
    on_disk = IndexCfg.from_json(cfg_path)
    assert on_disk.nprobe == 2 and on_disk.centroids == 4 and on_disk.index_builder_type == "ivf_simple"
    D_before, m_before = client.search(x[:5], 3, index_id)
    client.close()
    c2 = make_client([cluster["single_port"]])
    assert c2.load_index(index_id)                                   # cfg comes from the file
    assert c2.cfg.nprobe == 2 and c2.cfg.metric == "l2"
    D_after, m_after = c2.search(x[:5], 3, index_id)
    assert np.array_equal(D_before, D_after) and m_before == m_after  # reloaded shard answers identically
    cfg2 = IndexCfg(index_builder_type="ivf_simple", dim=d, centroids=4, metric="l2", nprobe=3)
    assert c2.load_index(index_id, cfg2, force_reload=False)          # already loaded: cfg is applied
    assert cluster["single"].indexes[index_id].faiss_index.nprobe == 3
    c2.close()


def test_flat_builder_ignores_l2_metric(cluster):
    """SURVEY quirk B1 (reference index.py:94 + client.py:206): the "flat" builder is always an
    inner-product index; with metric="l2" each shard still returns its LARGEST inner products and
    the client merges them keeping the SMALLEST, un-negated.  Reproduced, not fixed."""
    from oracle import oracle as O

    index_id = "t_l2_flat"
    rs = np.random.RandomState(6)
    d = 16
    cfg = IndexCfg(index_builder_type="flat", dim=d, metric="l2")
    client = make_client(cluster["multi_ports"])
    client.create_index(index_id, cfg)
    shards = []
    for s in range(4):
        x = rs.rand(50, d).astype(np.float32)
        shards.append(x)
        client.add_index_data(index_id, x, [f"s{s}_{i}" for i in range(50)], False)
    client.sync_train(index_id)
    wait_trained(client, index_id)
    q = rs.rand(3, d).astype(np.float32)
    D, meta = client.search(q, 2, index_id)
    # expected by hand from the rule above
    per_shard = [O.flat_search(O.METRIC_IP, x, q, 2)[0] for x in shards]  # top-2 largest IP per shard
    pool = np.concatenate(per_shard, axis=1)                               # [3, 8]
    expect = np.sort(pool, axis=1)[:, :2]
    assert np.array_equal(D, expect) and (D > 0).all()
    client.close()
