"""Test doubles: an oracle-backed engine factory and merge backend.

The product path has NO CPU fallback.  These helpers exist only so that the HOST logic
(state machine, round-robin placement, RPC, persistence, merge plumbing, the world_size>1
collective flow) can be exercised by `-m "not gpu"` tests in a container without a GPU; they
are injected explicitly by the tests and are never selected by product code."""
import numpy as np

from oracle import oracle as O


def oracle_engine_factory(cfg):
    """same builder table as distributed_faiss_b200.index.default_engine_factory"""
    b = cfg.index_builder_type
    if b == "flat":
        return O.make_index("flat", cfg.dim, metric=O.METRIC_IP)
    if b == "ivf_simple":
        ix = O.make_index("ivf_flat", cfg.dim, metric=cfg.get_metric(), nlist=int(cfg.centroids))
        ix.nprobe = cfg.nprobe
        return ix
    if b == "knnlm":
        ix = O.make_index("ivf_pq", cfg.dim, metric=cfg.get_metric(), nlist=int(cfg.centroids),
                          M=int(cfg.extra.get("code_size", 64)), nbits=int(cfg.extra.get("bits_per_vector", 8)))
        cfg.nprobe = ix.nprobe
        return ix
    if b == "ivfsq":
        ix = O.make_index("ivf_sq", cfg.dim, metric=cfg.get_metric(), nlist=int(cfg.centroids))
        ix.nprobe = cfg.nprobe
        return ix
    raise RuntimeError("Either faiss_factory or valid index_builder_type should be specified to initialize index")


def oracle_merge(Dall, Iall, negate):
    Dall = np.asarray(Dall, dtype=np.float32)
    return O.merge(-Dall if negate else Dall, np.asarray(Iall, dtype=np.int64))


class OracleBackend:
    """spmd.ShardGroup backend on CPU tensors (gloo tests)."""

    name = "oracle"

    def search(self, shard, x_t, k):
        import torch

        D, I = shard.search(x_t.numpy(), k)
        return torch.from_numpy(D), torch.from_numpy(I)

    def search_into(self, shard, x_t, k, D_out, I_out, table_t):
        D, I = self.search(shard, x_t, k)
        D_out.copy_(D)
        I_out.copy_(I if table_t is None else self.map_ids(I, table_t))

    def map_ids(self, ids_t, table_t):
        import torch

        out = table_t[ids_t.clamp(min=0)]
        return torch.where(ids_t < 0, torch.full_like(out, -1), out)

    def merge(self, D_t, I_t, negate):
        import torch

        D, I = oracle_merge(D_t.numpy(), I_t.numpy(), negate)
        return torch.from_numpy(D), torch.from_numpy(I)
