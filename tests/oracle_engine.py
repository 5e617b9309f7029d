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
    """spmd.ShardGroup / SearchPlane backend on CPU tensors (gloo tests): the device operations of
    spmd.CudaBackend restated with the oracle and numpy."""

    name = "oracle"

    def search_local(self, shard, x_t, k, D_out, I_out):
        import torch

        D, I = shard.search(x_t.numpy(), k)
        D_out.copy_(torch.from_numpy(D))
        I_out.copy_(torch.from_numpy(I))

    def map_ids(self, ids_t, table_t, out_t=None):
        import torch

        out = table_t[ids_t.clamp(min=0)]
        out = torch.where(ids_t < 0, torch.full_like(out, -1), out)
        if out_t is not None:
            out_t.copy_(out)
            return out_t
        return out

    def encode_ids(self, ids_t, tag, out_t, col_t=None, drop_code=-1):
        import torch

        e = (int(tag) << 40) | ids_t.clamp(min=0)
        if col_t is not None:
            c = col_t[ids_t.clamp(min=0)]
            e = torch.where((c == drop_code) | (c == -2), e | (1 << 62), e)
        out_t.copy_(torch.where(ids_t < 0, torch.full_like(e, -1), e))
        return out_t

    def merge_packed(self, packed_t, R, S_loc, nq, k, stride, off_I, negate):
        import torch

        raw = packed_t.numpy()
        n = S_loc * nq * k
        Ds, Is = [], []
        for r in range(R):
            blk = raw[r * stride:(r + 1) * stride]
            Ds.append(blk[:4 * n].view(np.float32).reshape(S_loc, nq, k))
            Is.append(blk[off_I:off_I + 8 * n].view(np.int64).reshape(S_loc, nq, k))
        D, I = oracle_merge(np.concatenate(Ds), np.concatenate(Is), negate)
        return torch.from_numpy(D), torch.from_numpy(I)

    def filter_compact(self, D_t, I_t, k_out):
        import torch

        D, I = D_t.numpy(), I_t.numpy()
        nq = D.shape[0]
        oD = np.full((nq, k_out), np.finfo(np.float32).max, dtype=np.float32)
        oI = np.full((nq, k_out), -1, dtype=np.int64)
        cnt = np.zeros(nq, dtype=np.int32)
        for q in range(nq):
            keep = [j for j in range(D.shape[1]) if I[q, j] >= 0 and not (I[q, j] >> 62) & 1][:k_out]
            oD[q, :len(keep)] = D[q, keep]
            oI[q, :len(keep)] = I[q, keep]
            cnt[q] = len(keep)
        return torch.from_numpy(oD), torch.from_numpy(oI), torch.from_numpy(cnt)

    def reconstruct_owned(self, shard, I_t, R_t, tag):
        import torch

        I = I_t.numpy().reshape(-1)
        own = (I >= 0) & (((I >> 40) & 0xFFFFF) == tag)
        if own.any():
            rows = shard.reconstruct_rows(I[own] & ((1 << 40) - 1))
            R_t.view(-1, R_t.shape[-1])[torch.from_numpy(np.nonzero(own)[0])] = torch.from_numpy(rows)

    def merge(self, D_t, I_t, negate):
        import torch

        D, I = oracle_merge(D_t.numpy(), I_t.numpy(), negate)
        return torch.from_numpy(D), torch.from_numpy(I)
