"""distributed_faiss_b200 -- a B200-native engine behind distributed-faiss's search API.

Drop-in for the SEARCH PATH of facebookresearch/distributed-faiss
(`IndexClient.search` -> per-shard coarse quantizer -> PQ/SQ tables -> inverted-list scan ->
cross-shard merge): same `IndexCfg` / `IndexServer` / `IndexClient` / `IndexState` names and
argument meaning, hand-written sm_100a CUDA underneath (libdfx.so, include/dfx.h), NCCL over
NVLink for the multi-GPU data plane.  See DESIGN.md and INTEGRATION.md.
"""
from .index_cfg import IndexCfg  # noqa: F401
from .index_state import IndexState  # noqa: F401

__all__ = ["IndexCfg", "IndexState", "IndexServer", "IndexClient"]


def __getattr__(name):
    if name == "IndexServer":
        from .server import IndexServer

        return IndexServer
    if name == "IndexClient":
        from .client import IndexClient

        return IndexClient
    raise AttributeError(name)
