"""NCCL data plane: the search fan-out of IndexClient.search over GPUs instead of sockets.

The reference sends the same pickled query to every IndexServer over TCP and merges the
pickled replies on the client's CPU (distributed_faiss/client.py:200-210, 265-310;
rpc.py:120-131).  Inside one 8xB200 box this module replaces that with:

    query  --(ncclBroadcast from the client rank)-->  every rank
    every rank: libdfx search of its resident shards (+ local id -> caller id on device)
    results --(ncclAllGather over NVLink/NVSwitch)--> [S, nq, k] on every rank
    merge kernel K6 (float_maxheap_array_t semantics)  -> (D, I) on device

One process per GPU (`IndexServer(rank=r)` <-> GPU r), `torch.distributed` for the plumbing.
A rank may hold several shards (the 1/2/4-GPU points of the scaling curve hold 8/4/2 shards
each); shards are numbered rank-major so the merge's "earlier shard wins ties" rule is the
same for every GPU count.  The path has exactly one exchange step, as in the reference.

`backend` abstracts the three device operations so that the world_size>1 host logic can be
exercised on CPU with gloo (tests/test_spmd.py injects an oracle-backed backend); the
default backend is CUDA-only and has no fallback.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


class CudaBackend:
    """The product backend: everything runs in libdfx kernels on torch's current stream."""

    name = "cuda"

    def search(self, shard, x_t, k):
        return shard.search_dev(x_t, k)

    def search_into(self, shard, x_t, k, D_out, I_out, table_t):
        """search one shard and leave (D, caller ids) in the given [nq,k] slices, no extra copies"""
        if table_t is None:
            shard.search_dev(x_t, k, D_out, I_out)
        else:
            _, I_local = shard.search_dev(x_t, k, D_out, None)
            self.map_ids(I_local, table_t, I_out)

    def map_ids(self, ids_t, table_t, out_t=None):
        from . import engine

        return engine.map_ids_dev(ids_t, table_t, out_t)

    def merge(self, D_t, I_t, negate):
        from . import engine

        return engine.merge_dev(D_t, I_t, negate=negate)


class ShardGroup:
    """The set of shards of one index spread over the ranks of a process group.

    shards     : this rank's engine objects (engine.GpuIndex), in global shard order
    id_tables  : per local shard, int64 device tensor mapping shard-local id -> caller id
                 (the integer-metadata convention of scripts/load_data.py:120-124), or None
                 to return shard-local ids
    Every rank must own the same number of shards.
    """

    def __init__(self, shards: Sequence, id_tables: Optional[Sequence] = None, group=None,
                 backend=None, device=None):
        self.shards = list(shards)
        self.id_tables = list(id_tables) if id_tables is not None else [None] * len(self.shards)
        assert len(self.id_tables) == len(self.shards)
        self.group = group
        self.backend = backend or CudaBackend()
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.device = device if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
        self._pin_q = None
        self._pin_D = None
        self._pin_I = None
        # side streams for intra-rank shard overlap (CUDA backend only)
        self._streams = []
        self._stream_max_nq = 0
        if self.device.type == "cuda" and isinstance(self.backend, CudaBackend) and len(self.shards) > 1:
            import os

            # latency-bound batches gain ~1.8x from overlapping the shards; large batches already
            # fill the machine and are run back to back (also keeps per-kernel timing clean)
            self._stream_max_nq = int(os.environ.get("DFX_SHARD_STREAMS_MAX_NQ", "1024"))
            n_side = int(os.environ.get("DFX_SHARD_STREAMS", "4"))
            self._streams = [torch.cuda.Stream(device=self.device) for _ in range(max(0, n_side))]
        # EXPERIMENTAL (DFX_GRAPHS=1, off by default, single rank only): latency-bound batches are
        # replayed from a CUDA graph captured per (nq, k, maximize) shape -- a batch-1 search over
        # 8 shards is ~70 short launches whose issue cost dominates.  Graphs are dropped whenever
        # nprobe changes; do not add to the shards while graphs are alive.
        self._graphs = {}
        self._graph_max_nq = 0
        if self.device.type == "cuda" and isinstance(self.backend, CudaBackend) and self.world == 1:
            import os

            if os.environ.get("DFX_GRAPHS") == "1":
                self._graph_max_nq = int(os.environ.get("DFX_GRAPHS_MAX_NQ", "256"))

    @property
    def num_shards(self) -> int:
        return len(self.shards) * self.world

    def set_nprobe(self, nprobe: int):
        for s in self.shards:
            s.nprobe = nprobe
        self._graphs.clear()

    def get_ntotal(self) -> int:
        n = torch.tensor([sum(s.ntotal for s in self.shards)], dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.all_reduce(n, group=self.group)
        return int(n.item())

    # ------------------------------------------------------------------ the hot path
    def search(self, x_t: torch.Tensor, k: int, maximize: bool = False, src: int = 0):
        """Collective.  x_t: float32 [nq, d] on this rank's device; the contents of rank `src`
        are used (broadcast) -- the other ranks only need to pass a tensor of the same shape.
        Returns (D [nq,k] float32 ascending -- negated scores when `maximize`, as the reference
        returns them for metric "dot" --, I [nq,k] int64 caller ids, -1 = no result), on device,
        identical on every rank."""
        if self._graph_max_nq and x_t.is_cuda and x_t.shape[0] <= self._graph_max_nq:
            return self._search_graphed(x_t, k, maximize)
        return self._search_eager(x_t, k, maximize, src)

    def _search_graphed(self, x_t, k, maximize):
        key = (tuple(x_t.shape), int(k), bool(maximize))
        ent = self._graphs.get(key)
        if ent is None:
            static_x = x_t.clone()
            cur = torch.cuda.current_stream(x_t.device)
            side = torch.cuda.Stream(device=x_t.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):  # two eager passes: workspaces reach their final size
                for _ in range(2):
                    self._search_eager(static_x, k, maximize, 0)
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(graph):
                    out = self._search_eager(static_x, k, maximize, 0)
            except RuntimeError as e:  # something in the path is not capturable: stay eager
                import logging

                logging.getLogger("distributed_faiss_b200").warning("CUDA-graph capture failed, disabled: %s", e)
                self._graph_max_nq = 0
                self._graphs.clear()
                torch.cuda.synchronize(x_t.device)
                return self._search_eager(x_t, k, maximize, 0)
            ent = (graph, static_x, out)
            self._graphs[key] = ent
        graph, static_x, out = ent
        static_x.copy_(x_t)
        graph.replay()
        return out[0].clone(), out[1].clone()

    def _search_eager(self, x_t: torch.Tensor, k: int, maximize: bool = False, src: int = 0):
        if self.world > 1:
            dist.broadcast(x_t, src=src, group=self.group)
        nq = x_t.shape[0]
        S_loc = len(self.shards)
        D_loc = torch.empty((S_loc, nq, k), dtype=torch.float32, device=x_t.device)
        I_loc = torch.empty((S_loc, nq, k), dtype=torch.int64, device=x_t.device)
        if self._streams and S_loc > 1 and nq <= self._stream_max_nq:
            # shards of one rank are independent: alternate them over a few side streams so the
            # short kernels of one shard (table build, re-rank, selection) fill the tail of
            # another shard's list scan
            main = torch.cuda.current_stream(x_t.device)
            ready = torch.cuda.Event()
            ready.record(main)
            for j, shard in enumerate(self.shards):
                st = self._streams[j % len(self._streams)]
                if j < len(self._streams):
                    st.wait_event(ready)
                with torch.cuda.stream(st):
                    self.backend.search_into(shard, x_t, k, D_loc[j], I_loc[j], self.id_tables[j])
            for st in self._streams[:min(len(self._streams), S_loc)]:
                done = torch.cuda.Event()
                done.record(st)
                main.wait_event(done)
            D_loc.record_stream(main)
        else:
            for j, shard in enumerate(self.shards):
                self.backend.search_into(shard, x_t, k, D_loc[j], I_loc[j], self.id_tables[j])
        if self.world > 1:
            D_all = torch.empty((self.world * S_loc, nq, k), dtype=torch.float32, device=x_t.device)
            I_all = torch.empty((self.world * S_loc, nq, k), dtype=torch.int64, device=x_t.device)
            dist.all_gather_into_tensor(D_all, D_loc, group=self.group)
            dist.all_gather_into_tensor(I_all, I_loc, group=self.group)
        else:
            D_all, I_all = D_loc, I_loc
        return self.backend.merge(D_all, I_all, maximize)

    def search_host(self, x: np.ndarray, k: int, maximize: bool = False, src: int = 0):
        """End-to-end form with HOST buffers: pinned staging, H2D of the query batch, the
        collective search, D2H of (D, I).  This is what IndexClient.search costs a caller."""
        nq, d = x.shape
        if self.device.type != "cuda":
            D, I = self.search(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)), k, maximize, src)
            return D.numpy(), I.numpy()
        if self._pin_q is None or self._pin_q.shape != (nq, d):
            self._pin_q = torch.empty((nq, d), dtype=torch.float32, pin_memory=True)
        if self._pin_D is None or self._pin_D.shape != (nq, k):
            self._pin_D = torch.empty((nq, k), dtype=torch.float32, pin_memory=True)
            self._pin_I = torch.empty((nq, k), dtype=torch.int64, pin_memory=True)
        self._pin_q.numpy()[...] = x
        x_t = self._pin_q.to(self.device, non_blocking=True)
        D, I = self.search(x_t, k, maximize, src)
        self._pin_D.copy_(D, non_blocking=True)
        self._pin_I.copy_(I, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._pin_D.numpy(), self._pin_I.numpy()


def init_process_group_from_env(backend: Optional[str] = None):
    """One process per GPU launched by torchrun: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*."""
    import os

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank % torch.cuda.device_count())
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kw = {}
        if torch.cuda.is_available():
            kw["device_id"] = torch.device("cuda", torch.cuda.current_device())
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"),
                                rank=rank, world_size=world, **kw)
    return rank, local_rank, world
