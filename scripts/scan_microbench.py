#!/usr/bin/env python
"""One shard of the bench workload (125 M vectors of the 1 B, nlist 65 536), batch 4096, nprobe 8:
per-launch time of the fused table-build + scan kernel and of the whole shard search, for the
kernel knobs currently in the environment (DFX_IL2_PREFETCH ...).  ~40 s per run.
    python scripts/scan_microbench.py [label]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from distributed_faiss_b200 import engine  # noqa: E402


class A:
    nlist = 0
    kmeans_niter = 10
    train_pts = 64


def main():
    label = sys.argv[1] if len(sys.argv) > 1 else "default"
    nvec = int(os.environ.get("MB_NVEC", 1_000_000_000))
    G = nvec // 10
    synth = engine.Synth(1234, bench.D, 16, 1, 1.0, 0.02, ngroups=G, eps=0.01, delta=0.1)
    idx, tab, info = bench.build_shard(engine, synth, nvec, 0, A, torch, None)
    g = torch.Generator(device="cpu").manual_seed(1236)
    q_rows = torch.randint(0, nvec, (12288,), generator=g, dtype=torch.int64).cuda()
    xq = synth.rows(0, 12288, rows_t=q_rows, noise_stream=7)
    idx.nprobe = int(os.environ.get("MB_NPROBE", 8))
    out = {"label": label, "info": info}
    # MB_VARIANTS="threads:prefetch,..." e.g. "256:4,256:0,512:4" (dfx_set_param il2_threads / il2_prefetch)
    variants = [tuple(int(x) for x in v.split(":")) for v in os.environ.get("MB_VARIANTS", "0:-1").split(",")]
    ref = None
    for (thr, pf) in variants:
      idx.set_param("il2_threads", thr)
      idx.set_param("il2_prefetch", pf)
      for B in (4096,):
          batches = [xq[i * B:(i + 1) * B].contiguous() for i in range(12288 // B)]
          for b in batches:
              idx.search_dev(b, 10)
          torch.cuda.synchronize()
          ndis = idx.last_stats()["ndis"]
          idx.profile(True)
          idx.profile_read(True)
          e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          e0.record()
          n = 30
          for i in range(n):
              idx.search_dev(batches[i % len(batches)], 10)
          e1.record()
          torch.cuda.synchronize()
          ms, nl = idx.profile_read(True)
          idx.profile(False)
          Dv, Iv = idx.search_dev(batches[0], 10)
          if ref is None:
              ref = (Dv.clone(), Iv.clone())
          same = bool(torch.equal(Dv, ref[0]) and torch.equal(Iv, ref[1]))
          out[f"t{thr}_pf{pf}_B{B}"] = {"same_as_first_variant": same, "scan_ms_per_launch": ms / nl, "search_ms": e0.elapsed_time(e1) / n,
                          "algorithmic_GBps": ndis * 32 / (ms / nl) / 1e6, "frac_of_6568": ndis * 32 / (ms / nl) / 1e6 / 6568.4}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
