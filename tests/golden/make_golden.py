"""Regenerates tests/golden/oracle_small.npz: seeded small inputs and the oracle's outputs for
them.  faiss is not installable here (see oracle/dfx_oracle.c), so these vectors pin the
ORACLE (drift detection) and give the GPU tests a committed fixture to compare against; the
reference's own golden vectors for this path are in merge_golden.json.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def main():
    rs = np.random.RandomState(20260922)
    d, n, nq, nlist, M, k = 32, 600, 6, 8, 8, 5
    centers = rs.randn(12, d).astype(np.float32)
    xb = (centers[rs.randint(0, 12, n)] + 0.3 * rs.randn(n, d)).astype(np.float32)
    xq = (centers[rs.randint(0, 12, nq)] + 0.3 * rs.randn(nq, d)).astype(np.float32)
    out = {"xb": xb, "xq": xq, "k": np.int64(k)}
    for metric, name in ((O.METRIC_IP, "ip"), (O.METRIC_L2, "l2")):
        D, I = O.flat_search(metric, xb, xq, k)
        out[f"flat_{name}_D"], out[f"flat_{name}_I"] = D, I
    for kind, kw in (("ivf_flat", {}), ("ivf_pq", {"M": M}), ("ivf_sq", {})):
        ix = O.make_index(kind, d, metric=O.METRIC_L2, nlist=nlist, **kw)
        ix.train_niter = 10
        ix.train(xb)
        ix.add(xb)
        st = ix.get_state()
        for key, val in st.items():
            if isinstance(val, np.ndarray):
                out[f"{kind}_state_{key}"] = val
        for nprobe in (2, nlist):
            ix.nprobe = nprobe
            D, I = ix.search(xq, k)
            out[f"{kind}_np{nprobe}_D"], out[f"{kind}_np{nprobe}_I"] = D, I
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small.npz"), **out)
    print("wrote oracle_small.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
