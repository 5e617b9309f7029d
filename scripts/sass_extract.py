#!/usr/bin/env python
"""SASS evidence of the hot kernels from the shipped object files (no GPU needed):
    python scripts/sass_extract.py profiles/r02_sass
writes <prefix>_<kernel>.txt (full listing) and <prefix>_summary.json (mnemonic histograms:
UTCHMMA / LDTM / UTMALDG / UBLKCP / SYNCS prove tcgen05 + TMEM + TMA; LDS / PRMT / FADD2 counts
of the scan's inner loop)."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "distributed_faiss_b200", "build")
KERNELS = {"dfx_tc.o": ["tc_coarse_kernel", "rerank2_kernel", "topg_collect_kernel", "tc_exact_rows_kernel"],
           "dfx_scan_il2.o": ["scan_pq_il2_kernel"], "dfx_search.o": ["scan_rows_kernel", "dfx_select_rows_kernel"]}
# full listings (instruction text only) of the instantiations the headline configuration runs:
# tc_coarse_kernel<KT=2 (d=128), L2, PRECISE / FAST> and scan_pq_il2_kernel<REG, 256 threads, 3 CTAs/SM>
LISTED = ["tc_coarse_kernelILi2ELi1ELi2E", "tc_coarse_kernelILi2ELi1ELi1E", "scan_pq_il2_kernelILb1ELi256ELi3E"]


def main():
    prefix = sys.argv[1]
    summary = {}
    for obj, names in KERNELS.items():
        txt = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
        funcs = re.split(r"\n\s*Function : ", txt)
        for f in funcs[1:]:
            mangled = f.split("\n", 1)[0].strip()
            for n in names:
                if n in mangled:
                    ops = collections.Counter()
                    for line in f.splitlines():
                        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
                        if m:
                            ops[m.group(1).split(".")[0]] += 1
                    summary[mangled] = dict(ops.most_common())
                    short = re.sub(r"[^A-Za-z0-9]+", "_", mangled)[:80]
                    if any(t in mangled for t in LISTED):
                        body = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln) for ln in f.splitlines()]
                        body = [ln for ln in body if ln.strip()]
                        open(f"{prefix}_{short}.txt", "w").write("Function : " + "\n".join(body) + "\n")
    json.dump(summary, open(prefix + "_summary.json", "w"), indent=1)
    for k, v in summary.items():
        keys = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "LDS", "PRMT", "FADD2", "LDG", "FFMA"]
        print(k[:70], {kk: v[kk] for kk in keys if kk in v})


if __name__ == "__main__":
    main()
