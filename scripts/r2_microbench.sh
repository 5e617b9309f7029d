#!/bin/bash
# One 125 M-vector shard of the bench workload: scan-kernel variants in round robin
# (MB_VARIANTS="threads:prefetch,..."), the gap statistics the screening tolerance is up against,
# and the coarse-quantizer precision modes.   gpurun --timeout 900 -- 'bash scripts/r2_microbench.sh'
set -u
mkdir -p gpurun_out
MB_VARIANTS="${MB_VARIANTS:-256:2,256:4,256:8,512:2}" MB_REPS="${MB_REPS:-3}" timeout 600 python scripts/scan_microbench.py mb \
    2>gpurun_out/mb.err | tail -1 | tee gpurun_out/microbench.json
