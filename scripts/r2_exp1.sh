#!/bin/bash
# experiment 1 (round 2): contiguous-range scan loop, CTA shapes, prefetch distance; ncu of K1
set -u
mkdir -p gpurun_out
echo "== parity subset"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/exp1_suite.log
echo "== scan variants"
MB_VARIANTS="256:4,256:0,256:2,256:8,512:4,512:0,512:8,256:4" timeout 600 python scripts/scan_microbench.py exp1 2>gpurun_out/exp1_mb.err | tail -1 | tee gpurun_out/exp1_microbench.json
echo "== ncu K1"
MB_VARIANTS="256:4" timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_coarse|topg_collect|rerank2" --launch-skip 12 -c 3 \
    -o gpurun_out/exp1_k1 python scripts/scan_microbench.py ncu > gpurun_out/exp1_ncu.log 2>&1
tail -3 gpurun_out/exp1_ncu.log
