#!/bin/bash
# experiment 6: N = 256 tiles in the PRECISE screening kernel
set -u
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core or flat" 2>&1 | tail -4 | tee gpurun_out/exp6_suite.log
MB_VARIANTS="256:4" MB_REPS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 40 --csv \
    --log-file gpurun_out/exp6_launches.csv python scripts/scan_microbench.py exp6 > gpurun_out/exp6_ncu.log 2>&1
tail -1 gpurun_out/exp6_ncu.log | cut -c1-200
timeout 300 python scripts/bench_other_configs.py > gpurun_out/exp6_other_configs.log 2>&1; tail -30 gpurun_out/exp6_other_configs.log
