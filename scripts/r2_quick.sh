#!/bin/bash
# quick GPU check: parity suite, bench (with the batch sweep), launch list of one microbench search
set -u
mkdir -p gpurun_out
TAG=${TAG:-r02q}
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/${TAG}_suite.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"], d["clocks"])
print(d["config"].get("qps_by_batch"), d["config"].get("e2e_qps_by_batch"))
PY
grep -v "^\[bench" gpurun_out/${TAG}_bench.err | head -5
MB_VARIANTS="256:4" MB_REPS=2 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 40 --csv \
    --log-file gpurun_out/${TAG}_launches_mb.csv python scripts/scan_microbench.py q > gpurun_out/${TAG}_ncu_mb.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_mb.log | cut -c1-150
