#!/bin/bash
# N-GPU bench through IndexClient over the NCCL search plane (both arms, as the driver runs them)
#   gpurun --gpus N --timeout 1500 -- 'N=N bash scripts/r2_multi.sh'
set -u
N=${N:-2}
mkdir -p gpurun_out
if [ "${WITH_TESTS:-0}" = "1" ]; then
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core or flat" 2>&1 | tail -3 | tee gpurun_out/multi${N}_suite.log
fi
timeout ${BENCH_TIMEOUT:-360} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/multi${N}_bench.json 2> gpurun_out/multi${N}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/multi${N}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["roofline"]["ms_per_launch"], d["clocks"])
print(d["config"].get("qps_by_batch"), d["config"].get("e2e_qps_by_batch"))
PY
grep -v "^\[bench" gpurun_out/multi${N}_bench.err | grep -iv "nccl info" | head -8
if [ "${WITH_REF:-0}" = "1" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
      bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/multi${N}_ref.json 2> gpurun_out/multi${N}_ref.err
  cut -c1-600 gpurun_out/multi${N}_ref.json
fi
