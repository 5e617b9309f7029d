#!/bin/bash
# First GPU call of round 2: validate and time everything that was written after the GPU budget of
# round 1 was spent (all of it passes on the CPU emulation of the library, none of it has run on
# hardware): scan_variant 2 / 3, prep_variant 2, rerank_variant 2, DFX_IL2_THREADS=384, DFX_GRAPHS=1.
# Run from the repo root (about 30 GPU-minutes):
#   gpurun --timeout 2700 -- 'bash scripts/r2_validate_scan_v3.sh'
# Everything lands in gpurun_out/ (copy what is worth keeping into profiles/).
set -u
mkdir -p gpurun_out
echo "== parity of the experimental variants (scan 2/3, prep 2, rerank 2) against the oracle"
DFX_EXPERIMENTAL=1 timeout 600 python -m pytest tests -m gpu -q -x -k "scan_variant_2 or interleaved or ivf_matches or tensor_core_coarse or flat_tensor_core" 2>&1 | tail -15 | tee gpurun_out/r2_v3_parity.log
if ! grep -q " passed" gpurun_out/r2_v3_parity.log || grep -q "failed" gpurun_out/r2_v3_parity.log; then
    echo "the experimental variants are NOT green: stop here, read gpurun_out/r2_v3_parity.log"
    exit 1
fi
echo "== whole gpu suite with variant 2 as the default layout of every new IVF-PQ index"
DFX_SCAN_VARIANT=2 DFX_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2_v3_suite.log
echo "== bench: default kernels, then every experimental variant on the SAME built shards (one 1B build)"
timeout 1200 python bench.py --steps 20 --warmup 3 --variant-sweep > gpurun_out/r2_bench_variants.json 2> gpurun_out/r2_bench_variants.err
tail -c 3000 gpurun_out/r2_bench_variants.json
echo "== the 12-warp CTA shape of scan 2 needs its own process (read once from the environment)"
DFX_SCAN_VARIANT=2 DFX_IL2_THREADS=384 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_v3_384.json 2> gpurun_out/r2_bench_v3_384.err
tail -c 1500 gpurun_out/r2_bench_v3_384.json
echo "== small batches with and without CUDA-graph replay (DFX_GRAPHS=1, experimental)"
DFX_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_api.py -m gpu -q -k graph_replay 2>&1 | tail -3 | tee gpurun_out/r2_graphs_parity.log
# (100 M vectors: the comparison is about launch overhead, a 1 B build per arm is not needed)
timeout 600 python bench.py --nvec 100000000 --steps 20 --warmup 3 --sweep --no-cpu > gpurun_out/r2_bench_sweep_eager.json 2> gpurun_out/r2_bench_sweep_eager.err
DFX_GRAPHS=1 timeout 600 python bench.py --nvec 100000000 --steps 20 --warmup 3 --sweep --no-cpu > gpurun_out/r2_bench_sweep_graphs.json 2> gpurun_out/r2_bench_sweep_graphs.err
grep -o '"qps_by_batch": {[^}]*}' gpurun_out/r2_bench_sweep_eager.json gpurun_out/r2_bench_sweep_graphs.json
echo "== the other configurations: default, then flat through the tensor-core path + 8 vectors in flight in the row scans"
timeout 600 python scripts/bench_other_configs.py > gpurun_out/r2_other_configs.log 2>&1; tail -5 gpurun_out/r2_other_configs.log
DFX_FLAT_TC=1 DFX_ROWS_INFLIGHT=8 timeout 600 python scripts/bench_other_configs.py > gpurun_out/r2_other_configs_flat_tc.log 2>&1; tail -5 gpurun_out/r2_other_configs_flat_tc.log
echo "== ncu: one full capture of the new scan kernel (bench.py opens the profiler window around the timed region)"
DFX_SCAN_VARIANT=2 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:scan_pq_il2 -c 2 \
    -o gpurun_out/r2_scan_pq_il2 python bench.py --steps 1 --warmup 1 > gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_ncu.log
