#!/bin/bash
# GPU check of the round-2 defaults: parity suite, smoke, the driver's bench line, the other
# configurations, the ncu launch list and one full capture of the dominant kernel.
#   gpurun --timeout 2400 -- 'bash scripts/r2_gpu_check.sh'
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2b}
echo "== gpu suite"
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_suite.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
echo "== bench (driver command)"
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 4000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
if [ "${QUICK:-0}" = "1" ]; then exit 0; fi
echo "== other configurations"
timeout 900 python scripts/bench_other_configs.py > gpurun_out/${TAG}_other_configs.log 2>&1; tail -40 gpurun_out/${TAG}_other_configs.log
echo "== ncu launch list of the timed region"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-sweep > gpurun_out/${TAG}_ncu_list.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_list.log
echo "== ncu full capture of the coarse-quantizer screening kernel (search shape)"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_coarse -c 2 \
    -o gpurun_out/${TAG}_tc_coarse python bench.py --steps 1 --warmup 1 --no-cpu --no-sweep > gpurun_out/${TAG}_ncu_full_tc.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full_tc.log
echo "== ncu full capture of the scan kernel"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:scan_pq_il2 -c 2 \
    -o gpurun_out/${TAG}_scan_pq_il2 python bench.py --steps 1 --warmup 1 --no-cpu --no-sweep > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
