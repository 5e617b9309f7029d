set -u
mkdir -p gpurun_out
echo "== gpu suite"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r2c_suite.log
for pf in 0 2 4 8 16; do
  DFX_IL2_PREFETCH=$pf timeout 300 python scripts/scan_microbench.py pf$pf 2>/dev/null | tail -1 | tee -a gpurun_out/r2c_microbench.jsonl
done
