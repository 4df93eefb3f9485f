#!/bin/bash
# round-2 first GPU pass: parity of the warp-specialised kernel, then a variant sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
for v in "4 0" "5 0" "5 1" "5 2"; do
  set -- $v
  WENO5_FLUX_VARIANT=$1 WENO5_WS_SPLIT=$2 timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench_v$1_s$2.log 2>&1
done
tail -3 gpurun_out/r2_pytest1.log
grep -h -o '"ms_per_step": [0-9.]*\|"avg_launch_ms": [0-9.]*' gpurun_out/r2_bench_v*.log
