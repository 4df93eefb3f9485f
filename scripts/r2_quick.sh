#!/bin/bash
# quick GPU check: parity files + bench of the given variants ("5 6" by default)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_reference_pin.py -m gpu -x -q > gpurun_out/r2_pytest_quick.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r2_pytest_quick.log
for v in ${@:-5 6}; do
  WENO5_FLUX_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 3 --e2e-steps 0 --no-cpu --strong-steps 0 > gpurun_out/r2_bench_quick_v$v.log 2>&1
  echo "variant=$v: $(grep -o '"ms_per_step": [0-9.]*\|"avg_launch_ms": [0-9.]*' gpurun_out/r2_bench_quick_v$v.log | tr '\n' ' ')"
done
