#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_quick2.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r2_pytest_quick2.log
timeout 300 python bench.py --steps 20 --warmup 3 --e2e-steps 0 --no-cpu --strong-steps 0 > gpurun_out/r2_bench_quick2.log 2>&1
echo "4096^2: $(grep -o '"ms_per_step": [0-9.]*\|"avg_launch_ms": [0-9.]*' gpurun_out/r2_bench_quick2.log | tr '\n' ' ')"
timeout 300 python bench.py --nx 8192 --ny 1024 --steps 20 --warmup 3 --e2e-steps 0 --no-cpu --strong-steps 0 > gpurun_out/r2_bench_quick2_slab.log 2>&1
echo "8192x1024 slab: $(grep -o '"ms_per_step": [0-9.]*\|"avg_launch_ms": [0-9.]*' gpurun_out/r2_bench_quick2_slab.log | tr '\n' ' ')"
timeout 300 python bench.py --nx 2048 --ny 2048 --steps 20 --warmup 3 --e2e-steps 0 --no-cpu --strong-steps 0 > gpurun_out/r2_bench_quick2_2048.log 2>&1
echo "2048^2: $(grep -o '"ms_per_step": [0-9.]*\|"avg_launch_ms": [0-9.]*' gpurun_out/r2_bench_quick2_2048.log | tr '\n' ' ')"
