#!/bin/bash
# GPU pass: full -m gpu suite, then the bench (device-resident + e2e), both logged under gpurun_out/
tag=${1:-x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_pytest_$tag.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_$tag.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_bench_$tag.log 2>&1
tail -4 gpurun_out/r2_pytest_$tag.log
grep -h -o '"ms_per_step": [0-9.]*\|"avg_launch_ms": [0-9.]*\|"gpu_launches": [0-9]*' gpurun_out/r2_bench_$tag.log
grep -h "worst relative" gpurun_out/r2_pytest_$tag.log
