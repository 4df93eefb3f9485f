#!/bin/bash
# per-kernel durations of one 8192 x 1024 slab (what one of 8 GPUs does in the strong-scaling run, without neighbours)
# against 1/8 of the undivided 8192 x 8192 grid: where the 6 % of strong-scaling loss that is NOT communication sits
mkdir -p gpurun_out
for ny in 1024 8192; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_strong_$ny.csv \
      python bench.py --nx 8192 --ny $ny --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --strong-steps 0 > gpurun_out/ncu_strong_$ny.log 2>&1
  echo "ny=$ny rc=$?"
done
