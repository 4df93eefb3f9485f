#!/bin/bash
# final scaling record of round 2 on one 8-GPU box: the driver-shaped bench at N = 1, 2, 4, 8
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c_bench_n1.log 2>&1
echo "N=1 rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2c_bench_n1.log | tr '\n' ' ')"
for N in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N \
     bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2c_bench_n${N}.log 2>&1
  echo "N=$N rc=$? $(grep -o '"ms_per_step": [0-9.]*\|"efficiency_vs_n1": [0-9.a-z]*' gpurun_out/r2c_bench_n${N}.log | tr '\n' ' ') $(grep -o '"slab_bit_identical": [a-z]*' gpurun_out/r2c_bench_n${N}.log | head -1)"
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2c_bench_ref.log 2>&1
echo "reference arm: $(grep -o '"value": [0-9.]*' gpurun_out/r2c_bench_ref.log | head -1) $(grep -o '"cores": [0-9]*' gpurun_out/r2c_bench_ref.log | head -1)"
