#!/bin/bash
# 8-GPU box: bit-identity at 4 and 8 ranks, then the driver-shaped bench at N = 4 and 8 (weak + strong legs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "4 or 8" > gpurun_out/r2_multi_test_n8.log 2>&1
tail -2 gpurun_out/r2_multi_test_n8.log
for N in 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
     bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n${N}.log 2>&1
  echo "N=$N rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench_n${N}.log | tr '\n' ' ') $(grep -o '"slab_bit_identical": [a-z]*' gpurun_out/r2_bench_n${N}.log | head -1)"
done
