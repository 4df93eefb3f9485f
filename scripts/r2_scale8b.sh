#!/bin/bash
# 8-GPU box: weak bench at N = 8 with the exchange variants, then the driver-shaped runs at N = 1, 4, 8 (weak + strong legs)
mkdir -p gpurun_out
run() { # N tag env...
  N=$1; tag=$2; shift; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N \
     bench.py --gpus $N --steps 20 --warmup 3 --e2e-steps 0 --no-cpu --strong-steps 0 --no-slab-check > gpurun_out/r2b_bench_n${N}_$tag.log 2>&1
  echo "N=$N $tag: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2b_bench_n${N}_$tag.log | head -1)"
}
timeout 600 python bench.py --steps 20 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2b_bench_n1.log 2>&1
echo "N=1: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2b_bench_n1.log | tr '\n' ' ')"
run 8 peer_all X=1
run 8 nccl_allreduce WENO5_ALLREDUCE=nccl
run 8 wait_kernel WENO5_NO_INKERNEL_WAIT=1
run 8 nccl WENO5_HALO=nccl
for N in 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N \
     bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2b_bench_full_n${N}.log 2>&1
  echo "full N=$N rc=$? $(grep -o '"ms_per_step": [0-9.]*\|"efficiency_vs_n1": [0-9.a-z]*' gpurun_out/r2b_bench_full_n${N}.log | tr '\n' ' ') $(grep -o '"slab_bit_identical": [a-z]*' gpurun_out/r2b_bench_full_n${N}.log | head -1)"
done
