#!/bin/bash
# multi-GPU pass on N GPUs: bit-identity test, then the weak bench with the three halo-exchange variants
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_multi_test_n$N.log 2>&1
tail -2 gpurun_out/r2_multi_test_n$N.log
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 20 --warmup 3 --e2e-steps 0 --no-cpu --strong-steps 0 --no-slab-check > gpurun_out/r2_bench_n${N}_$tag.log 2>&1
  echo "$tag: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench_n${N}_$tag.log | head -1) $(grep -o '"halo_exchange": "[^"]*"' gpurun_out/r2_bench_n${N}_$tag.log)"
}
run peer_all X=1
run peer_nccl_allreduce WENO5_ALLREDUCE=nccl
run peer_wait_kernel WENO5_NO_INKERNEL_WAIT=1
run nccl WENO5_HALO=nccl
