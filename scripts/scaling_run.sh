#!/bin/bash
# usage: scaling_run.sh NGPU NX NY_PER_GPU TAG   -- one bench.py run at the given per-GPU slab size
N=$1; NX=$2; NY=$3; TAG=$4
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --nx $NX --ny $NY --steps 10 --warmup 3 --e2e-steps 1 --no-cpu > gpurun_out/scale_$TAG.log 2>&1
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
      bench.py --gpus $N --nx $NX --ny $NY --steps 10 --warmup 3 --e2e-steps 1 --no-cpu > gpurun_out/scale_$TAG.log 2>&1
fi
echo "rc=$? $TAG"
grep "^{" gpurun_out/scale_$TAG.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$TAG', d['n_gpus'], 'gpus', '%.4e cell-upd/s' % d['value'], '%.3f ms/step' % d['ms_per_step'], d['config']['workload'])"
