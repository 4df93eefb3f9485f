#!/bin/bash
for n in 4096 2048; do
for w in 2 4 6 8 12; do
  export WENO5_WAVES=$w WENO5_MINPER=8
  python bench.py --nx $n --ny $n --steps 10 --warmup 3 --e2e-steps 1 --no-cpu > gpurun_out/b.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/b.log') if x.startswith('{')]
d=json.loads(l[-1]); print("n=$n waves=$w", "%.4e" % d['value'], "%.3f ms" % d['ms_per_step'])
PY
done; done
