#!/bin/bash
# development aid: time every flux-kernel launch variant on one B200 and check it against the oracle
mkdir -p gpurun_out
N=${1:-2048}
for v in 2 4; do
  export WENO5_FLUX_VARIANT=$v
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v$v.log 2>&1 || { echo "variant $v: SMOKE FAILED"; tail -3 gpurun_out/smoke_v$v.log; continue; }
  python bench.py --nx $N --ny $N --steps 10 --warmup 3 --e2e-steps 1 --no-cpu > gpurun_out/bench_v$v.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_v$v.log') if x.startswith('{')]
if not l: print('variant $v: bench failed'); raise SystemExit
d=json.loads(l[-1]); r=d['roofline']
print(f"variant $v: {d['value']:.4e} cell-upd/s  {d['ms_per_step']:.3f} ms/step  flux avg {r['avg_launch_ms']:.4f} ms  share {r['share_of_step']:.3f}  fp64 frac {r['fp64']['frac']:.3f}  clocks {d['clocks']['sm_mhz']} {d['clocks'].get('power_w_max')}")
PY
done
