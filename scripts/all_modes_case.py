"""Small end-to-end case that launches every kernel of the path once or twice (warp-specialised / TMEM-stash /
single-role sweeps through WENO5_FLUX_VARIANT, CT, boundary fill, fused CFL reduction, graph replay, 1-D solver,
dual-energy mode in 2-D and 1-D, snapshot hook, pipelined host step).  Written as the target of a compute-sanitizer
pass; the tool is closed on this GPU pool ("runs under it have left GPUs needing a reset"), so out-of-bounds
accesses are covered by the parity tests instead (ragged sizes, ghost cells compared, bit-identity of slab and
pipelined runs whose planes have different shapes)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from julia_b200 import problems as pr  # noqa: E402
from julia_b200.solver import Pipe, Solver  # noqa: E402

g, q, bx, by = pr.orszag_tang(80, ny=72, ysize=0.9)
with Solver(g) as s:
    s.upload(q, bx, by)
    s.advance(3)
    s.totals(); s.divb()
    out = s.download()
    assert np.isfinite(out[0]).all()
g, q, bx, by = pr.shock_cloud(72, 40, xsize=1.8, ysize=1.0)
with Solver(g) as s:
    s.upload(q, bx, by)
    dt = s.timestep()[0]
    s.rk4_step(dt)
    s.set_es_switch(1, 1e-5)
    s.rk4_step(dt)
    s.es_switch_count()
    s.arr2img(0, None, ncolours=64)
g1, q1 = pr.rj95_2a(200)
with Solver(g1) as s:
    s.upload(q1)
    s.advance(3)
    s.set_es_switch(2, 1e-5)
    s.rk4_step(1e-4)
    s.es_switch_mask()
g, q, bx, by = pr.orszag_tang(64, ny=160, ysize=2.5)
with Pipe(g, rows_per_slab=48) as p:
    p.step(q, bx, by, 1e-4)
print("all modes ran")
