"""Generates tests/golden/*.npz: outputs of the CPU oracle (strict build) for small seeded
cases.  The reference itself ships no golden vectors and cannot run here, so
these fixtures freeze the oracle; tests check the oracle against them (-m "not gpu") and the
CUDA path against them (-m gpu).  Re-run only when the oracle is deliberately changed."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from julia_b200 import problems as pr  # noqa: E402
from oracle import oracle_c as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES_2D = {
    # name: (constructor, kwargs, steps)
    "ot24": (pr.orszag_tang, dict(n=24), 3),
    "alfven32": (pr.alfven_wave, dict(n=32), 5),
    "blast24": (pr.blast, dict(n=24), 3),
    "shockcloud16x8": (pr.shock_cloud, dict(nx=16, ny=8), 2),
    "random24x16": (pr.random_state, dict(nx=24, ny=16, seed=1234), 2),
}
CASES_1D = {
    "rj95_2a_64": (pr.rj95_2a, dict(nx=64), 20),
    "briowu_64": (pr.brio_wu, dict(nx=64), 20),
}


def run_2d(name):
    ctor, kw, steps = CASES_2D[name]
    g, q, bx, by = ctor(**kw)
    dt = O.timestep_2d(g, q)[0]
    for _ in range(steps):
        q, bx, by = O.rk4_step_2d(g, q, bx, by, dt)
    return dict(q=q, bxb=bx, byb=by, dt=dt, steps=steps)


def run_1d(name):
    ctor, kw, steps = CASES_1D[name]
    g, q = ctor(**kw)
    dt = O.vmax_1d(g, q)[1]
    for _ in range(steps):
        q = O.rk4_step_1d(g, q, dt)
    return dict(q=q, dt=dt, steps=steps)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for n in CASES_2D:
        np.savez_compressed(os.path.join(OUT, n + ".npz"), **run_2d(n))
    for n in CASES_1D:
        np.savez_compressed(os.path.join(OUT, n + ".npz"), **run_1d(n))
    print("wrote", sorted(os.listdir(OUT)))
