"""GPU bring-up check (development aid): compares intermediate and final fields of the CUDA path
with the CPU oracle on small grids and prints error tables.  Run on a B200 via gpurun."""
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
from julia_b200 import problems as pr  # noqa: E402
from julia_b200.solver import Solver  # noqa: E402
from oracle import oracle_c as O  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def check_sweeps(name, g, q, bx, by):
    q8 = np.asfortranarray(q[..., [0, 1, 2, 3, 4, 5, 6, 8]])
    dqx, fsy = O.sweep_2d(g, q8, 0)
    dqy, gsx = O.sweep_2d(g, q8, 1)
    with Solver(g) as s:
        s.upload(q, bx, by)
        dt = s.timestep()[0]
        s.rk4_step(dt)   # after the step the buffers hold stage-4 data; re-run a stage-1-only check below
    # stage-1 intermediate: run one step on a fresh handle and fetch QA? rhs/fsy are overwritten by later
    # stages, so use a tiny dt: stages 2-4 see nearly the same state -> compare loosely, then exact via step.
    with Solver(g) as s:
        s.upload(q, bx, by)
        s.rk4_step(1e-14)
        i = slice(3, -3)
        rows = []
        six = [0, 1, 2, 3, 6, 7]
        for m, c in enumerate(six):
            r = s.debug_fetch(f"rhs{m}")
            rows.append((f"rhs{m}", rel(r[i, i], dqx[i, i, c])))
        f = s.debug_fetch("fsy")
        rows.append(("fsy", rel(f[2:-2, 2:-2], fsy[2:-2, 2:-2])))
        gq = s.debug_fetch("gsx")
        rows.append(("gsx", rel(gq[2:-2, 2:-2], gsx[2:-2, 2:-2])))
    print(f"[{name}] sweep intermediates (dt=1e-14 so all stages see ~q0):")
    for k, v in rows:
        print(f"    {k:6s} rel err {v:.3e}")


def check_steps(name, g, q, bx, by, nsteps, roundtrip=True, full=False):
    with Solver(g, es_roundtrip=roundtrip) as s:
        s.upload(q, bx, by)
        dt, vx, vy = s.timestep()
        dto = O.timestep_2d(g, q)
        qo, bxo, byo = q, bx, by
        t0 = time.time()
        for _ in range(nsteps):
            s.rk4_step(dt)
        tg = time.time() - t0
        for _ in range(nsteps):
            qo, bxo, byo = O.rk4_step_2d(g, qo, bxo, byo, dt)
        qg, bxg, byg = s.download()
        dt2 = s.timestep()
        dto2 = O.timestep_2d(g, qo)
        divb = np.abs(s.divb()).max()
    i = slice(None) if full else slice(3, -3)
    errs = [rel(qg[i, i, c], qo[i, i, c]) for c in range(9)]
    print(f"[{name}] {nsteps} steps (roundtrip={roundtrip}, full={full}): dt gpu {dt:.15e} oracle {dto[0]:.15e}")
    print("    q rel err per comp:", " ".join(f"{e:.1e}" for e in errs))
    print(f"    bxb {rel(bxg[i, i], bxo[i, i]):.2e} byb {rel(byg[i, i], byo[i, i]):.2e}  "
          f"dt-after gpu {dt2[0]:.12e} oracle {dto2[0]:.12e}  max|divB| {divb:.2e}  gpu wall {tg:.3f}s")
    return max(errs)


def check_1d():
    g, q = pr.rj95_2a(256)
    with Solver(g) as s:
        s.upload(q)
        dt, vx, _ = s.timestep()
        vo, dto = O.vmax_1d(g, q)
        qo = q
        for _ in range(20):
            s.rk4_step(dt)
            qo = O.rk4_step_1d(g, qo, dt)
        qg = s.download()
    print(f"[1d rj2a] dt gpu {dt:.15e} oracle {dto:.15e}; 20 steps rel err per comp:",
          " ".join(f"{rel(qg[:, c], qo[:, c]):.1e}" for c in range(9)))


if __name__ == "__main__":
    probs = {
        "alfven48": pr.alfven_wave(48),
        "OT64": pr.orszag_tang(64),
        "blast40": pr.blast(40),
        "rand40x24": pr.random_state(40, 24),
        "shockcloud64x16": pr.shock_cloud(64, 16),
    }
    for name, (g, q, bx, by) in probs.items():
        check_sweeps(name, g, q, bx, by)
    for name, (g, q, bx, by) in probs.items():
        check_steps(name, g, q, bx, by, 3, full=name.startswith("shock"))
    g, q, bx, by = pr.alfven_wave(64)
    check_steps("alfven64 no-roundtrip", g, q, bx, by, 10, roundtrip=False)
    g, q, bx, by = pr.shock_cloud(16, 8)
    check_steps("shockcloud16x8 full", g, q, bx, by, 2, full=True)
    check_1d()
    # larger grid sanity + timing
    g, q, bx, by = pr.orszag_tang(1024)
    with Solver(g) as s:
        s.upload(q, bx, by)
        s.advance(2)
        t0 = time.time()
        t, n = s.advance(10)
        dtw = time.time() - t0
        print(f"[OT1024] 10 steps in {dtw:.3f}s -> {1024 * 1024 * 10 / dtw:.3e} cell-updates/s; t={t:.5f}; "
              f"max|divB| {np.abs(s.divb()).max():.2e}; launches {s.launch_count()}")
