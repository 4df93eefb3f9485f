"""Throughput of the 3-D extension on one B200: `python scripts/bench3d.py [n] [steps]` -> one JSON line
(cell-updates/s of weno5_advance_3d on the 3-D Orszag-Tang vortex n^3, device-resident, dt recomputed every step;
share of the three sweep kernels from CUDA events)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    ge.build()
    from julia_b200 import problems3d as p3
    from julia_b200.solver3d import Solver3D
    g, q, bx, by, bz = p3.orszag_tang_3d(n)
    with Solver3D(g) as s:
        s.upload(q, bx, by, bz)
        s.advance(2)                                   # warm-up
        t0 = time.perf_counter()
        t, nd = s.advance(steps)
        wall = time.perf_counter() - t0                # weno5_advance_3d returns after a stream synchronize
        s.profile(True)
        s.advance(2)
        ms, nl = s.profile_read()
        s.profile(False)
        divb = float(abs(s.divb()[3:-3, 3:-3, 3:-3]).max())
    cells = n ** 3
    print(json.dumps({"workload": f"orszag_tang_3d {n}^3 periodic", "steps": nd, "ms_per_step": 1e3 * wall / steps,
                      "cell_updates_per_s": cells * steps / wall,
                      "sweep_ms_per_launch_xyz": [3 * m / max(nl, 1) for m in ms],
                      "sweep_share": (sum(ms) / 2) / (1e3 * wall / steps), "sweep_launches_timed": nl,
                      "lib": os.path.basename(os.environ.get("WENO5_LIB", "libweno5mhd.so")),
                      "max_divb": divb, "t_end": t}))


if __name__ == "__main__":
    main()
