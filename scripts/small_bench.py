"""ms/step of the launch-bound configurations (BASELINE configs 1 and 2, plus 1024^2 / 2048^2), device-resident
weno5_advance, with and without the captured-graph replay (WENO5_NO_GRAPH=1).  The 1-D line runs as one
cluster-resident kernel (weno5_line1d.cu); WENO5_NO_LINE1D=1 in the environment gives the multi-kernel figures."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401  (device synchronisation only)

from julia_b200 import problems as pr  # noqa: E402
from julia_b200.solver import Solver  # noqa: E402


def run(name, make, steps):
    out = {}
    for graph in (0, 1):
        if graph:
            os.environ.pop("WENO5_NO_GRAPH", None)
        else:
            os.environ["WENO5_NO_GRAPH"] = "1"
        a = make()
        g = a[0]
        with Solver(g) as s:
            s.upload(*a[1:])
            s.advance(10)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s.advance(steps)
            torch.cuda.synchronize()
            out[graph] = (time.perf_counter() - t0) / steps * 1e3
    cells = g.nx * g.ny
    print(f"{name:28s} {cells:9d} cells: {out[0]:8.4f} ms/step enqueued, {out[1]:8.4f} ms/step graph replay "
          f"({cells / (out[1] * 1e-3):.3e} cell-updates/s)")


if __name__ == "__main__":
    run("1-D RJ95-2a 1024 (config 1)", lambda: pr.rj95_2a(1024), 2000)
    run("alfven 512^2 (config 2)", lambda: pr.alfven_wave(512), 500)
    run("orszag_tang 1024^2", lambda: pr.orszag_tang(1024), 200)
    run("orszag_tang 2048^2 (config 3)", lambda: pr.orszag_tang(2048), 100)
