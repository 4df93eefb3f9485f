"""Development aid: device-resident ms/step of the E-only step and of the dual-energy step (both branches +
ES_Switch) on the Orszag-Tang vortex, and how many cells take the E state."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from julia_b200 import problems as pr  # noqa: E402
from julia_b200.solver import Solver  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
g, q, bx, by = pr.orszag_tang(n)
with Solver(g, es_roundtrip=True) as s:
    for dual in (False, True):
        s.set_es_switch(dual, 1e-5)
        s.upload(q, bx, by)
        s.advance(3)
        t0 = time.time()
        t, nd = s.advance(10)
        ms = (time.time() - t0) / 10 * 1e3
        extra = f", {s.es_switch_count()} of {n * n} cells on the E state in the last stage" if dual else ""
        print(f"{n}^2 {'dual-energy' if dual else 'E only   '}: {ms:7.3f} ms/step  {n * n / ms * 1e3:.3e} cell-updates/s{extra}")
