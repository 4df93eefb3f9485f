"""Development aid: relative drift of the conserved totals on the GPU for several grid sizes."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from julia_b200 import problems as pr
from julia_b200.solver import Solver
for n in (128, 256, 512, 1024, 2048):
    g, q, bx, by = pr.orszag_tang(n)
    with Solver(g, es_roundtrip=False) as s:
        s.upload(q, bx, by)
        t0 = s.totals()
        s.advance(5)
        t1 = s.totals()
    cells = float(n) * n
    rel = [(t1[c] - t0[c]) / max(abs(t0[c]), cells * 1e-3) for c in (0, 1, 2, 4, 5, 8, 9, 10)]
    print(n, " ".join(f"{r:+.2e}" for r in rel))
