"""Development aid: e2e throughput of the pipelined host step vs rows per slab (pinned host arrays)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from julia_b200 import lib as L, problems as pr
from julia_b200.solver import Pipe, Solver
lib = L.load()
def pinned(shape):
    n = int(np.prod(shape)); p = C.c_void_p(); L.check(lib.weno5_host_alloc(C.byref(p), n * 8))
    return np.frombuffer((C.c_double * n).from_address(p.value), dtype=np.float64).reshape(shape, order="F")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g, q, bx, by = pr.orszag_tang(n)
a = [pinned(x.shape) for x in (q, bx, by)]; b = [pinned(x.shape) for x in (q, bx, by)]
for dst, src in zip(a, (q, bx, by)): dst[...] = src
with Solver(g) as s:
    s.upload(q, bx, by); dt0 = s.timestep()[0]
rows_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else (128, 256, 512, 1024)
ref = None
for rows in rows_list:
    with Pipe(g, rows_per_slab=rows, es_roundtrip=False) as p:
        i, o = a, b
        *_, dtn = p.step(*i, dt0, out=tuple(o)); i, o = o, i
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(5):
            *_, dtn = p.step(*i, dtn, out=tuple(o)); i, o = o, i
        torch.cuda.synchronize(); t = (time.time() - t0) / 5
    # every slab size must give the same bits (the windows carry the whole domain of dependence of a step)
    res = [x.copy() for x in i]
    if ref is None:
        ref = res
    same = all(np.array_equal(x, y) for x, y in zip(res, ref))
    print(f"rows {rows}: {t*1e3:.2f} ms/step  {n*n/t:.3e} cell-upd/s  direct={'WENO5_PIPE_DIRECT' in os.environ}  same_bits_as_first={same}")
    for dst, src in zip(a, (q, bx, by)): dst[...] = src
