"""
Slab-decomposed run vs the same run on one GPU: must be BIT-identical (SURVEY.md 8e -- every rank does the
same per-cell arithmetic on identical data, the dt maximum is order-independent).

Used by tests/test_gpu_multi.py (through scripts/multi_gpu_check.py), and by bench.py as a pre-check of
every multi-GPU run whose outcome lands in the JSON line ("slab_bit_identical").  Every rank advances its
slab (4 ghost rows exchanged per RK stage, max-allreduce per step); rank 0 then repeats the run on the whole
grid on its own GPU and compares all gathered rows, face fields and time steps.
"""
import numpy as np

from . import partition as pt
from . import problems as pr
from .solver import Solver

CASES = {
    # name: (constructor, nx, rows per rank (0: 1024 rows split over the ranks), steps, advance steps)
    "orszag_tang_512x1024": ("orszag_tang", 512, 0, 2, 3),
    "blast_512x1024": ("blast", 512, 0, 2, 3),
    "orszag_tang_1024row_slabs": ("orszag_tang", 256, 1024, 1, 2),
    "shock_cloud_outflow": ("shock_cloud", 256, 0, 2, 3),
    # dual-energy mode (both branches + live ES_Switch, Weno5_2D.jl:1734-1768): the switch flags of the neighbours' edge
    # rows travel with the halo
    "blast_dual_energy": ("blast", 256, 0, 2, 2, True),
}


def _problem(name, nx, ny):
    if name == "shock_cloud":
        return pr.shock_cloud(nx, ny, xsize=1.0, ysize=float(ny) / nx)
    return getattr(pr, name)(nx, ny=ny, ysize=float(ny) / nx)


def run_case(case, rank, world, local, dist):
    """Returns (max |difference|, max |dt difference|) on rank 0, (None, None) elsewhere."""
    ctor, nx, rows, steps, adv = CASES[case][:5]
    dual = len(CASES[case]) > 5 and CASES[case][5]
    ny = rows * world if rows else 1024
    g, q, bx, by = _problem(ctor, nx, ny)
    off, nl = pt.slab(g.ny, world, rank)
    periodic = g.bc_y == pr.PERIODIC
    uid = pt.broadcast_unique_id(Solver.comm_unique_id, rank, dist)
    with Solver(g, device=local, ny_local=nl, j_offset=off) as s:
        s.comm_init(rank, world, uid)
        if dual:
            s.set_es_switch(True, 1e-5)
        cut = lambda x: pt.cut_slab(x, off, nl, periodic=periodic)       # noqa: E731
        s.upload(cut(q), cut(bx), cut(by))
        dts = []
        for _ in range(steps):
            dt = s.timestep()[0]
            dts.append(dt)
            s.rk4_step(dt)
        t_end, _ = s.advance(adv)
        dts.append(t_end)
        qs, bxs, bys = s.download()
    gathered = [None] * world
    dist.all_gather_object(gathered, (qs[:, 3:-3], bxs[:, 3:-3], bys[:, 3:-3]))
    if rank != 0:
        return None, None
    with Solver(g, device=local) as s1:
        if dual:
            s1.set_es_switch(True, 1e-5)
        s1.upload(q, bx, by)
        dts1 = []
        for _ in range(steps):
            dt = s1.timestep()[0]
            dts1.append(dt)
            s1.rk4_step(dt)
        t_end, _ = s1.advance(adv)
        dts1.append(t_end)
        q1, bx1, by1 = s1.download()
    diff = 0.0
    for r, (qq, bb, cc) in enumerate(gathered):
        o, n_ = pt.slab(g.ny, world, r)
        sl = slice(o + 3, o + 3 + n_)
        diff = max(diff, float(np.abs(qq - q1[:, sl]).max()), float(np.abs(bb - bx1[:, sl]).max()),
                   float(np.abs(cc - by1[:, sl]).max()))
    ddt = max(abs(x - y) for x, y in zip(dts, dts1))
    return diff, ddt


def run_all(rank, world, local, dist, cases=None):
    """All cases; on rank 0 a dict {"slab_bit_identical": bool, "ranks": N, "cases": {...}}."""
    out = {}
    for case in (cases or CASES):
        d, ddt = run_case(case, rank, world, local, dist)
        if rank == 0:
            out[case] = {"max_abs_diff": d, "max_dt_diff": ddt}
    if rank != 0:
        return None
    ok = all(v["max_abs_diff"] == 0.0 and v["max_dt_diff"] == 0.0 for v in out.values())
    return {"slab_bit_identical": bool(ok), "ranks": world, "cases": out}
