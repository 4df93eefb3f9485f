"""Launched under torchrun (one rank per GPU): slab-decomposed run vs the same run on one GPU."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from julia_b200 import partition as pt  # noqa: E402
from julia_b200 import problems as pr  # noqa: E402
from julia_b200.solver import Solver  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bc", default="periodic")
    ap.add_argument("--n", type=int, default=192)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    if a.bc == "periodic":
        g, q, bx, by = pr.orszag_tang(a.n)
    else:
        g, q, bx, by = pr.shock_cloud(a.n, a.n // 2)
    off, nl = pt.slab(g.ny, world, rank)
    periodic = g.bc_y == pr.PERIODIC
    uid = pt.broadcast_unique_id(Solver.comm_unique_id, rank, dist)
    with Solver(g, device=local, ny_local=nl, j_offset=off) as s:
        s.comm_init(rank, world, uid)
        cut = lambda x: pt.cut_slab(x, off, nl, periodic=periodic)
        s.upload(cut(q), cut(bx), cut(by))
        dts = []
        for _ in range(a.steps):
            dt = s.timestep()[0]
            dts.append(dt)
            s.rk4_step(dt)
        t, n = s.advance(3)
        qs, bxs, bys = s.download()
    gathered = [None] * world
    dist.all_gather_object(gathered, (qs, bxs, bys))
    if rank == 0:
        with Solver(g, device=local) as s1:
            s1.upload(q, bx, by)
            dts1 = []
            for _ in range(a.steps):
                dt = s1.timestep()[0]
                dts1.append(dt)
                s1.rk4_step(dt)
            s1.advance(3)
            q1, bx1, by1 = s1.download()
        diff = 0.0
        for r, (qq, bb, cc) in enumerate(gathered):
            o, n_ = pt.slab(g.ny, world, r)
            sl = slice(o + 3, o + 3 + n_)
            diff = max(diff, np.abs(qq[:, 3:-3] - q1[:, sl]).max(), np.abs(bb[:, 3:-3] - bx1[:, sl]).max(),
                       np.abs(cc[:, 3:-3] - by1[:, sl]).max())
        ddt = max(abs(x - y) for x, y in zip(dts, dts1))
        print(f"[multi_gpu_check] bc={a.bc} world={world} max|diff|={diff:.3e} max|dt diff|={ddt:.3e}")
        if a.out:
            np.savez(a.out, max_abs_diff=diff, dt_diff=ddt)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
