"""Host-side logic of the multi-GPU path on CPU: slab bookkeeping, and -- with two gloo
ranks -- that a slab-split run of the oracle (each rank advances its slab plus an overlap
wide enough for the domain of dependence of one RK4 step) reproduces the monolithic run
bit for bit, which is the property the NCCL halo exchange has to preserve on the GPUs."""
import os
import sys

import numpy as np
import pytest

from julia_b200 import partition as pt
from julia_b200 import problems as pr
from util import ROOT


def test_slab_covers_the_grid():
    for ny, P in ((64, 1), (64, 2), (100, 3), (8192, 8), (16384, 8), (67, 4)):
        rows = []
        for r in range(P):
            off, n = pt.slab(ny, P, r)
            assert off == len(rows)
            rows += list(range(off, off + n))
        assert rows == list(range(ny))
    with pytest.raises(ValueError):
        pt.slab(20, 4, 0)


def test_neighbours():
    assert pt.neighbours(0, 1, True) == (-1, -1)
    assert pt.neighbours(0, 4, True) == (3, 1) and pt.neighbours(3, 4, True) == (2, 0)
    assert pt.neighbours(0, 4, False) == (-1, 1) and pt.neighbours(3, 4, False) == (2, -1)
    assert pt.neighbours(0, 2, True) == (1, 1)


def test_cut_and_assemble_roundtrip():
    g, q, bx, by = pr.orszag_tang(24)
    slabs = [pt.cut_slab(q, *pt.slab(24, 3, r), periodic=True) for r in range(3)]
    assert all(s.shape == (g.Nx, 8 + 6, 9) for s in slabs)
    assert np.array_equal(pt.assemble(slabs), q)
    # ghost rows of a slab are the neighbour's interior rows
    assert np.array_equal(slabs[1][:, :3], slabs[0][:, -6:-3])
    assert np.array_equal(slabs[0][:, :3], slabs[2][:, -6:-3])       # periodic wrap


def test_slab_ic_equals_cut_of_global_ic():
    n, P = 32, 2
    g, q, bx, by = pr.orszag_tang(n)
    for r in range(P):
        off, nl = pt.slab(n, P, r)
        gs, qs, bxs, bys = pr.orszag_tang(n, ny=nl, ysize=nl / n, y0=off / n)
        ref = pt.cut_slab(q, off, nl, periodic=True)
        assert np.abs(qs[:, 3:-3] - ref[:, 3:-3]).max() < 1e-13
        assert np.abs(bxs[:, 3:-3] - pt.cut_slab(bx, off, nl, periodic=True)[:, 3:-3]).max() < 1e-13


def _worker(rank, world, port, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import oracle_c as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, overlap = 48, 22                       # 4 stages x (3 WENO + 2 CT) rows of dependence
    g, q, bx, by = pr.random_state(24, n, seed=11)                   # periodic: the overlap rows are true images
    dt = 0.5 * O.timestep_2d(g, q)[0]
    off, nl = pt.slab(n, world, rank)
    # the pseudo-domain's own y boundary (outflow) is wrong, but its error travels at most
    # 3 + 4 stages x 4 rows = 19 rows per step and never reaches the owned rows
    gs = pr.Grid(nx=g.nx, ny=nl + 2 * (overlap - 3), xsize=g.xsize, ysize=g.dy * (nl + 2 * (overlap - 3)),
                 gam=g.gam, bc_x=g.bc_x, bc_y=pr.OUTFLOW)
    cut = lambda a: pt.cut_slab(a, off, nl, ghost=overlap, periodic=True)
    q4, bx4, by4 = O.rk4_step_2d(gs, cut(q), cut(bx), cut(by), dt)
    mine = np.ascontiguousarray(q4[:, overlap:overlap + nl])
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        full = O.rk4_step_2d(g, q, bx, by, dt)[0]
        got = np.concatenate(gathered, axis=1)
        np.save(out, np.abs(got - full[:, 3:-3]).max())
    dist.barrier()
    dist.destroy_process_group()


def test_slab_split_oracle_is_bit_identical_gloo(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "err.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert float(np.load(out)) == 0.0
