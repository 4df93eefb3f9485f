"""
GPU parity tests of the 3-D extension, through the C ABI (weno5_*_3d).  The reference has no 3-D code;
the 3-D step is pinned to its 2-D step by reduction (a state uniform along one axis, in all three orientations) and
compared with the 3-D oracle (itself pinned the same way, tests/test_oracle3d.py) on states that depend on all three
coordinates, for every mix of boundary conditions and on grids whose lines cross CTA-tile boundaries.
"""
import numpy as np
import pytest

from julia_b200 import problems as pr
from julia_b200 import problems3d as p3
from julia_b200.solver3d import Solver3D, classical_RK4_step_3d
from oracle import oracle_c as O

pytestmark = pytest.mark.gpu
I = slice(3, -3)


def compare(got, want, tol):
    assert np.isfinite(got[0]).all()
    for c in range(9):
        s = max(np.abs(want[0][..., c]).max(), 1.0)
        assert np.abs(got[0][..., c] - want[0][..., c]).max() <= tol * s, c
    for a, b in zip(got[1:], want[1:]):
        assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1.0)


@pytest.mark.parametrize("shape", [(70, 5, 4), (5, 36, 4), (4, 5, 36), (40, 37, 35)])
@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 0, 0), (1, 0, 1), (0, 1, 0)])
def test_step_matches_oracle_all_cells(built, shape, bc):
    """one RK4 step on a smooth state, every cell and every face value including the ghost layers"""
    g, q, bx, by, bz = p3.smooth_3d(*shape, seed=sum(shape) + bc[0])
    g.bc_x, g.bc_y, g.bc_z = bc
    q, bx, by, bz = (np.asfortranarray(a.copy()) for a in (q, bx, by, bz))
    O.boundaries_3d(g, q, bx, by, bz)
    with Solver3D(g) as s:
        s.upload(q, bx, by, bz)
        dt, vm = s.timestep()
        dto, vmo = O.timestep_3d(g, q)
        assert abs(dt - dto) <= 1e-12 * dto and np.allclose(vm, vmo, rtol=1e-12, atol=0)
        s.rk4_step(0.8 * dto)
        got = s.download()
    want = O.rk4_step_3d(g, q, bx, by, bz, 0.8 * dto)
    compare(got, want, 1e-11)


@pytest.mark.parametrize("orientation", [0, 1, 2])
def test_reduces_to_the_2d_step_in_every_orientation(built, orientation):
    """uniform along one axis: the 3-D step must return the 2-D step of the same library (the path pinned against the
    reference) and of the 2-D oracle, whichever two axes carry the problem"""
    from julia_b200.solver import Solver
    g2, q, bx, by = pr.orszag_tang(96, ny=64, ysize=2.0 / 3.0)
    with Solver(g2) as s2:
        s2.upload(q, bx, by)
        dt = s2.timestep()[0]
        s2.rk4_step(dt)
        q2, bx2, by2 = s2.download()
    qo, bxo, byo = O.rk4_step_2d(g2, q, bx, by, dt, fast=True)
    g, q3, b0, b1, b2 = p3.embed_2d(g2, q, bx, by, orientation, n3=5)
    got = classical_RK4_step_3d(g, q3, b0, b1, b2, dt)
    for ref_q, ref_bx, ref_by, tol in ((q2, bx2, by2, 1e-12), (qo, bxo, byo, 1e-9)):   # oracle: degenerate faces, DESIGN 6
        for idx in (3, 5, 7):
            for c in (0, 7, 8):
                a = p3.extract_2d(got[0][..., c], orientation, idx)
                assert np.abs(a[I, I] - ref_q[I, I, c]).max() <= tol * np.abs(ref_q[..., c]).max()
            for base in ([1, 2, 3], [4, 5, 6]):
                for m2 in range(3):
                    a = p3.extract_2d(got[0][..., base[p3.vector_component(orientation, m2)]], orientation, idx)
                    assert np.abs(a[I, I] - ref_q[I, I, base[m2]]).max() <= tol
            for ax, ref in ((orientation % 3, ref_bx), ((orientation + 1) % 3, ref_by)):
                assert np.abs(p3.extract_2d(got[1 + ax], orientation, idx)[I, I] - ref[I, I]).max() <= tol


def test_advance_conserves_and_keeps_divb_zero(built):
    """the device-resident driver loop on the 3-D Orszag-Tang vortex: conservation and div B to round-off, agreement
    with the oracle's loop in L1 (degenerate faces: L1 as north_star prescribes for non-smooth problems)"""
    g, q, bx, by, bz = p3.orszag_tang_3d(48, nz=40)
    with Solver3D(g) as s:
        s.upload(q, bx, by, bz)
        s0 = s.totals()
        t, n = s.advance(6)
        assert n == 6 and t > 0
        s1 = s.totals()
        divb = s.divb()
        got = s.download()
        nl = s.launch_count()
    assert nl >= 6 * (4 * 5 + 2)          # per step: 4 x (3 sweeps, EMF + face update, cell update) + CFL reduction
    scale = np.abs(q[I, I, I]).reshape(-1, 9).sum(axis=0)
    for c in (0, 1, 2, 3, 8):
        assert abs(s1[c] - s0[c]) <= 1e-12 * scale[c], c
    for m in range(3):
        assert abs(s1[9 + m] - s0[9 + m]) <= 1e-12 * max(np.abs((bx, by, bz)[m][I, I, I]).sum(), 1.0)
    assert np.abs(divb[I, I, I]).max() <= 1e-12 * np.abs(q[..., 4:7]).max() / g.dx
    qo, bxo, byo, bzo, to = O.advance_3d(g, q, bx, by, bz, 6, fast=True)
    assert abs(t - to) <= 1e-11 * to
    for c in range(9):
        l1 = np.abs(got[0][I, I, I, c] - qo[I, I, I, c]).sum() / max(np.abs(qo[I, I, I, c]).sum(), 1e-300)
        assert l1 <= 1e-8, (c, l1)


def test_alfven_wave_3d_one_period(built):
    """an exact, genuinely 3-D solution on the GPU: the circularly polarised Alfven wave along (1,1,1) run for one period
    with the device-resident driver loop returns to its initial state within the scheme's truncation error (2.5e-4 at
    32^3, the figure of the CPU oracle and of the 2-D scheme at that resolution) and agrees with the oracle's run"""
    n = 32
    g, q, bx, by, bz = p3.alfven_wave_3d(n)
    T = 1.0 / np.sqrt(3.0)
    with Solver3D(g) as s:
        s.upload(q, bx, by, bz)
        t, nsteps = s.advance(100000, tmax=T)
        divb = s.divb()
        got = s.download()
    assert abs(t - T) < 1e-14 and 60 < nsteps < 120
    err = max(np.abs(got[0][I, I, I, c] - q[I, I, I, c]).mean() for c in (1, 2, 3, 4, 5, 6))
    assert err < 3.0e-4, err
    assert np.abs(divb[I, I, I]).max() < 1e-12
    want, tt = (q, bx, by, bz), 0.0
    while tt < T - 1e-14:
        dt = min(O.timestep_3d(g, want[0], fast=True)[0], T - tt)
        want = O.rk4_step_3d(g, *want, dt, fast=True)
        tt += dt
    for c in range(9):
        assert np.abs(got[0][I, I, I, c] - want[0][I, I, I, c]).max() <= 1e-10 * max(np.abs(want[0][..., c]).max(), 1.0), c


def test_blast_3d_outflow_and_tmax(built):
    g, q, bx, by, bz = p3.blast_3d(32, bc=pr.OUTFLOW)
    with Solver3D(g, es_roundtrip=False) as s:
        s.upload(q)                                   # face fields derived on the device (interpolateB_center2face)
        t, n = s.advance(1000, tmax=0.004)
        assert abs(t - 0.004) < 1e-15 and 0 < n < 1000
        got = s.download()
    f = O.center2face_3d(g, q)
    O.boundaries_3d(g, None, *f)
    want = (q, *f)
    tt, steps = 0.0, 0
    while tt < 0.004:
        dt = min(O.timestep_3d(g, want[0], fast=True)[0], 0.004 - tt)
        want = O.rk4_step_3d(g, *want, dt, roundtrip=False, fast=True)
        tt += dt
        steps += 1
    assert steps == n
    for c in range(9):
        l1 = np.abs(got[0][I, I, I, c] - want[0][I, I, I, c]).sum() / max(np.abs(want[0][I, I, I, c]).sum(), 1e-300)
        assert l1 <= 1e-9, (c, l1)


@pytest.mark.parametrize("env", [{"WENO5_3D_SWEEP": "tile"}, {"WENO5_3D_SWEEP": "xtile"}, {"WENO5_3D_SWEEP": "xmarch"},
                                 {"WENO5_3D_CT": "fused"}, {"WENO5_3D_LPL": "32"}, {"WENO5_3D_NSEG": "3"}])
def test_kernel_variants_agree(built, monkeypatch, env):
    """the CTA-tile sweeps (the first version), the thread-per-line x sweep, the fused EMF + face-update kernel and other
    lane / segment counts are the same arithmetic in a different division of labour: two steps must agree with the
    default kernels to round-off"""
    g, q, bx, by, bz = p3.smooth_3d(70, 37, 19, seed=11)
    res = []
    for e in ({}, env):
        for k in ("WENO5_3D_SWEEP", "WENO5_3D_CT", "WENO5_3D_LPL", "WENO5_3D_NSEG"):
            monkeypatch.delenv(k, raising=False)
        for k, v in e.items():
            monkeypatch.setenv(k, v)
        with Solver3D(g) as s:
            s.upload(q, bx, by, bz)
            s.advance(2)
            res.append(s.download())
    for a, b in zip(*res):
        assert np.abs(a - b).max() <= 1e-13 * max(np.abs(b).max(), 1.0)


def test_pure_c_driver_3d(built, tmp_path):
    """tests/abi/c_driver3d.c -- create_3d, upload_3d, advance_3d, divb_3d, download_3d from plain C11, linked against
    the library alone -- on a smooth 3-D state with outflow sides, against the 3-D oracle's driver loop"""
    import os
    import subprocess
    from julia_b200 import lib as L
    from util import ROOT
    g, q, bx, by, bz = p3.smooth_3d(20, 14, 9, seed=3, bc=pr.OUTFLOW)
    q, bx, by, bz = (np.asfortranarray(a.copy()) for a in (q, bx, by, bz))
    O.boundaries_3d(g, q, bx, by, bz)
    exe, fin, fout = str(tmp_path / "c_driver3d"), str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "abi", "c_driver3d.c"), "-o", exe, "-L", os.path.dirname(L.LIB_PATH),
                           "-lweno5mhd", "-Wl,-rpath," + os.path.dirname(L.LIB_PATH)])
    with open(fin, "wb") as f:
        np.array([g.nx, g.ny, g.nz, 3, pr.OUTFLOW], dtype=np.int32).tofile(f)
        np.array([g.xsize, g.ysize, g.zsize]).tofile(f)
        for a in (q, bx, by, bz):
            a.ravel(order="F").tofile(f)
    subprocess.check_call([exe, fin, fout])
    raw = np.fromfile(fout)
    n = int(np.prod(g.shape))
    t, done = raw[0], int(raw[1])
    got_q = raw[2:2 + 9 * n].reshape(g.shape + (9,), order="F")
    got_b = [raw[2 + (9 + m) * n:2 + (10 + m) * n].reshape(g.shape, order="F") for m in range(3)]
    divb = raw[2 + 12 * n:2 + 13 * n].reshape(g.shape, order="F")
    qo, bxo, byo, bzo, to = O.advance_3d(g, q, bx, by, bz, 3)
    assert done == 3 and abs(t - to) <= 1e-12 * to
    compare((got_q, *got_b), (qo, bxo, byo, bzo), 1e-11)
    assert np.abs(divb - O.divb_3d(g, *got_b)).max() <= 1e-12


def test_argument_validation_3d(built):
    import ctypes as C
    from julia_b200 import lib as L
    lib = L.load()
    h = C.c_void_p()
    bad = L.Params3D(3, 8, 8, 0.1, 0.1, 0.1, 5 / 3, 0.8, 1e-6, 1, 1, 1, 1)
    assert lib.weno5_create_3d(C.byref(bad), 0, C.byref(h)) == L.ERR_ARG
    bad = L.Params3D(8, 8, 8, 0.1, 0.1, 0.1, 5 / 3, 0.8, 1e-6, 1, 5, 1, 1)
    assert lib.weno5_create_3d(C.byref(bad), 0, C.byref(h)) == L.ERR_ARG
    assert lib.weno5_rk4_step_3d(None, 0.1) == L.ERR_ARG
    g, q, bx, by, bz = p3.smooth_3d(8, 8, 8)
    with Solver3D(g) as s:
        with pytest.raises(L.Weno5Error):
            s.rk4_step(0.1)                            # no state uploaded
