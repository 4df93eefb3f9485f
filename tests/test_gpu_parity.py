"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle
on the same inputs, against the committed golden fixtures, and -- at BASELINE.json's sizes --
through size-independent properties (conservation, div B, symmetry).

Tolerances (BASELINE.json north_star): smooth problems relative L-inf <= 1e-10; shock problems
are compared in relative L1 with the tolerance stated at the assert."""
import ctypes as C
import numpy as np
import pytest

from julia_b200 import lib as L
from julia_b200 import problems as pr
from julia_b200 import solver as S
from julia_b200.solver import Solver
from oracle import oracle_c as O
from util import degenerate_distance_2d, golden, interior, rel_l1, rel_linf

pytestmark = pytest.mark.gpu

ALL9 = range(9)
SMOOTH_TOL = 1e-10


@pytest.fixture(scope="module", autouse=True)
def _require_gpu(built):
    # fail loudly: on a GPU box a missing device / library is an error, never a skip
    assert L.load().weno5_device_count() > 0, "no sm_100 GPU visible: the CUDA path cannot run"


def run_gpu(g, q, bx, by, steps, dt=None, **kw):
    with Solver(g, **kw) as s:
        s.upload(q, bx, by)
        if dt is None:
            dt = s.timestep()[0]
        for _ in range(steps):
            s.rk4_step(dt)
        return s.download() + (dt,)


def run_oracle(g, q, bx, by, steps, dt):
    for _ in range(steps):
        q, bx, by = O.rk4_step_2d(g, q, bx, by, dt)
    return q, bx, by


# ------------------------------------------------------------------ oracle parity, small grids
@pytest.mark.parametrize("ctor,kw,steps", [
    (pr.alfven_wave, dict(n=64), 10),
    (pr.random_state, dict(nx=72, ny=40, seed=3), 5),
    (pr.blast, dict(n=48), 5),
])
def test_smooth_problems_match_oracle(ctor, kw, steps):
    g, q, bx, by = ctor(**kw)
    qg, bxg, byg, dt = run_gpu(g, q, bx, by, steps)
    assert abs(dt - O.timestep_2d(g, q)[0]) <= 1e-14 * dt
    qo, bxo, byo = run_oracle(g, q, bx, by, steps, dt)
    for c in ALL9:
        scale = max(np.abs(qo[..., c]).max(), 1e-3)       # vz,Bz of the blast are round-off sized
        assert np.abs(qg[..., c] - qo[..., c]).max() <= SMOOTH_TOL * scale, c
    assert rel_linf(bxg, bxo) <= SMOOTH_TOL and rel_linf(byg, byo) <= SMOOTH_TOL


def test_alfven_512_config2():
    """BASELINE config 2: smooth Alfven wave at 512^2, tolerance 1e-10 (40 steps here; the
    oracle needs ~1 s per step at this size)."""
    g, q, bx, by = pr.alfven_wave(512)
    steps = 40
    qg, bxg, byg, dt = run_gpu(g, q, bx, by, steps)
    qo, bxo, byo = run_oracle(g, q, bx, by, steps, dt)
    for c in ALL9:
        assert rel_linf(qg[..., c], qo[..., c]) <= SMOOTH_TOL, c
    assert rel_linf(bxg, bxo) <= SMOOTH_TOL and rel_linf(byg, byo) <= SMOOTH_TOL


def test_orszag_tang_l1():
    """Shock-type problem with degenerate faces (By changes sign, Bz = 0): the reference
    algorithm itself is ill-conditioned there (alpha_s = sqrt(round-off)), so the
    comparison is in L1: tolerance 1e-8 relative after 20 steps at 128^2."""
    g, q, bx, by = pr.orszag_tang(128)
    qg, bxg, byg, dt = run_gpu(g, q, bx, by, 20)
    qo, bxo, byo = run_oracle(g, q, bx, by, 20, dt)
    for c in (0, 1, 2, 4, 5, 7, 8):
        assert rel_l1(interior(qg[..., c]), interior(qo[..., c])) <= 1e-8, c
    # vz, Bz are zero in exact arithmetic; both codes keep them at the WENO-noise level
    assert np.abs(qg[..., 3]).max() < 1e-6 and np.abs(qo[..., 3]).max() < 1e-6


def test_error_is_confined_to_degenerate_faces():
    """The tolerance on Orszag-Tang (1e-7 / L1) is a statement about DEGENERATE faces only -- lines where the
    face-averaged tangential field vanishes (By = 0 at x = k/4, Bx = 0 at y = k/2) and alpha_s is the square root of
    a round-off residual (Weno5_2D.jl:1355-1356).  Shown here, not asserted: one RK4 step at 512^2 against the
    oracle (bit-identical to the reference, tests/test_reference_pin.py); cells farther than 8 cells (measured decay:
    2.7e-10 on the face, 5e-11, 4e-12, 2e-12, 1e-12, 5e-13, 1e-13, 2e-14 at distance 1..7, scripts/
    degenerate_probe.py) from every degenerate face agree to 1e-12 -- two orders below the smooth-problem tolerance
    of north_star -- and the masked band itself stays below 1e-8.  (The reference-run goldens are 16^2 / 24^2 grids,
    where every cell lies inside that band.)"""
    g, q, bx, by = pr.orszag_tang(512)
    qg, bxg, byg, dt = run_gpu(g, q, bx, by, 1)
    qo, bxo, byo = run_oracle(g, q, bx, by, 1, dt)
    dist = degenerate_distance_2d(q, 1e-20)
    band = dist <= 8
    err = np.zeros(q.shape[:2])
    for c in (0, 1, 2, 4, 5, 7, 8):
        err = np.maximum(err, np.abs(qg[..., c] - qo[..., c]) / np.abs(qo[..., c]).max())
    assert (dist == 0).sum() > 0 and 0.05 < band.mean() < 0.5, band.mean()
    print(f"degenerate-face band: {int(band.sum())} of {band.size} cells masked ({100 * band.mean():.1f} %), "
          f"L-inf inside {err[band].max():.2e}, outside {err[~band].max():.2e}")
    assert err[~band].max() <= 1e-12
    assert err[band].max() <= 1e-8


def test_brio_wu_interface_is_the_only_deviation():
    """Brio-Wu, 1024 cells, 8 steps: the initial interface (By changes sign, Bz = 0) is the one degenerate face;
    outside its domain of dependence (13 cells per step) GPU and oracle agree to round-off, inside to 1e-8."""
    g, q = pr.brio_wu(1024)
    dt, steps = 0.4 * g.dx, 8
    with Solver(g) as s:
        s.upload(q)
        for _ in range(steps):
            s.rk4_step(dt)
        qg = s.download()
    qo = q
    for _ in range(steps):
        qo = O.rk4_step_1d(g, qo, dt)
    err = np.zeros(q.shape[0])
    for c in (0, 1, 2, 5, 8):
        err = np.maximum(err, np.abs(qg[:, c] - qo[:, c]) / np.abs(qo[:, c]).max())
    mid, w = q.shape[0] // 2, 13 * steps + 1
    band = np.zeros(q.shape[0], dtype=bool)
    band[mid - w:mid + w] = True
    print(f"Brio-Wu: {int(band.sum())} of {band.size} cells in the interface band, L-inf inside {err[band].max():.2e}, "
          f"outside {err[~band].max():.2e}")
    assert err[~band].max() <= 1e-13
    assert err[band].max() <= 1e-8


def test_shock_cloud_reference_problem_outflow():
    """The reference's own 2-D problem (Weno5_2D.jl:2001-2041) with its outflow boundaries, all
    cells including ghosts and the flux-array rim rules (:1704-1729)."""
    g, q, bx, by = pr.shock_cloud(64, 16)
    qg, bxg, byg, dt = run_gpu(g, q, bx, by, 5)
    qo, bxo, byo = run_oracle(g, q, bx, by, 5, dt)
    for c in (0, 1, 5, 6, 7, 8):
        assert rel_l1(qg[..., c], qo[..., c]) <= 1e-10, c
        assert rel_linf(qg[..., c], qo[..., c]) <= 1e-8, c
    assert rel_linf(bxg, bxo) <= 1e-8 and rel_linf(byg, byo) <= 1e-10


@pytest.mark.parametrize("name,ctor,kw", [
    ("ot24", pr.orszag_tang, dict(n=24)), ("alfven32", pr.alfven_wave, dict(n=32)),
    ("blast24", pr.blast, dict(n=24)), ("shockcloud16x8", pr.shock_cloud, dict(nx=16, ny=8)),
    ("random24x16", pr.random_state, dict(nx=24, ny=16, seed=1234))])
def test_golden_fixtures_2d(name, ctor, kw):
    ref = golden(name)
    g, q, bx, by = ctor(**kw)
    qg, bxg, byg, _ = run_gpu(g, q, bx, by, int(ref["steps"]), dt=float(ref["dt"]))
    tol = 1e-7 if name == "ot24" else 1e-10            # ot24: degenerate faces
    for c in (0, 1, 2, 4, 5, 7, 8):
        assert rel_linf(qg[..., c], ref["q"][..., c]) <= tol, (name, c)
    assert rel_linf(bxg, ref["bxb"]) <= tol and rel_linf(byg, ref["byb"]) <= tol


@pytest.mark.parametrize("name,ctor", [("rj95_2a_64", pr.rj95_2a), ("briowu_64", pr.brio_wu)])
def test_golden_fixtures_1d(name, ctor):
    ref = golden(name)
    g, q = ctor(64)
    with Solver(g) as s:
        s.upload(q)
        for _ in range(int(ref["steps"])):
            s.rk4_step(float(ref["dt"]))
        qg = s.download()
    # both are shock tubes: L1, tolerance 1e-10; Brio-Wu additionally starts with a degenerate face
    # (By changes sign, Bz = 0), so its L-inf bound is looser (DESIGN.md "degenerate faces")
    assert rel_l1(qg, ref["q"]) <= 1e-10
    assert rel_linf(qg, ref["q"]) <= (1e-8 if name == "briowu_64" else 1e-10)


def test_1d_config1_shock_tube_1024():
    """BASELINE config 1: 1-D MHD shock tube, 1024 cells, run to t = 0.2 with the device-side
    driver loop; compared with the oracle in L1 (shock problem; tolerance 1e-9)."""
    g, q = pr.rj95_2a(1024)
    with Solver(g) as s:
        s.upload(q)
        vx = s.timestep()[1]
        assert abs(vx - O.vmax_1d(g, q)[0]) <= 1e-14 * vx
        t, n = s.advance(10 ** 6, tmax=0.2)
        qg = s.download()
    assert t == pytest.approx(0.2, abs=1e-15) and n > 100
    qo, to = q, 0.0
    while to < 0.2:
        dt = min(O.vmax_1d(g, qo)[1], 0.2 - to)
        qo = O.rk4_step_1d(g, qo, dt)
        to += dt
    for c in ALL9:
        assert rel_l1(qg[:, c], qo[:, c]) <= 1e-9, c
    rho = qg[:, 0]
    assert 0.9 < rho.min() and rho.max() < 1.75        # Weno5_1D.jl:79


# ------------------------------------------------------------------ API behaviour
@pytest.mark.parametrize("case", ["rj95_1024", "briowu_200", "periodic_300_ragged"])
def test_cluster_1d_kernel_matches_multi_kernel_path(case, monkeypatch):
    """The 1-D solver as one cluster-resident kernel (weno5_line1d.cu: 8 CTAs, DSMEM halos, step loop and CFL
    reduction inside the kernel) against the same run through the multi-kernel path (WENO5_NO_LINE1D): same device
    functions and update formulas, so round-off agreement (separate instantiations: ptxas may contract a few
    multiply-adds differently), identical step count and time."""
    if case == "rj95_1024":
        g, q = pr.rj95_2a(1024)
    elif case == "briowu_200":
        g, q = pr.brio_wu(200)
    else:
        g = pr.Grid(nx=300, ny=1, xsize=1.0, bc_x=pr.PERIODIC)
        x = g.xc()
        q = np.zeros((g.Nx, 9), order="F")
        for i in range(g.Nx):
            s_, c_ = np.sin(2 * np.pi * x[i]), np.cos(2 * np.pi * x[i])
            q[i] = pr.state2conserved(1.0 + 0.2 * s_, 0.5 + 0.1 * c_, 0.1 * s_, -0.1 * c_, 0.75, 0.3 + 0.1 * c_, 0.2 * s_,
                                      1.0 + 0.1 * s_, g.gam)
    res = {}
    for mode in ("cluster", "multi"):
        if mode == "multi":
            monkeypatch.setenv("WENO5_NO_LINE1D", "1")
        else:
            monkeypatch.delenv("WENO5_NO_LINE1D", raising=False)
        with Solver(g) as s:
            s.upload(q)
            l0 = s.launch_count()
            t, n = s.advance(7)
            adv_launches = s.launch_count() - l0
            dt = s.timestep()[0]
            s.rk4_step(dt)
            res[mode] = (s.download(), t, n, adv_launches)
    (qa, ta, na, la), (qb, tb, nb, lb) = res["cluster"], res["multi"]
    assert na == nb == 7 and abs(ta - tb) <= 1e-14 * tb
    assert la == 1 and lb > 50                     # one launch for the whole loop vs ~15 per step
    for c in range(9):
        scale = max(np.abs(qb[:, c]).max(), 1e-3)
        assert np.abs(qa[:, c] - qb[:, c]).max() <= 1e-12 * scale, (case, c)


def test_one_shot_call_equals_handle_path():
    g, q, bx, by = pr.random_state(40, 24, seed=9)
    dt = 0.4 * O.timestep_2d(g, q)[0]
    a = S.classical_RK4_step(g, q, bx, by, dt)
    b = run_gpu(g, q, bx, by, 1, dt=dt)[:3]
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    g1, q1 = pr.rj95_2a(64)                  # (Brio-Wu starts with a degenerate face)
    q4 = S.classical_RK4_step(g1, q1, dt=1e-3)
    assert rel_linf(q4, O.rk4_step_1d(g1, q1, 1e-3)) <= 1e-12


def test_reference_named_helpers():
    g, q, bx, by = pr.shock_cloud(32, 16)
    dt, vx, vy = S.timestep(g, q)
    dto, vxo, vyo = O.timestep_2d(g, q)
    assert (abs(dt - dto) <= 1e-14 * dto) and abs(vx - vxo) <= 1e-14 * vxo and abs(vy - vyo) <= 1e-14 * vyo
    q2, bx2, by2 = S.interpolateB_center2face(g, q)
    bxo, byo = O.center2face(g, q)
    qo = q.copy(order="F")
    O.boundaries_2d(g, qo, bxo, byo)
    # the device contracts the 4th-order stencil into FMAs: last-bit differences are allowed
    assert rel_linf(bx2, bxo) <= 4e-16 and rel_linf(by2, byo) <= 4e-16 and np.array_equal(q2, qo)
    d = S.compute_divB(g, bx, by)
    assert np.abs(d - O.divb(g, bx, by)).max() <= 1e-12 * max(1.0, np.abs(d).max())


def test_es_roundtrip_flag_is_roundoff_only():
    g, q, bx, by = pr.alfven_wave(48)
    a = run_gpu(g, q, bx, by, 5, es_roundtrip=True)
    b = run_gpu(g, q, bx, by, 5, es_roundtrip=False)
    for c in ALL9:
        assert rel_linf(a[0][..., c], b[0][..., c]) <= 1e-13


def test_uniform_state_preserved_exactly_on_gpu():
    g = pr.Grid(nx=64, ny=40, bc_x=pr.PERIODIC, bc_y=pr.PERIODIC)
    s = pr.state2conserved(1.3, 0.4, -0.2, 0.1, 0.5, -0.3, 0.2, 0.7, g.gam)
    q = np.zeros((g.Nx, g.Ny, 9), order="F")
    for c in range(9):
        q[..., c] = s[c]
    bx = np.full((g.Nx, g.Ny), 0.5, order="F")
    by = np.full((g.Nx, g.Ny), -0.3, order="F")
    qg, bxg, byg, _ = run_gpu(g, q, bx, by, 3, dt=0.01, es_roundtrip=False)
    # 0.5 is exact in binary, -0.3 is not: (1/3)(-b + b + 2b + b) may differ from b in the last bit
    assert rel_linf(qg[..., :7], q[..., :7]) <= 2e-16 and np.array_equal(bxg, bx) and rel_linf(byg, by) <= 2e-16


def test_xy_symmetry_on_gpu():
    g, q, bx, by = pr.random_state(48, 40, seed=5)
    gt = pr.Grid(nx=g.ny, ny=g.nx, xsize=g.ysize, ysize=g.xsize, gam=g.gam, bc_x=g.bc_y, bc_y=g.bc_x)
    perm = [0, 2, 1, 3, 5, 4, 6, 7, 8]
    qt = np.asfortranarray(np.transpose(q[..., perm], (1, 0, 2)).copy())
    qt[..., 3] *= -1
    qt[..., 6] *= -1
    dt = 0.5 * O.timestep_2d(g, q)[0]
    a = run_gpu(g, q, bx, by, 2, dt=dt)
    b = run_gpu(gt, qt, np.asfortranarray(by.T.copy()), np.asfortranarray(bx.T.copy()), 2, dt=dt)
    back = np.transpose(b[0], (1, 0, 2))[..., perm].copy()
    back[..., 3] *= -1
    back[..., 6] *= -1
    assert rel_linf(interior(back), interior(a[0])) <= 1e-12
    assert rel_linf(interior(b[2].T), interior(a[1])) <= 1e-12


def test_advance_lands_on_tmax_and_is_idempotent_after():
    g, q, bx, by = pr.orszag_tang(64)
    with Solver(g) as s:
        s.upload(q, bx, by)
        t, n = s.advance(10 ** 6, tmax=0.02)
        assert t == pytest.approx(0.02, abs=1e-16) and 3 < n < 1000
        q1 = s.download()[0]
        t2, n2 = s.advance(5, tmax=0.02, t=t)           # already there: nothing may change
        assert n2 == 0 and t2 == t and np.array_equal(s.download()[0], q1)


def test_error_paths():
    g, q, bx, by = pr.orszag_tang(32)
    with Solver(g) as s:
        with pytest.raises(L.Weno5Error) as e:
            s.rk4_step(0.1)
        assert e.value.code == L.ERR_STATE
        s.upload(q, bx, by)
        with pytest.raises(L.Weno5Error) as e:
            s.rk4_step(float("nan"))
        assert e.value.code == L.ERR_ARG
        bad = q.copy(order="F")
        bad[10, 10, 0] = np.nan
        s.upload(bad, bx, by)
        with pytest.raises(L.Weno5Error) as e:
            s.timestep()
        assert e.value.code == L.ERR_NUMERIC


# ------------------------------------------------------------------ BASELINE sizes: invariants
@pytest.mark.parametrize("n", [2048, 4096])
def test_orszag_tang_full_size_invariants(n):
    """Config 3 (2048^2) and the roofline size (4096^2): conservation of mass, momentum, energy
    and magnetic flux to round-off, div B at round-off, state stays finite and physical."""
    g, q, bx, by = pr.orszag_tang(n)
    with Solver(g, es_roundtrip=False) as s:
        s.upload(q, bx, by)
        t0 = s.totals()
        t, nst = s.advance(5)
        t1 = s.totals()
        divb = s.divb()
        assert nst == 5 and t > 0
        qs = s.download(faces=False)
    cells = float(n) * n
    for c in (0, 1, 2, 3, 4, 5, 6, 8, 9, 10):       # S (7) is not conserved
        scale = max(abs(t0[c]), cells * 1e-3)
        assert abs(t1[c] - t0[c]) <= 1e-12 * scale, (c, t0[c], t1[c])
    assert np.abs(divb).max() * g.dx < 1e-12        # |div B| dx relative to |B| ~ 0.3
    assert np.isfinite(qs).all() and qs[..., 0].min() > 0


@pytest.mark.parametrize("n,steps", [(2048, 3), (4096, 1)])
def test_orszag_tang_full_size_parity_vs_oracle(n, steps):
    """BASELINE config 3 (2048^2, 3 steps) and the bench / roofline size (4096^2, 1 step) against the CPU oracle on
    the same input, every step with its own CFL time step on both sides: relative L1 <= 1e-8 on every component
    (Orszag-Tang has degenerate faces, so L1 as north_star prescribes; the L-inf figure is printed)."""
    g, q, bx, by = pr.orszag_tang(n)
    with Solver(g) as s:
        s.upload(q, bx, by)
        t, nst = s.advance(steps)
        qg, bxg, byg = s.download()
    qo, bxo, byo, to = O.advance_2d(g, q, bx, by, steps, fast=True)
    assert nst == steps and abs(t - to) <= 1e-12 * to, (t, to)
    worst = 0.0
    for c in ALL9:
        a, b = interior(qg[..., c]), interior(qo[..., c])
        scale = max(np.abs(b).mean(), 1e-3)           # Mz, Bz vanish identically in this problem
        e1 = float(np.abs(a - b).mean() / scale)
        worst = max(worst, float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-3)))
        assert e1 <= 1e-8, (c, e1)
    assert rel_l1(interior(bxg), interior(bxo)) <= 1e-8 and rel_l1(interior(byg), interior(byo)) <= 1e-8
    print(f"orszag_tang {n}^2 x {steps}: worst relative L-inf vs oracle {worst:.3e}")


@pytest.mark.parametrize("case", ["periodic", "outflow"])
def test_fused_cfl_reduction_equals_stepwise(case):
    """weno5_advance() takes the wave-speed maxima of step k+1 from stage 4 of step k (ct_centers_entropy_vmax:
    no separate pass over the state); the result must equal timestep() -> classical_RK4_step() called step by
    step, bit for bit -- widths that are no multiple of the kernel's 63-column blocks, both boundary types."""
    if case == "periodic":
        g, q, bx, by = pr.orszag_tang(100, ny=76, ysize=0.76)
    else:
        g, q, bx, by = pr.shock_cloud(130, 50, xsize=2.6, ysize=1.0)
    with Solver(g) as s:
        s.upload(q, bx, by)
        t_adv, n = s.advance(4)
        qa, bxa, bya = s.download()
    assert n == 4
    with Solver(g) as s:
        s.upload(q, bx, by)
        t = 0.0
        for _ in range(4):
            dt = s.timestep()[0]
            s.rk4_step(dt)
            t += dt
        qb, bxb_, byb_ = s.download()
    assert t_adv == t
    assert np.array_equal(qa, qb) and np.array_equal(bxa, bxb_) and np.array_equal(bya, byb_)


def test_tma_and_ldgsts_staging_agree_to_roundoff(tmp_path):
    """Four instantiations of the flux sweep: the warp-specialised kernel (WENO5_FLUX_VARIANT=5: face
    warps + helper warps, TMA staging), the TMEM-stash kernel (6: 168 registers, 3 CTAs per SM), the single-role
    kernel with TMA staging (4) and with cp.async (2).  The arithmetic is the same source, but they are separate
    template instantiations and ptxas contracts a few multiply-adds of the difference-first projection
    differently (measured 2.7e-15 after 3 steps; the 12-projection form, -DW5_FACE_PLAIN, is bit-equal):
    results agree to round-off.  Runs on one GPU vs slabs and the pipelined host step use ONE
    instantiation and stay bit-identical (test_gpu_multi.py, test_pipelined_host_step_is_bit_identical)."""
    import os
    import subprocess
    import sys
    from util import ROOT
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from julia_b200 import problems as pr\n"
        "from julia_b200.solver import Solver\n"
        "g, q, bx, by = pr.random_state(136, 72, seed=21)\n"
        "s = Solver(g); s.upload(q, bx, by); s.advance(3); out = s.download()\n"
        "np.savez(sys.argv[1], q=out[0], bx=out[1], by=out[2])\n" % ROOT)
    res = {}
    for variant in ("5", "6", "4", "2"):
        out = str(tmp_path / f"v{variant}.npz")
        env = dict(os.environ, WENO5_FLUX_VARIANT=variant)
        subprocess.check_call([sys.executable, "-c", code, out], env=env)
        res[variant] = np.load(out)
    for k in ("q", "bx", "by"):
        for other in ("6", "4", "2"):
            assert np.abs(res["5"][k] - res[other][k]).max() <= 1e-13 * np.abs(res["5"][k]).max(), (k, other)
    assert np.isfinite(res["5"]["q"]).all()


@pytest.mark.parametrize("rows,direct", [(64, False), (40, False), (100, False), (64, True)])
@pytest.mark.parametrize("case", ["periodic", "outflow_ragged"])
def test_pipelined_host_step_is_bit_identical(case, rows, direct, monkeypatch):
    """weno5_pipe_step streams y-slabs (+16 overlap rows = domain of dependence of one RK4 step)
    through the GPU -- inputs through an image of the grid on the device that the slabs gather their windows from,
    or (direct) every slab uploading its own window -- for slab sizes that do and do not divide the grid;
    every row of the result, the ghost rows and the returned next time step must
    equal the monolithic upload -> step -> download path bit for bit."""
    from julia_b200.solver import Pipe
    if case == "periodic":
        g, q, bx, by = pr.orszag_tang(96, ny=256, ysize=256 / 96)
    else:
        g, q, bx, by = pr.shock_cloud(96, 200, xsize=2.0, ysize=4.0)
    with Solver(g) as s:
        s.upload(q, bx, by)
        dt = s.timestep()[0]
        s.rk4_step(dt)
        dt_next = s.timestep()[0]
        q_ref, bx_ref, by_ref = s.download()
    if direct:
        monkeypatch.setenv("WENO5_PIPE_DIRECT", "1")
    else:
        monkeypatch.delenv("WENO5_PIPE_DIRECT", raising=False)
    with Pipe(g, rows_per_slab=rows) as p:
        q4, bx4, by4, dtn = p.step(q, bx, by, dt)
        assert np.array_equal(q4, q_ref) and np.array_equal(bx4, bx_ref) and np.array_equal(by4, by_ref)
        assert dtn == dt_next
        # second step from the pipelined result, to exercise buffer reuse
        q5, bx5, by5, _ = p.step(q4, bx4, by4, dtn)
    with Solver(g) as s:
        s.upload(q_ref, bx_ref, by_ref)
        s.rk4_step(dt_next)
        q_ref2 = s.download()[0]
    assert np.array_equal(q5, q_ref2)


def test_pipelined_step_argument_checks():
    """Partially overlapping input / output arrays are rejected (not only identical pointers); pageable arrays are
    accepted, flagged as not overlapped, and give the same result as page-locked ones."""
    from julia_b200.solver import Pipe
    g, q, bx, by = pr.orszag_tang(96, ny=256, ysize=256 / 96)
    with Pipe(g, rows_per_slab=64) as p:
        big = np.zeros(q.size + 64)
        qin = big[:q.size].reshape(q.shape, order="F")
        qin[...] = q
        qout = big[64:64 + q.size].reshape(q.shape, order="F")          # shifted view of the same buffer
        with pytest.raises(L.Weno5Error) as e:
            p.step(qin, bx, by, 1e-4, out=(qout, np.zeros_like(bx, order="F"), np.zeros_like(by, order="F")))
        assert e.value.code == L.ERR_ARG and "overlap" in str(e.value)
        q4, bx4, by4, dtn = p.step(q, bx, by, 1e-4)
        assert not p.overlapped()
        hq, hbx, hby = p.pinned(q.shape), p.pinned(bx.shape), p.pinned(by.shape)
        oq, obx, oby = p.pinned(q.shape), p.pinned(bx.shape), p.pinned(by.shape)
        hq[...], hbx[...], hby[...] = q, bx, by
        r = p.step(hq, hbx, hby, 1e-4, out=(oq, obx, oby))
        assert p.overlapped()
        assert np.array_equal(r[0], q4) and np.array_equal(r[1], bx4) and np.array_equal(r[2], by4) and r[3] == dtn
