"""Runs the REFERENCE ITSELF (Weno5_1D.jl / Weno5_2D.jl, translated mechanically by oracle/jl2py.py and
executed on oracle/jlrt.py) and writes its inputs/outputs as golden vectors tests/golden/ref_*.npz.

Only runnable in the build container (needs /root/reference); the fixtures and this script are committed,
the translated source goes to a scratch directory outside the repository and is never committed.  The translated functions
are called unmodified; what the harness does around them is stated per case:

  ref1d_rj95_2a      Weno5_1D: reference IC (N = 256+6), the driver's loop `vxmax = compute_global_vmax(q);
                     dt = courFac*dx/vxmax; q = classical_RK4_step(q, dt)` for 6 steps.
  ref1d_briowu       same functions with N = 64+6, gamma = 2 (module constants patched) on the Brio-Wu IC.
  ref2d_shockcloud   Weno5_2D exactly as shipped (16x4 cells, 4:1 box, shock-cloud IC, outflow): IC ->
                     interpolateB_center2face -> boundaries! -> 2 x (timestep, classical_RK4_step), compute_divB.
                     (The driver's debugging lines `dt = 1e-8` / `@assert false`, Weno5_2D.jl:54,61, are not used.)
  ref2d_shockcloud32 same with Nx,Ny = 32+6, 8+6 (module constants patched), 1 step.
  ref2d_functions    every hot-path function called once on a seeded random state (20x14 incl. ghosts): primitives,
                     eigenvalues, fluxes, L/R eigenvectors, WENO face fluxes, x and y flux differences (E and S
                     variants), rotate_state, corner_fluxes, face<->centre interpolation, E2S/S2E, timestep,
                     compute_divB, boundaries!.
  ref2d_periodic_ot  the RK4 step on a 16x16 Orszag-Tang state with `boundaries!` REPLACED by a periodic wrap
                     (the reference has outflow only; this pins the interior numerics the bench workload uses).

  ref2d_dual_*       "next" row f3: the same RK4 step with the reference's own commented-out switching criterion
                     restored (the only source edit, see DUAL_PATCH), so that the S branch is live: periodic blast
                     wave 16x16 (2 steps) and the shock-cloud problem (1 step); `nbad` = cells switched to E per stage.

  ref1d_switch       Weno5_1D as shipped on RJ95-2a (N = 64+6, 4 steps): q per step AND `idx`, the list of cells whose
                     entropy-branch candidate differs from the energy branch by |S(E) - S| >= 1e-5 in the last stage
                     (ES_Switch, Weno5_1D.jl:238-258 -- it evolves E only but still runs the S sweeps for this list).
  ref1d_dual_live    the same with the three "E only" override lines of ES_Switch (:252-254) removed (DUAL_PATCH_1D),
                     i.e. the switch as the author wrote it above them: S state by default, E state where flagged.

  ref_arr2img        "next" row f4: Arr2Img (JColorMaps.jl:92-146, the only function translated from that file) on
                     fields of the shock-cloud result with an identity colour table.

`oob` in each file counts the @inbounds reads that left their column: "wrapped" ones read
the neighbouring column like the Julia binary would, "outside" ones (past the allocation) read NaN here.
"""
import importlib.util
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from julia_b200 import problems as pr  # noqa: E402
from oracle import jl2py, jlrt  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
# translated source never enters the repository tree: it lives in a scratch directory for the duration of the run
GEN = os.path.join(tempfile.gettempdir(), "weno5_jl2py_ref")


# The one source edit any case makes (case_dual only): the reference's own commented-out switching criterion
# (Weno5_2D.jl:1740-1741) is restored, i.e. `#dS = abs.(...)` is uncommented and the override `dS = 1` dropped.
DUAL_PATCH = (("#dS = abs.(q_SE[:,:,8] - q_S[:,:,8])", "dS = abs.(q_SE[:,:,8] - q_S[:,:,8])"), ("\ndS = 1\n", "\n"))


DUAL_PATCH_1D = (("    q[:,1:7] = q_E[:,1:7] # E only\n    q[:,8] = E2S(q_E)\n    q[:,9] = q_E[:,8]\n", ""),)


def load(name, patch=(), tag="", only=None):
    os.makedirs(GEN, exist_ok=True)
    with open(os.path.join(REF, name + ".jl")) as f:
        src = f.read()
    for old, new in patch:
        assert src.count(old) == 1, old
        src = src.replace(old, new)
    py, failed = jl2py.translate(src, only=only)
    assert not failed, failed
    path = os.path.join(GEN, name.lower() + tag + "_ref.py")
    with open(path, "w") as f:
        f.write(py)
    spec = importlib.util.spec_from_file_location(name.lower() + tag + "_ref", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def reset_oob():
    jlrt.OOB["wrapped"] = jlrt.OOB["outside"] = 0


def oob():
    return np.array([jlrt.OOB["wrapped"], jlrt.OOB["outside"]])


def configure_2d(m, nx, ny, xsize, ysize):
    m.Nx, m.Ny = nx + 2 * m.NBnd, ny + 2 * m.NBnd
    m.iMin, m.iMax, m.jMin, m.jMax = 1 + m.NBnd, m.Nx - m.NBnd, 1 + m.NBnd, m.Ny - m.NBnd
    m.xsize, m.ysize = xsize, ysize
    m.dx, m.dy = xsize / (m.Nx - 2 * m.NBnd), ysize / (m.Ny - 2 * m.NBnd)
    m.flag = False


def configure_1d(m, n, gamma):
    m.N = n + 2 * m.NBnd
    m.iMin, m.iMax = 1 + m.NBnd, m.N - m.NBnd
    m.dx = m.xsize / (m.N - 2 * m.NBnd)
    m.gamma = gamma
    m.gam0, m.gam1, m.gam2, m.gamS = 1 - gamma, (gamma - 1) / 2, (gamma - 2) / (gamma - 1), (gamma - 1) / gamma


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print(f"wrote {name}.npz", flush=True)


# ------------------------------------------------------------------------------------------ 1-D
def run_1d(m, q, steps):
    qs, dts = [np.array(q)], []
    for _ in range(steps):
        vxmax = m.compute_global_vmax(q)
        dt = m.courFac * m.dx / vxmax
        q, idx, omega = m.classical_RK4_step(q, dt)
        qs.append(np.array(q))
        dts.append(float(dt))
    return np.stack(qs), np.array(dts)


def case_1d(m1):
    reset_oob()
    configure_1d(m1, 256, 5.0 / 3.0)
    q0 = m1.initial_conditions()
    qs, dts = run_1d(m1, q0, 6)
    save("ref1d_rj95_2a", q=qs, dt=dts, nx=256, gamma=5.0 / 3.0, oob=oob())
    reset_oob()
    configure_1d(m1, 64, 2.0)
    g, q0 = pr.brio_wu(64)
    qs, dts = run_1d(m1, np.asfortranarray(q0.reshape(70, 9)), 8)
    save("ref1d_briowu", q=qs, dt=dts, nx=64, gamma=2.0, oob=oob())
    configure_1d(m1, 256, 5.0 / 3.0)


def case_1d_switch():
    """The 1-D ES_Switch: `idx` of the shipped code, and the live switch (override lines removed)."""
    for name, patch, tag in (("ref1d_switch", (), ""), ("ref1d_dual_live", DUAL_PATCH_1D, "_live")):
        m1 = load("Weno5_1D", patch=patch, tag=tag)
        reset_oob()
        configure_1d(m1, 64, 5.0 / 3.0)
        q = m1.initial_conditions()
        qs, dts, idxs = [np.array(q)], [], []
        for _ in range(4):
            vxmax = m1.compute_global_vmax(q)
            dt = m1.courFac * m1.dx / vxmax
            q, idx, omega = m1.classical_RK4_step(q, dt)
            qs.append(np.array(q))
            dts.append(float(dt))
            flag = np.zeros(q.shape[0])
            flag[np.asarray(idx, dtype=int).ravel() - 1] = 1.0          # Julia indices are 1-based
            idxs.append(flag)
        save(name, q=np.stack(qs), dt=np.array(dts), flag=np.stack(idxs), nx=64, gamma=5.0 / 3.0, oob=oob())


# ------------------------------------------------------------------------------------------ 2-D
def run_2d(m, steps):
    q = m.initial_conditions()
    bxb, byb = m.interpolateB_center2face(q)
    m.boundaries_b(q, bxb, byb)
    out = dict(q0=np.array(q), bxb0=np.array(bxb), byb0=np.array(byb))
    dts, vm = [], []
    for s in range(1, steps + 1):
        dt, vxmax, vymax = m.timestep(q)
        q, bxb, byb = m.classical_RK4_step(q, bxb, byb, dt)
        dts.append(float(dt))
        vm.append((float(vxmax), float(vymax)))
        out[f"q{s}"], out[f"bxb{s}"], out[f"byb{s}"] = np.array(q), np.array(bxb), np.array(byb)
        out[f"divb{s}"] = m.compute_divB(bxb, byb)
    out["dt"], out["vmax"] = np.array(dts), np.array(vm)
    return out


def case_shockcloud(m2):
    reset_oob()
    configure_2d(m2, 16, 4, 4, 1)
    save("ref2d_shockcloud", nx=16, ny=4, xsize=4.0, ysize=1.0, **run_2d(m2, 2), oob=oob())
    reset_oob()
    configure_2d(m2, 32, 8, 4, 1)
    save("ref2d_shockcloud32", nx=32, ny=8, xsize=4.0, ysize=1.0, **run_2d(m2, 1), oob=oob())
    configure_2d(m2, 16, 4, 4, 1)


def case_functions(m2):
    reset_oob()
    nx, ny = 14, 8
    configure_2d(m2, nx, ny, 1.0, 0.75)
    g, q, bxb, byb = pr.random_state(nx, ny, seed=4321, bc=pr.OUTFLOW, smooth=False)
    q, bxb, byb = (np.asfortranarray(a) for a in (q, bxb, byb))
    out = dict(nx=nx, ny=ny, xsize=1.0, ysize=0.75, q=q, bxb=bxb, byb=byb)
    idxE, idxS = [1, 2, 3, 4, 5, 6, 7, 9], [1, 2, 3, 4, 5, 6, 7, 8]
    qE = jlrt.jget(q, (jlrt.COLON, jlrt.COLON, idxE))
    qS = jlrt.jget(q, (jlrt.COLON, jlrt.COLON, idxS))
    u = m2.compute_primitive_variables_E(qE)
    a = m2.compute_eigenvalues(u)
    F = m2.compute_fluxes_E(qE, u)
    L, R = m2.compute_eigenvectors_E_vec(qE, u, F, m2.gam)
    dF = m2.weno5_interpolation(qE, a, F, L, R)
    out.update(uE=u, aE=a, FE=F, LE=L, RE=R, dFE=dF)
    dq, bsy, bsz = m2.weno5_flux_difference_E(qE)
    out.update(dqx_E=dq, fsy_E=bsy, bsz_x_E=bsz)
    qrot = m2.rotate_state(qE, "fwd")
    dqr, tmp, gsxr = m2.weno5_flux_difference_E(qrot)
    out.update(qrot_E=qrot, dqy_E=m2.rotate_state(dqr, "bwd"), gsx_E=jlrt.transpose(gsxr), bsy_y_E=jlrt.transpose(tmp))
    # S branch (dead in the shipped RK4 step; kept for the "next" row f3)
    uS = m2.compute_primitive_variables_S(qS)
    FS = m2.compute_fluxes_S(qS, uS)
    LS, RS = m2.compute_eigenvectors_S_vec(qS, uS, FS, m2.gam)
    out.update(uS=uS, FS=FS, LS=LS, RS=RS, aS=m2.compute_eigenvalues(uS))
    dq, bsy, bsz = m2.weno5_flux_difference_S(qS)
    out.update(dqx_S=dq, fsy_S=bsy, bsz_x_S=bsz)
    dqr, tmp, gsxr = m2.weno5_flux_difference_S(m2.rotate_state(qS, "fwd"))
    out.update(dqy_S=m2.rotate_state(dqr, "bwd"), gsx_S=jlrt.transpose(gsxr))
    # constrained transport pieces, switches, CFL, divB, boundaries
    fsy, gsx = np.array(out["fsy_E"]), np.array(out["gsx_E"])
    Ox, Oxi, Oxj = m2.corner_fluxes(fsy, gsx)
    out.update(Ox=Ox, Oxi=Oxi, Oxj=Oxj)
    out["Bxy"] = m2.interpolateB_face2center(bxb, byb)
    c2f = m2.interpolateB_center2face(q)
    out.update(c2f_bxb=c2f[0], c2f_byb=c2f[1])
    out["E2S"] = m2.E2S(qE)
    out["S2E"] = m2.S2E(qS)
    out["timestep"] = np.array([float(x) for x in m2.timestep(q)])
    out["divb"] = m2.compute_divB(bxb, byb)
    qb, fb, gb = np.array(q, order="F"), np.array(fsy, order="F"), np.array(gsx, order="F")
    m2.boundaries_b(qb, fb, gb)
    out.update(bnd_q=qb, bnd_fsy=fb, bnd_gsx=gb)
    out["state2conserved"] = m2.state2conserved(1.3, 0.2, -0.4, 0.6, 0.7, -0.1, 0.25, 0.9)
    save("ref2d_functions", **out, oob=oob())
    configure_2d(m2, 16, 4, 4, 1)


def case_periodic(m2):
    reset_oob()
    n = 16
    configure_2d(m2, n, n, 1.0, 1.0)
    g, q, bxb, byb = pr.orszag_tang(n)
    q, bxb, byb = (np.asfortranarray(a) for a in (q, bxb, byb))
    nb = m2.NBnd

    def wrap(a):
        Nx, Ny = a.shape[0], a.shape[1]
        a[0:nb] = a[Nx - 2 * nb:Nx - nb]
        a[Nx - nb:Nx] = a[nb:2 * nb]
        a[:, 0:nb] = a[:, Ny - 2 * nb:Ny - nb]
        a[:, Ny - nb:Ny] = a[:, nb:2 * nb]

    def periodic_boundaries(q_, fsy, gsx):
        for a in (q_, fsy, gsx):
            wrap(a)

    saved = m2.boundaries_b
    m2.boundaries_b = periodic_boundaries
    try:
        periodic_boundaries(q, bxb, byb)
        dt, vxmax, vymax = m2.timestep(q)
        q1, bx1, by1 = m2.classical_RK4_step(q, bxb, byb, dt)
    finally:
        m2.boundaries_b = saved
    save("ref2d_periodic_ot", nx=n, ny=n, xsize=1.0, ysize=1.0, q0=q, bxb0=bxb, byb0=byb, dt=float(dt),
         vmax=np.array([float(vxmax), float(vymax)]), q1=q1, bxb1=bx1, byb1=by1, oob=oob())
    configure_2d(m2, 16, 4, 4, 1)


def case_dual():
    """f3: the RK4 step with the S branch live.  Blast wave 16x16 with periodic ghosts (uniform far field keeps
    the S state, the blast front switches to E) and the reference's shock-cloud problem with its own boundaries!."""
    m2 = load("Weno5_2D", DUAL_PATCH, "_dual")
    nb = m2.NBnd
    calls = []
    orig_switch = m2.ES_Switch

    def counting_switch(*a):
        r = orig_switch(*a)
        calls.append(len(r[3]))
        return r

    m2.ES_Switch = counting_switch
    # periodic blast
    reset_oob()
    n = 16
    configure_2d(m2, n, n, 1.0, 1.0)
    g, q, bxb, byb = pr.blast(n)
    q, bxb, byb = (np.asfortranarray(a) for a in (q, bxb, byb))

    def wrap(a):
        Nx, Ny = a.shape[0], a.shape[1]
        a[0:nb] = a[Nx - 2 * nb:Nx - nb]
        a[Nx - nb:Nx] = a[nb:2 * nb]
        a[:, 0:nb] = a[:, Ny - 2 * nb:Ny - nb]
        a[:, Ny - nb:Ny] = a[:, nb:2 * nb]

    def periodic_boundaries(q_, fsy, gsx):
        for a in (q_, fsy, gsx):
            wrap(a)

    saved = m2.boundaries_b
    m2.boundaries_b = periodic_boundaries
    try:
        periodic_boundaries(q, bxb, byb)
        out = dict(nx=n, ny=n, xsize=1.0, ysize=1.0, q0=q, bxb0=bxb, byb0=byb)
        dts = []
        for s in (1, 2):
            dt, vxmax, vymax = m2.timestep(q)
            q, bxb, byb = m2.classical_RK4_step(q, bxb, byb, dt)
            dts.append(float(dt))
            out[f"q{s}"], out[f"bxb{s}"], out[f"byb{s}"] = np.array(q), np.array(bxb), np.array(byb)
    finally:
        m2.boundaries_b = saved
    save("ref2d_dual_blast", **out, dt=np.array(dts), nbad=np.array(calls), oob=oob())
    # the reference's own problem
    reset_oob()
    del calls[:]
    configure_2d(m2, 16, 4, 4, 1)
    out = run_2d(m2, 1)
    save("ref2d_dual_shockcloud", nx=16, ny=4, xsize=4.0, ysize=1.0, **out, nbad=np.array(calls), oob=oob())


def case_arr2img():
    """f4: the snapshot hook.  Arr2Img (JColorMaps.jl:92-146) on fields of the shock-cloud state after one step,
    with an identity colour table of n entries (the returned "image" is then the colour index itself): fixed
    range, automatic range, log scaling, a field with NaNs."""
    mc = load("JColorMaps", only={"Arr2Img"})
    d = np.load(os.path.join(OUT, "ref2d_shockcloud32.npz"))
    out = {}
    rho = np.asfortranarray(d["q1"][:, :, 0])
    by = np.asfortranarray(d["q1"][:, :, 5])
    hole = rho.copy(order="F")
    hole[5, 5] = np.nan
    hole[20, 7] = np.nan
    cases = {"fixed": (rho, 256, dict(range_=jlrt.jvect([1.0, 3.0]))),
             "auto": (by, 256, {}),
             "log": (rho, 64, dict(log=True)),
             "nan": (hole, 256, {}),
             "logneg": (np.asfortranarray(d["q1"][:, :, 1]), 200, dict(log=True))}
    for name, (arr, n, kw) in cases.items():
        cmap = np.arange(1.0, n + 1.0)
        out["in_" + name] = arr
        out["out_" + name] = mc.Arr2Img(np.array(arr, order="F"), cmap, **kw)
        out["n_" + name] = n
    save("ref_arr2img", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["functions", "1d", "1d_switch", "shockcloud", "periodic", "dual", "arr2img"]
    os.makedirs(OUT, exist_ok=True)
    t0 = time.time()
    if "1d" in which:
        case_1d(load("Weno5_1D"))
    if "1d_switch" in which:
        case_1d_switch()
    if set(which) & {"functions", "shockcloud", "periodic"}:
        m2 = load("Weno5_2D")
        if "functions" in which:
            case_functions(m2)
        if "shockcloud" in which:
            case_shockcloud(m2)
        if "periodic" in which:
            case_periodic(m2)
    if "dual" in which:
        case_dual()
    if "arr2img" in which:
        case_arr2img()
    print(f"done in {time.time() - t0:.0f} s")
