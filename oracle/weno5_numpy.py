"""
TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference WENO5 MHD numerics.

This file is one of two independent CPU restatements (the other is the C oracle in
``oracle/weno5_oracle.c``) of the algorithm of jdonnert/julia's
``Weno5_2D.jl`` / ``Weno5_1D.jl`` (the shipped E-branch step, the entropy-variable S sweeps, the dual-energy
step with a live ES_Switch, and Arr2Img of JColorMaps.jl).  It is array-at-a-time like the reference
(whole-grid temporaries, transposed copy for the y sweep) and is only ever imported
by ``tests/`` -- never by the product path.

PARITY PINNED against outputs of the reference itself: the Julia 0.6 source is translated
mechanically (oracle/jl2py.py) and its own functions are executed unmodified by
scripts/make_reference_golden.py; tests/test_reference_pin.py checks this file against the
committed vectors tests/golden/ref_*.npz (bit-identical functions, sweeps and full RK4
step; the only deviation is quirk Q3 below).  Further pins: algebraic identities
(L.R = I, R diag(a) L = dF/dq), invariants, and agreement with the independently written
C oracle (tests/test_oracle.py).

Index convention: arrays are numpy (Nx, Ny, C) with 0-based indices; a reference
1-based index ``i`` is ``i-1`` here.  NBnd = 3 ghost cells on every side.

Decisions for the reference quirks:
  Q1  E branch only (the S branch is dead code in the shipped configuration).
  Q2  2D: E is round-tripped through S each stage (Weno5_2D.jl:1738,1765); 1D keeps E.
  Q3  the face at i = Nq-2 reads one cell past the array in the reference
      (Weno5_2D.jl:1812-1824); here that cell is the natural continuation of the
      boundary condition (zero-gradient clamp for outflow, wrap for periodic).
  Q13 timestep uses x-face averages for both directions and the S-based pressure.
  guard: compute_global_vmax takes sqrt(max(0, .)) where the reference has a bare sqrt
      (Weno5_2D.jl:1972-1973); identical whenever the reference result is not NaN.
"""
import numpy as np

NB = 3  # Weno5_2D.jl:9


class Params:
    def __init__(self, nx, ny=1, dx=None, dy=None, gam=5.0 / 3.0, cfl=0.8, eps=1e-6,
                 bc_x="outflow", bc_y="outflow", xsize=1.0, ysize=1.0):
        self.nx, self.ny = nx, ny
        self.Nx, self.Ny = nx + 2 * NB, ny + 2 * NB
        self.dx = xsize / nx if dx is None else dx
        self.dy = ysize / ny if dy is None else dy
        self.gam, self.cfl, self.eps = gam, cfl, eps
        self.bc_x, self.bc_y = bc_x, bc_y


# ----------------------------------------------------------------------------- cell centred
def primitives_E(q, gam):
    """Weno5_2D.jl:1573-1606.  q[...,8] = [rho,Mx,My,Mz,Bx,By,Bz,E] -> [rho,v,B,P]."""
    u = np.empty_like(q)
    u[..., 0] = q[..., 0]
    for c in (1, 2, 3):
        u[..., c] = q[..., c] / q[..., 0]
    u[..., 4:7] = q[..., 4:7]
    v2 = u[..., 1] * u[..., 1] + u[..., 2] * u[..., 2] + u[..., 3] * u[..., 3]
    B2 = u[..., 4] * u[..., 4] + u[..., 5] * u[..., 5] + u[..., 6] * u[..., 6]
    u[..., 7] = (gam - 1.0) * (q[..., 7] - 0.5 * (q[..., 0] * v2 + B2))
    return u


def eigenvalues(u, gam):
    """Weno5_2D.jl:1608-1667."""
    rho, vx, Bx, By, Bz, P = u[..., 0], u[..., 1], u[..., 4], u[..., 5], u[..., 6], u[..., 7]
    B2 = Bx * Bx + By * By + Bz * Bz
    cs2 = np.maximum(0.0, gam * np.abs(P / rho))
    bnx2 = np.maximum(0.0, Bx * Bx / rho)
    bbn2 = B2 / rho
    s = bbn2 + cs2
    root = np.sqrt(np.maximum(0.0, s * s - 4.0 * bnx2 * cs2))
    lf = np.sqrt(np.maximum(0.0, 0.5 * (bbn2 + cs2 + root)))
    ls = np.sqrt(np.maximum(0.0, 0.5 * (bbn2 + cs2 - root)))
    la = np.sqrt(bnx2)
    a = np.empty(u.shape[:-1] + (7,))
    a[..., 0] = vx - lf
    a[..., 1] = vx - la
    a[..., 2] = vx - ls
    a[..., 3] = vx
    a[..., 4] = vx + ls
    a[..., 5] = vx + la
    a[..., 6] = vx + lf
    return a


def fluxes_E(q, u):
    """Weno5_2D.jl:1533-1555.  Seven flux components (Bx omitted)."""
    rho, vx, vy, vz, Bx, By, Bz, P = [u[..., c] for c in range(8)]
    pt = P + 0.5 * (Bx * Bx + By * By + Bz * Bz)
    vB = vx * Bx + vy * By + vz * Bz
    F = np.empty(u.shape[:-1] + (7,))
    F[..., 0] = rho * vx
    F[..., 1] = q[..., 1] * vx + pt - Bx * Bx
    F[..., 2] = q[..., 2] * vx - Bx * By
    F[..., 3] = q[..., 3] * vx - Bx * Bz
    F[..., 4] = By * vx - Bx * vy
    F[..., 5] = Bz * vx - Bx * vz
    F[..., 6] = (q[..., 7] + pt) * vx - Bx * vB
    return F


# ----------------------------------------------------------------------------- face centred
def eigenvectors_E(uL, uR, EL, ER, gam, variant="2d"):
    """Left/right eigenvectors at the face between two cells.

    2D form: Weno5_2D.jl:1250-1531 (the ``_vec`` variant, which is the one called, :985).
    1D form: Weno5_1D.jl:784-1003 (differs in round-off only: af/as and Bt2 threshold).
    Returns L[..., m, c], R[..., c, m] with component order [rho,Mx,My,Mz,By,Bz,E].
    The sgnBt continuity flip (:1506-1528) is applied as in the reference although it
    cancels exactly in R.weno(L.).
    """
    g0, g1, g2 = 1.0 - gam, (gam - 1.0) / 2.0, (gam - 2.0) / (gam - 1.0)
    rho = 0.5 * (uL[..., 0] + uR[..., 0])
    r = np.sqrt(rho)
    vx = 0.5 * (uL[..., 1] + uR[..., 1])
    vy = 0.5 * (uL[..., 2] + uR[..., 2])
    vz = 0.5 * (uL[..., 3] + uR[..., 3])
    Bx = 0.5 * (uL[..., 4] + uR[..., 4])
    By = 0.5 * (uL[..., 5] + uR[..., 5])
    Bz = 0.5 * (uL[..., 6] + uR[..., 6])
    E = 0.5 * (EL + ER)                                    # conserved E, "q NOT u"
    B2 = Bx * Bx + By * By + Bz * Bz
    v2 = vx * vx + vy * vy + vz * vz
    pg = (gam - 1.0) * (E - 0.5 * (rho * v2 + B2))
    bbn2 = B2 / rho
    cs2 = np.maximum(0.0, gam * np.abs(pg / rho))
    bnx2 = Bx * Bx / rho
    a = np.sqrt(cs2)
    s_ = bbn2 + cs2
    root = np.sqrt(np.maximum(0.0, s_ * s_ - 4.0 * bnx2 * cs2))
    cf = np.sqrt(np.maximum(0.0, (bbn2 + cs2 + root) / 2.0))
    ca = np.sqrt(np.maximum(0.0, bnx2))
    cs = np.sqrt(np.maximum(0.0, (bbn2 + cs2 - root) / 2.0))
    Bt2 = By * By + Bz * Bz
    s = np.sign(Bx)
    s = np.where(s == 0.0, 1.0, s)
    with np.errstate(divide="ignore", invalid="ignore"):
        if variant == "2d":
            degenerate = Bt2 <= 1e-30
        else:
            degenerate = ~(Bt2 >= 1e-30)
        bty = np.where(degenerate, 1.0 / np.sqrt(2.0), By / np.sqrt(Bt2))
        btz = np.where(degenerate, 1.0 / np.sqrt(2.0), Bz / np.sqrt(Bt2))
        dl2 = cf * cf - cs * cs
        ok = dl2 >= 1e-30
        if variant == "2d":
            af = np.where(ok, np.sqrt(np.maximum(0.0, cs2 - cs * cs) / dl2), 1.0)
            as_ = np.where(ok, np.sqrt(np.maximum(0.0, cf * cf - cs2) / dl2), 1.0)
        else:
            af = np.where(ok, np.sqrt(np.maximum(0.0, cs2 - cs * cs)) / np.sqrt(dl2), 1.0)
            as_ = np.where(ok, np.sqrt(np.maximum(0.0, cf * cf - cs2)) / np.sqrt(dl2), 1.0)
    sig = bty * vy + btz * vz
    shp = rho.shape
    L = np.zeros(shp + (7, 7))
    R = np.zeros(shp + (7, 7))
    zero, one = np.zeros(shp), np.ones(shp)

    def rowL(m, vals):
        for c, v in enumerate(vals):
            L[..., m, c] = v

    def colR(m, vals):
        for c, v in enumerate(vals):
            R[..., c, m] = v

    # fast -, Alfven -, slow -, entropy, slow +, Alfven +, fast +
    rowL(0, [af * (g1 * v2 + cf * vx) - as_ * cs * sig * s, af * (g0 * vx - cf),
             g0 * af * vy + as_ * cs * bty * s, g0 * af * vz + as_ * cs * btz * s,
             g0 * af * By + a * as_ * bty * r, g0 * af * Bz + a * as_ * btz * r, -g0 * af])
    rowL(1, [btz * vy - bty * vz, zero, -btz, bty, -btz * s * r, bty * s * r, zero])
    rowL(2, [as_ * (g1 * v2 + cs * vx) + af * cf * sig * s, g0 * as_ * vx - as_ * cs,
             g0 * as_ * vy - af * cf * bty * s, g0 * as_ * vz - af * cf * btz * s,
             g0 * as_ * By - a * af * bty * r, g0 * as_ * Bz - a * af * btz * r, -g0 * as_])
    rowL(3, [-cs2 / g0 - 0.5 * v2, vx, vy, vz, By, Bz, -one])
    rowL(4, [as_ * (g1 * v2 - cs * vx) - af * cf * sig * s, as_ * (g0 * vx + cs),
             g0 * as_ * vy + af * cf * bty * s, g0 * as_ * vz + af * cf * btz * s,
             g0 * as_ * By - a * af * bty * r, g0 * as_ * Bz - a * af * btz * r, -g0 * as_])
    rowL(5, [btz * vy - bty * vz, zero, -btz, bty, btz * s * r, -bty * s * r, zero])
    rowL(6, [af * (g1 * v2 - cf * vx) + as_ * cs * sig * s, af * (g0 * vx + cf),
             g0 * af * vy - as_ * cs * bty * s, g0 * af * vz - as_ * cs * btz * s,
             g0 * af * By + a * as_ * bty * r, g0 * af * Bz + a * as_ * btz * r, -g0 * af])

    colR(0, [af, af * (vx - cf), af * vy + as_ * cs * bty * s, af * vz + as_ * cs * btz * s,
             a * as_ * bty / r, a * as_ * btz / r,
             af * (cf * cf - cf * vx + 0.5 * v2 - g2 * cs2) + as_ * cs * sig * s])
    colR(1, [zero, zero, -btz, bty, -btz * s / r, bty * s / r, bty * vz - btz * vy])
    colR(2, [as_, as_ * (vx - cs), as_ * vy - af * cf * bty * s, as_ * vz - af * cf * btz * s,
             -a * af * bty / r, -a * af * btz / r,
             as_ * (cs * cs - cs * vx + 0.5 * v2 - g2 * cs2) - af * cf * sig * s])
    colR(3, [one, vx, vy, vz, zero, zero, 0.5 * v2])
    colR(4, [as_, as_ * (vx + cs), as_ * vy + af * cf * bty * s, as_ * vz + af * cf * btz * s,
             -a * af * bty / r, -a * af * btz / r,
             as_ * (cs * cs + cs * vx + 0.5 * v2 - g2 * cs2) + af * cf * sig * s])
    colR(5, [zero, zero, -btz, bty, btz * s / r, -bty * s / r, bty * vz - btz * vy])
    colR(6, [af, af * (vx + cf), af * vy - as_ * cs * bty * s, af * vz - as_ * cs * btz * s,
             a * as_ * bty / r, a * as_ * btz / r,
             af * (cf * cf + cf * vx + 0.5 * v2 - g2 * cs2) - as_ * cs * sig * s])

    with np.errstate(divide="ignore", invalid="ignore"):
        scale = [0.5 / cs2, 0.5 * one, 0.5 / cs2, -g0 / cs2, 0.5 / cs2, 0.5 * one, 0.5 / cs2]
    for m in range(7):
        L[..., m, :] *= scale[m][..., None]

    # continuity flip
    if variant == "2d":
        sgnBt = np.sign(By)
        sgnBt = np.where(By == 0.0, np.sign(Bz), sgnBt)
    else:
        sgnBt = np.where(By != 0.0, np.sign(Bx), np.sign(Bz))
    sgnBt = np.where(sgnBt == 0.0, 1.0, sgnBt)
    slowflip = np.where(a >= ca, sgnBt, 1.0)
    fastflip = np.where(a >= ca, 1.0, sgnBt)
    for m, flip in ((2, slowflip), (4, slowflip), (0, fastflip), (6, fastflip)):
        L[..., m, :] *= flip[..., None]
        R[..., :, m] *= flip[..., None]
    return L, R


# ----------------------------------------------------------------------------- S branch
def primitives_S(q, gam):
    """Weno5_2D.jl:953-969: q = [rho,Mx,My,Mz,Bx,By,Bz,S] -> u = [rho,vx,vy,vz,Bx,By,Bz,P(S)]."""
    u = q.copy()
    u[..., 1] = u[..., 1] / u[..., 0]
    u[..., 2] = u[..., 2] / u[..., 0]
    u[..., 3] = u[..., 3] / u[..., 0]
    u[..., 7] = u[..., 7] * u[..., 0] ** (gam - 1.0)
    return u


def fluxes_S(q, u):
    """Weno5_2D.jl:927-950: like compute_fluxes_E, the 7th flux is S*vx."""
    F = np.zeros(u.shape[:-1] + (7,))
    pt = u[..., 7] + 0.5 * (u[..., 4] * u[..., 4] + u[..., 5] * u[..., 5] + u[..., 6] * u[..., 6])
    F[..., 0] = u[..., 0] * u[..., 1]
    F[..., 1] = q[..., 1] * u[..., 1] + pt - u[..., 4] * u[..., 4]
    F[..., 2] = q[..., 2] * u[..., 1] - u[..., 4] * u[..., 5]
    F[..., 3] = q[..., 3] * u[..., 1] - u[..., 4] * u[..., 6]
    F[..., 4] = u[..., 5] * u[..., 1] - u[..., 4] * u[..., 2]
    F[..., 5] = u[..., 6] * u[..., 1] - u[..., 4] * u[..., 3]
    F[..., 6] = q[..., 7] * u[..., 1]
    return F


def eigenvectors_S(uL, uR, gam):
    """Weno5_2D.jl:640-924 (the ``_vec`` variant, the one called at :373): entropy-variable eigenvectors at
    the face between two cells.  The face density is the mean of rho^(1-gam) (":657-667"), the pressure the
    arithmetic mean of P(S).  Component order [rho,Mx,My,Mz,By,Bz,S]."""
    g0, gS = 1.0 - gam, (gam - 1.0) / gam
    drho = np.abs(uL[..., 0] - uR[..., 0])
    with np.errstate(divide="ignore", invalid="ignore"):
        rho_log = np.abs(g0 * drho / (uR[..., 0] ** g0 - uL[..., 0] ** g0)) ** (1.0 / gam)
    rho = np.where(drho <= 1e-12, 0.5 * (uL[..., 0] + uR[..., 0]), rho_log)
    vx = 0.5 * (uL[..., 1] + uR[..., 1])
    vy = 0.5 * (uL[..., 2] + uR[..., 2])
    vz = 0.5 * (uL[..., 3] + uR[..., 3])
    Bx = 0.5 * (uL[..., 4] + uR[..., 4])
    By = 0.5 * (uL[..., 5] + uR[..., 5])
    Bz = 0.5 * (uL[..., 6] + uR[..., 6])
    pg = 0.5 * (uL[..., 7] + uR[..., 7])
    r = np.sqrt(rho)
    S = pg / rho ** (gam - 1.0)
    B2 = Bx * Bx + By * By + Bz * Bz
    bbn2 = B2 / rho
    cs2 = np.maximum(0.0, gam * np.abs(pg / rho))
    bnx2 = Bx * Bx / rho
    a = np.sqrt(cs2)
    s_ = bbn2 + cs2
    root = np.sqrt(np.maximum(0.0, s_ * s_ - 4.0 * bnx2 * cs2))
    cf = np.sqrt(np.maximum(0.0, (bbn2 + cs2 + root) / 2.0))
    ca = np.sqrt(np.maximum(0.0, bnx2))
    cs = np.sqrt(np.maximum(0.0, (bbn2 + cs2 - root) / 2.0))
    Bt2 = By * By + Bz * Bz
    s = np.sign(Bx)
    s = np.where(s == 0.0, 1.0, s)
    with np.errstate(divide="ignore", invalid="ignore"):
        degenerate = Bt2 <= 1e-30
        bty = np.where(degenerate, 1.0 / np.sqrt(2.0), By / np.sqrt(Bt2))
        btz = np.where(degenerate, 1.0 / np.sqrt(2.0), Bz / np.sqrt(Bt2))
        dl2 = cf * cf - cs * cs
        ok = dl2 >= 1e-30
        af = np.where(ok, np.sqrt(np.maximum(0.0, cs2 - cs * cs) / dl2), 1.0)
        as_ = np.where(ok, np.sqrt(np.maximum(0.0, cf * cf - cs2) / dl2), 1.0)
    sig = bty * vy + btz * vz
    shp = rho.shape
    L = np.zeros(shp + (7, 7))
    R = np.zeros(shp + (7, 7))
    zero, one = np.zeros(shp), np.ones(shp)
    kap = rho / (gam * S)

    def rowL(m, vals):
        for c, v in enumerate(vals):
            L[..., m, c] = v

    def colR(m, vals):
        for c, v in enumerate(vals):
            R[..., c, m] = v

    rowL(0, [af * (gS * cs2 + cf * vx) - as_ * cs * sig * s, -af * cf, as_ * cs * bty * s, as_ * cs * btz * s,
             a * as_ * bty * r, a * as_ * btz * r, af * cs2 * kap])
    rowL(1, [btz * vy - bty * vz, zero, -btz, bty, -btz * s * r, bty * s * r, zero])
    rowL(2, [as_ * (gS * cs2 + cs * vx) + af * cf * sig * s, -as_ * cs, -af * cf * bty * s, -af * cf * btz * s,
             -a * af * bty * r, -a * af * btz * r, as_ * cs2 * kap])
    rowL(3, [one / gam, zero, zero, zero, zero, zero, -kap])
    rowL(4, [as_ * (gS * cs2 - cs * vx) - af * cf * sig * s, as_ * cs, af * cf * bty * s, af * cf * btz * s,
             -a * af * bty * r, -a * af * btz * r, as_ * cs2 * kap])
    rowL(5, [btz * vy - bty * vz, zero, -btz, bty, btz * s * r, -bty * s * r, zero])
    rowL(6, [af * (gS * cs2 - cf * vx) + as_ * cs * sig * s, af * cf, -as_ * cs * bty * s, -as_ * cs * btz * s,
             a * as_ * bty * r, a * as_ * btz * r, af * cs2 * kap])
    Sr = S / rho
    colR(0, [af, af * (vx - cf), af * vy + as_ * cs * bty * s, af * vz + as_ * cs * btz * s,
             a * as_ * bty / r, a * as_ * btz / r, af * Sr])
    colR(1, [zero, zero, -btz, bty, -btz * s / r, bty * s / r, zero])
    colR(2, [as_, as_ * (vx - cs), as_ * vy - af * cf * bty * s, as_ * vz - af * cf * btz * s,
             -a * af * bty / r, -a * af * btz / r, as_ * Sr])
    colR(3, [one, vx, vy, vz, zero, zero, g0 * Sr])
    colR(4, [as_, as_ * (vx + cs), as_ * vy + af * cf * bty * s, as_ * vz + af * cf * btz * s,
             -a * af * bty / r, -a * af * btz / r, as_ * Sr])
    colR(5, [zero, zero, -btz, bty, btz * s / r, -bty * s / r, zero])
    colR(6, [af, af * (vx + cf), af * vy - as_ * cs * bty * s, af * vz - as_ * cs * btz * s,
             a * as_ * bty / r, a * as_ * btz / r, af * Sr])
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = [0.5 / cs2, 0.5 * one, 0.5 / cs2, one, 0.5 / cs2, 0.5 * one, 0.5 / cs2]    # row 4 unscaled, :872
    for m in range(7):
        L[..., m, :] *= scale[m][..., None]
    sgnBt = np.sign(By)
    sgnBt = np.where(By == 0.0, np.sign(Bz), sgnBt)
    sgnBt = np.where(sgnBt == 0.0, 1.0, sgnBt)
    slowflip = np.where(a >= ca, sgnBt, 1.0)
    fastflip = np.where(a >= ca, 1.0, sgnBt)
    for m, flip in ((2, slowflip), (4, slowflip), (0, fastflip), (6, fastflip)):
        L[..., m, :] *= flip[..., None]
        R[..., :, m] *= flip[..., None]
    return L, R


def _ext_index(idx, Nq, bc):
    """Continuation of a stencil index past the high end of the array (quirk Q3)."""
    n = Nq - 2 * NB
    if bc == "periodic":
        return np.where(idx >= Nq, idx - n, idx)
    return np.minimum(idx, Nq - 1)


def weno5_face_flux(q, a, F, gam, eps, bc, variant="2d", branch="E"):
    """Numerical flux at faces fi = NB-1 .. Nq-NB (0-based; reference 1-based 3..Nq-2).

    Weno5_2D.jl:1795-1895 (1D twin Weno5_1D.jl:287-405).  q is (Nq, Nr, 8) and the sweep
    runs along axis 0.  Returns dF (Nq, Nr, 7), zero at faces that are not computed.
    """
    Nq = q.shape[0]
    faces = np.arange(NB - 1, Nq - NB + 1)
    if branch == "E":
        u = primitives_E(q, gam)
        L, R = eigenvectors_E(u[faces], u[faces + 1], q[faces, ..., 7], q[faces + 1, ..., 7], gam, variant)
    else:
        u = primitives_S(q, gam)
        L, R = eigenvectors_S(u[faces], u[faces + 1], gam)
    q7 = q[..., [0, 1, 2, 3, 5, 6, 7]]
    Fs = np.zeros(L.shape[:-2] + (7,))
    for m in range(7):
        w, v = [], []
        for k in range(6):
            idx = _ext_index(faces - 2 + k, Nq, bc)
            Fk, qk = F[idx], q7[idx]
            accw = L[..., m, 0] * Fk[..., 0]
            accv = L[..., m, 0] * qk[..., 0]
            for c in range(1, 7):
                accw = accw + L[..., m, c] * Fk[..., c]
                accv = accv + L[..., m, c] * qk[..., c]
            w.append(accw)
            v.append(accv)
        first = (-w[1] + 7.0 * w[2] + 7.0 * w[3] - w[4]) / 12.0
        dw = [w[k + 1] - w[k] for k in range(5)]
        dv = [v[k + 1] - v[k] for k in range(5)]
        amax = np.maximum(np.abs(a[faces, ..., m]), np.abs(a[faces + 1, ..., m]))

        def phi(at, bt, ct, dt_):
            IS0 = 13.0 * (at - bt) ** 2 + 3.0 * (at - 3.0 * bt) ** 2
            IS1 = 13.0 * (bt - ct) ** 2 + 3.0 * (bt + ct) ** 2
            IS2 = 13.0 * (ct - dt_) ** 2 + 3.0 * (3.0 * ct - dt_) ** 2
            al0 = 1.0 / (eps + IS0) ** 2
            al1 = 6.0 / (eps + IS1) ** 2
            al2 = 3.0 / (eps + IS2) ** 2
            om0 = al0 / (al0 + al1 + al2)
            om2 = al2 / (al0 + al1 + al2)
            return om0 * (at - 2.0 * bt + ct) / 3.0 + (om2 - 0.5) * (bt - 2.0 * ct + dt_) / 6.0

        second = phi(*[0.5 * (dw[k] + amax * dv[k]) for k in (0, 1, 2, 3)])
        third = phi(*[0.5 * (dw[k] - amax * dv[k]) for k in (4, 3, 2, 1)])
        Fs[..., m] = first - second + third
    dF = np.zeros(q.shape[:-1] + (7,))
    for c in range(7):
        acc = Fs[..., 0] * R[..., c, 0]
        for m in range(1, 7):
            acc = acc + Fs[..., m] * R[..., c, m]
        dF[faces, ..., c] = acc
    return dF, u


def flux_difference_S(q, gam, eps, bc):
    """Weno5_2D.jl:361-401: the entropy-variable twin of flux_difference_E (q[...,7] = S)."""
    return flux_difference_E(q, gam, eps, bc, branch="S")


def flux_difference_E(q, gam, eps, bc, variant="2d", branch="E"):
    """Weno5_2D.jl:973-1018: one directional sweep along axis 0 of q (Nq, Nr, 8)."""
    Nq, Nr = q.shape[0], q.shape[1]
    if branch == "E":
        u = primitives_E(q, gam)
        F = fluxes_E(q, u)
    else:
        u = primitives_S(q, gam)
        F = fluxes_S(q, u)
    a = eigenvalues(u, gam)
    dF, _ = weno5_face_flux(q, a, F, gam, eps, bc, variant, branch)
    dq = np.zeros((Nq, Nr, 8))
    bsy = np.zeros((Nq, Nr))
    bsz = np.zeros((Nq, Nr))
    i = np.arange(NB - 1, Nq - NB + 1)          # 1-based NBnd .. Nq-NBnd+1
    jl, jh = (NB - 1, Nr - NB + 1) if Nr > 1 else (0, 1)
    js = slice(jl, jh)
    for c_out, c_in in ((0, 0), (1, 1), (2, 2), (3, 3), (5, 4), (6, 5), (7, 6)):
        dq[i, js, c_out] = dF[i, js, c_in] - dF[i - 1, js, c_in]
    bsy[i, js] = dF[i, js, 4] + 0.5 * (u[i, js, 4] * u[i, js, 2] + u[i + 1, js, 4] * u[i + 1, js, 2])
    bsz[i, js] = dF[i, js, 5] + 0.5 * (u[i, js, 4] * u[i, js, 3] + u[i + 1, js, 4] * u[i + 1, js, 3])
    return dq, bsy, bsz


# ----------------------------------------------------------------------------- grid services
def rotate_fwd(q):
    """Weno5_2D.jl:1903-1910: (x,y,z) -> (y,z,x) component permutation + transpose."""
    return np.ascontiguousarray(np.transpose(q[..., [0, 2, 3, 1, 5, 6, 4, 7]], (1, 0, 2)))


def rotate_bwd(p):
    """Weno5_2D.jl:1914-1921."""
    return np.ascontiguousarray(np.transpose(p[..., [0, 3, 1, 2, 6, 4, 5, 7]], (1, 0, 2)))


def _fill_axis(arr, axis, bc):
    N = arr.shape[axis]
    n = N - 2 * NB
    a = np.moveaxis(arr, axis, 0)
    if bc == "periodic":
        a[0:NB] = a[n:n + NB]
        a[N - NB:N] = a[NB:2 * NB]
    else:
        a[0:NB] = a[NB]
        a[N - NB:N] = a[N - NB - 1]


def boundaries_q(q, P):
    """Weno5_2D.jl:1686-1702 (outflow) + periodic addition.  x first, then y."""
    _fill_axis(q, 0, P.bc_x)
    if q.shape[1] > 1:
        _fill_axis(q, 1, P.bc_y)


def boundaries_flux(f, P):
    """Weno5_2D.jl:1704-1729 for one of the two 2-D arrays."""
    Nx, Ny = f.shape
    if P.bc_x == "periodic":
        _fill_axis(f, 0, "periodic")
    else:
        f[1, :] = f[2, :]
        f[0, :] = f[2, :]
    if P.bc_y == "periodic":
        _fill_axis(f, 1, "periodic")
    else:
        f[:, 1] = f[:, 2]
        f[:, 0] = f[:, 2]
    if P.bc_x != "periodic":
        f[Nx - 2, :] = f[Nx - 3, :]
        f[Nx - 1, :] = f[Nx - 3, :]
    if P.bc_y != "periodic":
        f[:, Ny - 2] = f[:, Ny - 3]
        f[:, Ny - 1] = f[:, Ny - 3]
    # rim (note reference order: fsy[1,:] = fsy[:,1] = 0 then fsy[Nx,:] = fsy[:,Ny] = 0)
    if P.bc_x != "periodic":
        f[0, :] = 0.0
        f[Nx - 1, :] = 0.0
    if P.bc_y != "periodic":
        f[:, 0] = 0.0
        f[:, Ny - 1] = 0.0


def boundaries(q, f, g, P):
    boundaries_q(q, P)
    boundaries_flux(f, P)
    boundaries_flux(g, P)


def corner_fluxes(fsy, gsx):
    """Weno5_2D.jl:319-342."""
    Nx, Ny = fsy.shape
    Ox = np.zeros((Nx, Ny))
    Oxi = np.zeros((Nx, Ny))
    Oxj = np.zeros((Nx, Ny))
    Ox[1:Nx - 1, 1:Ny - 1] = 0.5 * (gsx[1:Nx - 1, 1:Ny - 1] + gsx[2:Nx, 1:Ny - 1]
                                    - fsy[1:Nx - 1, 1:Ny - 1] - fsy[1:Nx - 1, 2:Ny])
    Oxi[1:Nx - 1, 1:Ny - 1] = Ox[0:Nx - 2, 1:Ny - 1]
    Oxj[1:Nx - 1, 1:Ny - 1] = Ox[1:Nx - 1, 0:Ny - 2]
    return Ox, Oxi, Oxj


def face2center(bxb, byb):
    """Weno5_2D.jl:344-356."""
    Nx, Ny = bxb.shape
    Bxy = np.zeros((Nx, Ny, 2))
    i = slice(NB, Nx - 2)       # 1-based NBnd+1 .. Nx-2
    j = slice(NB, Ny - 2)
    Bxy[i, j, 0] = (-bxb[NB - 2:Nx - 4, j] + 9.0 * bxb[NB - 1:Nx - 3, j] + 9.0 * bxb[i, j] - bxb[NB + 1:Nx - 1, j]) / 16.0
    Bxy[i, j, 1] = (-byb[i, NB - 2:Ny - 4] + 9.0 * byb[i, NB - 1:Ny - 3] + 9.0 * byb[i, j] - byb[i, NB + 1:Ny - 1]) / 16.0
    return Bxy


def center2face(q):
    """Weno5_2D.jl:1669-1682."""
    Nx, Ny = q.shape[0], q.shape[1]
    bxb = np.zeros((Nx, Ny))
    byb = np.zeros((Nx, Ny))
    i = slice(1, Nx - 2)        # 1-based 2 .. Nx-2
    j = slice(1, Ny - 2)
    bxb[i, j] = (-q[0:Nx - 3, j, 4] + 9 * q[i, j, 4] + 9 * q[2:Nx - 1, j, 4] - q[3:Nx, j, 4]) / 16
    byb[i, j] = (-q[i, 0:Ny - 3, 5] + 9 * q[i, j, 5] + 9 * q[i, 2:Ny - 1, 5] - q[i, 3:Ny, 5]) / 16
    return bxb, byb


def E2S(q, gam):
    """Weno5_2D.jl:1770-1781 (component 7 holds E)."""
    rho = q[..., 0]
    mom2 = q[..., 1] ** 2 + q[..., 2] ** 2 + q[..., 3] ** 2
    B2 = q[..., 4] ** 2 + q[..., 5] ** 2 + q[..., 6] ** 2
    return (q[..., 7] - 0.5 * mom2 / rho - B2 / 2) * (gam - 1) / rho ** (gam - 1)


def S2E(q, gam):
    """Weno5_2D.jl:1783-1793 (component 7 holds S)."""
    rho = q[..., 0]
    mom2 = q[..., 1] ** 2 + q[..., 2] ** 2 + q[..., 3] ** 2
    B2 = q[..., 4] ** 2 + q[..., 5] ** 2 + q[..., 6] ** 2
    return rho ** (gam - 1) * q[..., 7] / (gam - 1) + B2 / 2 + 0.5 * mom2 / rho


def state2conserved(rho, vx, vy, vz, Bx, By, Bz, Pr, gam=5.0 / 3.0):
    """Weno5_2D.jl:2043-2053."""
    v2 = vx ** 2 + vy ** 2 + vz ** 2
    B2 = Bx ** 2 + By ** 2 + Bz ** 2
    E = Pr / (gam - 1) + 0.5 * (rho * v2 + B2)
    S = Pr / rho ** (gam - 1)
    return np.array([rho, rho * vx, rho * vy, rho * vz, Bx, By, Bz, S, E])


def compute_global_vmax(q, P):
    """Weno5_2D.jl:1942-1984 (x-face averages for both speeds, S-based pressure)."""
    Nx, Ny = q.shape[0], q.shape[1]
    i = slice(NB, Nx - NB)
    ip = slice(NB + 1, Nx - NB + 1)
    j = slice(NB, Ny - NB) if Ny > 1 else slice(0, 1)
    avg = lambda c: (q[i, j, c] + q[ip, j, c]) / 2
    rho = avg(0)
    vx, vy = avg(1) / rho, avg(2) / rho
    Bx, By, Bz, S = avg(4), avg(5), avg(6), avg(7)
    pres = S * rho ** (P.gam - 1)
    cs2 = P.gam * np.abs(pres / rho)
    B2 = Bx ** 2 + By ** 2 + Bz ** 2
    bbn2, bnx2, bny2 = B2 / rho, Bx ** 2 / rho, By ** 2 / rho
    rootx = np.sqrt(np.maximum(0.0, (bbn2 + cs2) ** 2 - 4 * bnx2 * cs2))
    rooty = np.sqrt(np.maximum(0.0, (bbn2 + cs2) ** 2 - 4 * bny2 * cs2))
    lfx = np.sqrt(0.5 * (bbn2 + cs2 + rootx))
    lfy = np.sqrt(0.5 * (bbn2 + cs2 + rooty))
    return float(np.max(np.abs(vx) + lfx)), float(np.max(np.abs(vy) + lfy))


def timestep(q, P):
    """Weno5_2D.jl:1930-1940."""
    vxmax, vymax = compute_global_vmax(q, P)
    dtx, dty = P.dx / vxmax, P.dy / vymax
    return P.cfl / (1 / dtx + 1 / dty), vxmax, vymax


def compute_divB(bxb, byb, P):
    """Weno5_2D.jl:1986-1999."""
    Nx, Ny = bxb.shape
    d = np.zeros((Nx, Ny))
    i = slice(NB - 1, Nx - NB)
    j = slice(NB - 1, Ny - NB)
    d[i, j] = (bxb[i, j] - bxb[NB - 2:Nx - NB - 1, j]) / P.dx + (byb[i, j] - byb[i, NB - 2:Ny - NB - 1]) / P.dy
    return d


# ----------------------------------------------------------------------------- time integration
def _stage_2d(qk9, base8, bbx, bby, c, dt, P, stage):
    """One E-branch RK stage of Weno5_2D.jl:134-180 and the residue of ES_Switch.

    The order of the two flux terms follows the reference stage by stage (quirk Q11)."""
    qE = np.ascontiguousarray(qk9[..., [0, 1, 2, 3, 4, 5, 6, 8]])
    dFx, fsy, _ = flux_difference_E(qE, P.gam, P.eps, P.bc_x)
    dFy_rot, _, gsx_rot = flux_difference_E(rotate_fwd(qE), P.gam, P.eps, P.bc_y)
    dFy = rotate_bwd(dFy_rot)
    gsx = np.ascontiguousarray(gsx_rot.T)
    if stage <= 2:
        qn = base8 - c * dt / P.dy * dFy - c * dt / P.dx * dFx          # :145, :190
    elif stage == 3:
        qn = base8 - dt / P.dx * dFx - dt / P.dy * dFy                  # :236
    else:
        qn = base8 + (-1 / 6 * dt / P.dx * dFx - 1 / 6 * dt / P.dy * dFy)   # :279-280
    boundaries(qn, fsy, gsx, P)
    Ox, Oxi, Oxj = corner_fluxes(fsy, gsx)
    bxn = bbx - c * dt / P.dy * (Ox - Oxj)
    byn = bby - c * dt / P.dx * (Oxi - Ox)
    qn[..., 4:6] = face2center(bxn, byn)
    boundaries(qn, fsy, gsx, P)
    q9 = np.zeros(qk9.shape)
    q9[..., 0:7] = qn[..., 0:7]
    q9[..., 7] = E2S(qn, P.gam)
    q9[..., 8] = S2E(q9[..., 0:8], P.gam)
    boundaries_q(q9, P)
    return q9, bxn, byn


def classical_RK4_step_2d(q0, bxb0, byb0, dt, P):
    """Weno5_2D.jl:125-317, E branch.  Returns (q4, bxb4, byb4)."""
    idxE = [0, 1, 2, 3, 4, 5, 6, 8]
    q0E = np.ascontiguousarray(q0[..., idxE])
    q1, bx1, by1 = _stage_2d(q0, q0E, bxb0, byb0, 0.5, dt, P, 1)
    q2, bx2, by2 = _stage_2d(q1, q0E, bxb0, byb0, 0.5, dt, P, 2)
    q3, bx3, by3 = _stage_2d(q2, q0E, bxb0, byb0, 1.0, dt, P, 3)
    base = 1 / 3 * (-q0E + q1[..., idxE] + 2 * q2[..., idxE] + q3[..., idxE])
    bbx = 1 / 3 * (-bxb0 + bx1 + 2 * bx2 + bx3)
    bby = 1 / 3 * (-byb0 + by1 + 2 * by2 + by3)
    q4, bx4, by4 = _stage_2d(q3, base, bbx, bby, 1.0 / 6.0, dt, P, 4)
    boundaries_flux(bx4, P)          # Weno5_2D.jl:313 boundaries!(q4, bxb4, byb4)
    boundaries_flux(by4, P)
    return q4, bx4, by4


# ----------------------------------------------------------------------------- dual-energy step
def ES_Switch(q_E, q_S, fsy_S, gsx_S, fsy_E, gsx_E, gam, threshold=1e-5):
    """Weno5_2D.jl:1734-1768 with the reference's own (commented-out) criterion restored:
    ``dS = abs.(q_SE[:,:,8] - q_S[:,:,8])`` instead of ``dS = 1``.  Cells (1-based i,j >= 2) whose entropy
    derived from the E state differs from the advected entropy by >= threshold take the E state and pull the
    E fluxes of their own and their low-side neighbour faces; all others keep the S state."""
    Nx, Ny = q_E.shape[:2]
    q_SE = q_E.copy()
    q_SE[..., 7] = E2S(q_E, gam)
    dS = np.abs(q_SE[..., 7] - q_S[..., 7])
    bad = dS >= threshold
    bad[0, :] = False
    bad[:, 0] = False
    q = np.zeros((Nx, Ny, 9))
    q[..., 0:8] = np.where(bad[..., None], q_SE, q_S)
    fE = bad.copy()
    fE[:-1, :] |= bad[1:, :]                     # fsy[i-1,j] of a bad cell
    gE = bad.copy()
    gE[:, :-1] |= bad[:, 1:]                     # gsx[i,j-1] of a bad cell
    fsy = np.where(fE, fsy_E, fsy_S)
    gsx = np.where(gE, gsx_E, gsx_S)
    q[..., 8] = S2E(q[..., 0:8], gam)
    return q, fsy, gsx, bad


def _stage_2d_dual(qk9, baseE, baseS, bbx, bby, c, dt, P, stage, threshold):
    """One stage of Weno5_2D.jl:134-180 with both branches live."""
    def ct(qn, fsy, gsx):
        boundaries(qn, fsy, gsx, P)
        Ox, Oxi, Oxj = corner_fluxes(fsy, gsx)
        bxn = bbx - c * dt / P.dy * (Ox - Oxj)
        byn = bby - c * dt / P.dx * (Oxi - Ox)
        qn[..., 4:6] = face2center(bxn, byn)
        boundaries(qn, fsy, gsx, P)
        return bxn, byn

    qE = np.ascontiguousarray(qk9[..., [0, 1, 2, 3, 4, 5, 6, 8]])
    qS = np.ascontiguousarray(qk9[..., 0:8])
    out = []
    for branch, qb, base in (("E", qE, baseE), ("S", qS, baseS)):
        dFx, fsy, _ = flux_difference_E(qb, P.gam, P.eps, P.bc_x, branch=branch)
        dFy_rot, _, gsx_rot = flux_difference_E(rotate_fwd(qb), P.gam, P.eps, P.bc_y, branch=branch)
        dFy = rotate_bwd(dFy_rot)
        gsx = np.ascontiguousarray(gsx_rot.T)
        y_first = branch == "E" and stage <= 2                         # :145,:190 vs :164,:210
        if stage <= 3:
            cc = c if stage <= 2 else 1.0
            qn = (base - cc * dt / P.dy * dFy - cc * dt / P.dx * dFx) if y_first else \
                 (base - cc * dt / P.dx * dFx - cc * dt / P.dy * dFy)
        else:
            qn = base + (-1 / 6 * dt / P.dx * dFx - 1 / 6 * dt / P.dy * dFy)
        ct(qn, fsy, gsx)
        out.append((qn, fsy, gsx))
    (qnE, fsyE, gsxE), (qnS, fsyS, gsxS) = out
    q9, fsy, gsx, bad = ES_Switch(qnE, qnS, fsyS, gsxS, fsyE, gsxE, P.gam, threshold)
    q8 = np.ascontiguousarray(q9[..., 0:8])
    bxn, byn = ct(q8, fsy, gsx)
    q9[..., 0:8] = q8
    # the reference leaves q[:,:,9] = S2E(q) from before the final face->centre update (:1765, :176-180)
    boundaries_q(q9, P)
    return q9, bxn, byn, bad


def classical_RK4_step_2d_dual(q0, bxb0, byb0, dt, P, threshold=1e-5):
    """Weno5_2D.jl:125-317 with a working ES_Switch (see ES_Switch above).  Returns (q4, bxb4, byb4, nbad) with nbad = interior cells that took the E state, per stage."""
    idxE = [0, 1, 2, 3, 4, 5, 6, 8]
    q0E = np.ascontiguousarray(q0[..., idxE])
    q0S = np.ascontiguousarray(q0[..., 0:8])
    q1, bx1, by1, b1 = _stage_2d_dual(q0, q0E, q0S, bxb0, byb0, 0.5, dt, P, 1, threshold)
    q2, bx2, by2, b2 = _stage_2d_dual(q1, q0E, q0S, bxb0, byb0, 0.5, dt, P, 2, threshold)
    q3, bx3, by3, b3 = _stage_2d_dual(q2, q0E, q0S, bxb0, byb0, 1.0, dt, P, 3, threshold)
    baseE = 1 / 3 * (-q0E + q1[..., idxE] + 2 * q2[..., idxE] + q3[..., idxE])
    baseS = 1 / 3 * (-q0S + q1[..., 0:8] + 2 * q2[..., 0:8] + q3[..., 0:8])
    bbx = 1 / 3 * (-bxb0 + bx1 + 2 * bx2 + bx3)
    bby = 1 / 3 * (-byb0 + by1 + 2 * by2 + by3)
    q4, bx4, by4, b4 = _stage_2d_dual(q3, baseE, baseS, bbx, bby, 1.0 / 6.0, dt, P, 4, threshold)
    boundaries_flux(bx4, P)
    boundaries_flux(by4, P)
    return q4, bx4, by4, [int(b[NB:-NB, NB:-NB].sum()) for b in (b1, b2, b3, b4)]     # interior cells switched to E


# ----------------------------------------------------------------------------- snapshot hook
def arr2img_index(arr, ncolours, range_=(0.0, 0.0), log=False):
    """Arr2Img, JColorMaps.jl:92-146, up to the colour-table lookup: the 1-based colour index of every pixel
    of the transposed field (Int16 in the reference).  `range_ == (0,0)` selects the automatic range over the
    finite pixels; with `log`, non-positive pixels are first set to the smallest positive one.  NaN pixels
    take the first colour.  (The reference's `sqrt` option cannot run: its keyword shadows Base.sqrt.)"""
    img = np.array(arr, dtype=np.float64).T.copy()
    if log:
        img[img <= 0] = img[img > 0].min()
    lo, hi = float(range_[0]), float(range_[1])
    if lo == 0.0 and hi == 0.0:
        good = img[np.isfinite(img)]
        lo, hi = good.min(), good.max()
    if log:
        img = np.log10(img)
        lo, hi = np.log10(lo), np.log10(hi)
    img[np.isnan(img)] = lo
    r = np.rint((img - lo) / (hi - lo) * (ncolours - 1)) + 1
    return np.clip(r, 1, ncolours).astype(np.int16)


def boundaries_1d(q, P):
    """Weno5_1D.jl:1076-1090."""
    _fill_axis(q, 0, P.bc_x)


def classical_RK4_step_1d(q, dt, P):
    """Weno5_1D.jl:152-236 with the E-only override of ES_Switch (:252-254)."""
    def sweep(q9):
        qE = np.ascontiguousarray(q9[:, None, [0, 1, 2, 3, 4, 5, 6, 8]])
        u = primitives_E(qE, P.gam)
        a = eigenvalues(u, P.gam)
        F = fluxes_E(qE, u)
        dF, _ = weno5_face_flux(qE, a, F, P.gam, P.eps, P.bc_x, variant="1d")
        N = q9.shape[0]
        Q = np.zeros((N, 8))
        i = np.arange(NB, N - NB)
        for c_out, c_in in ((0, 0), (1, 1), (2, 2), (3, 3), (5, 4), (6, 5), (7, 6)):
            Q[i, c_out] = dF[i, 0, c_in] - dF[i - 1, 0, c_in]
        return qE[:, 0, :], Q

    def switch(qE):
        q9 = np.zeros((qE.shape[0], 9))
        q9[:, 0:7] = qE[:, 0:7]
        q9[:, 7] = E2S(qE, P.gam)
        q9[:, 8] = qE[:, 7]
        boundaries_1d(q9, P)
        return q9

    q0E, Q0 = sweep(q)
    q1 = switch(q0E - 1 / 2 * dt / P.dx * Q0)
    q1E, Q1 = sweep(q1)
    q2 = switch(q0E - 1 / 2 * dt / P.dx * Q1)
    q2E, Q2 = sweep(q2)
    q3 = switch(q0E - dt / P.dx * Q2)
    q3E, Q3 = sweep(q3)
    q4 = switch(1 / 3 * (-q0E + q1E + 2 * q2E + q3E - 1 / 2 * dt / P.dx * Q3))
    return q4


def classical_RK4_step_1d_dual(q, dt, P, threshold=1e-5, live=True):
    """Weno5_1D.jl:152-236 with BOTH branches, as the reference runs them: every stage sweeps the switched state
    once in energy variables (weno5_flux_difference_E, :728) and once in entropy variables
    (weno5_flux_difference_S, :441), and ES_Switch (:238-258) flags the cells with |S(E) - S| >= threshold.
    live = False is the code as shipped -- the flags are computed and returned (`idx`), the state is overridden by
    the energy branch (:252-254); live = True drops those three override lines: entropy state by default, energy
    state in flagged cells.  Returns (q4, flag of the last stage)."""
    def sweep(q9, branch):
        cols = [0, 1, 2, 3, 4, 5, 6, 8] if branch == "E" else [0, 1, 2, 3, 4, 5, 6, 7]
        qb = np.ascontiguousarray(q9[:, None, cols])
        if branch == "E":
            u = primitives_E(qb, P.gam)
            F = fluxes_E(qb, u)
        else:
            u = primitives_S(qb, P.gam)
            F = fluxes_S(qb, u)
        a = eigenvalues(u, P.gam)
        dF, _ = weno5_face_flux(qb, a, F, P.gam, P.eps, P.bc_x, variant="1d", branch=branch)
        N = q9.shape[0]
        Q = np.zeros((N, 8))
        i = np.arange(NB, N - NB)
        for c_out, c_in in ((0, 0), (1, 1), (2, 2), (3, 3), (5, 4), (6, 5), (7, 6)):
            Q[i, c_out] = dF[i, 0, c_in] - dF[i - 1, 0, c_in]
        return qb[:, 0, :], Q

    def switch(qE, qS):
        qSE = qE.copy()
        qSE[:, 7] = E2S(qE, P.gam)
        bad = np.abs(qSE[:, 7] - qS[:, 7]) >= threshold
        q9 = np.zeros((qE.shape[0], 9))
        q9[:, 0:8] = np.where(bad[:, None], qSE, qS)
        q9[:, 8] = S2E(q9[:, 0:8], P.gam)
        if not live:
            q9[:, 0:7] = qE[:, 0:7]
            q9[:, 7] = E2S(qE, P.gam)
            q9[:, 8] = qE[:, 7]
        boundaries_1d(q9, P)
        return q9, bad

    q0E, Q0E = sweep(q, "E")
    q0S, Q0S = sweep(q, "S")
    q1, _ = switch(q0E - 1 / 2 * dt / P.dx * Q0E, q0S - 1 / 2 * dt / P.dx * Q0S)
    q1E, Q1E = sweep(q1, "E")
    q1S, Q1S = sweep(q1, "S")
    q2, _ = switch(q0E - 1 / 2 * dt / P.dx * Q1E, q0S - 1 / 2 * dt / P.dx * Q1S)
    q2E, Q2E = sweep(q2, "E")
    q2S, Q2S = sweep(q2, "S")
    q3, _ = switch(q0E - dt / P.dx * Q2E, q0S - dt / P.dx * Q2S)
    q3E, Q3E = sweep(q3, "E")
    q3S, Q3S = sweep(q3, "S")
    q4, bad = switch(1 / 3 * (-q0E + q1E + 2 * q2E + q3E - 1 / 2 * dt / P.dx * Q3E),
                     1 / 3 * (-q0S + q1S + 2 * q2S + q3S - 1 / 2 * dt / P.dx * Q3S))
    return q4, bad


def compute_global_vmax_1d(q, P):
    """Weno5_1D.jl:1038-1074."""
    vx, _ = compute_global_vmax(q[:, None, :], P)
    return vx


# ----------------------------------------------------------------------------- reference ICs
def initial_conditions_1d(N, gam=5.0 / 3.0):
    """RJ95-2a shock tube, Weno5_1D.jl:1092-1113.  N includes ghosts."""
    s4p = np.sqrt(4.0 * np.pi)
    ql = state2conserved(1.08, 1.2, 0.01, 0.5, 2 / s4p, 3.6 / s4p, 2 / s4p, 0.95, gam)
    qr = state2conserved(1.0, 0.0, 0.0, 0.0, 2 / s4p, 4 / s4p, 2 / s4p, 1.0, gam)
    q0 = np.zeros((N, 9))
    for i1 in range(1, N + 1):
        q0[i1 - 1] = ql if i1 < N / 2 + 1 else qr
    return q0


def initial_conditions_2d(P):
    """Shock-cloud interaction, Weno5_2D.jl:2001-2041."""
    R0, R1 = 0.12, 0.19
    q = np.zeros((P.Nx, P.Ny, 9))
    sl = [2.11820, 13.69836, 0, 0, 0, 19.33647, 19.33647, 65.88816]
    sr = [1, 0, 0, 0, 0, 9.12871, 9.12871, 1.0]
    for j in range(1, P.Ny + 1):
        for i in range(1, P.Nx + 1):
            x = (i - 1.0 + 0.5) * P.dx - NB * P.dx
            y = (j - 1.0 + 0.5) * P.dy - NB * P.dy
            rr = np.sqrt((x - 1.5) ** 2 + (y - 0.5) ** 2)
            s = list(sl) if x <= 1.2 else list(sr)
            if rr <= R0:
                s[0] = 10
            if (rr > R0) and (rr <= R1):
                s[0] = 1 + 9 * (R1 - rr) / (R1 - R0)
            q[i - 1, j - 1] = state2conserved(*[float(v) for v in s], gam=P.gam)
    return q
