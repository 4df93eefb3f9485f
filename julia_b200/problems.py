"""
Host-side initial conditions in the reference's public array layout.

The reference builds its state on the host (``initial_conditions`` /
``state2conserved``, Weno5_2D.jl:2001-2053, Weno5_1D.jl:1092-1124) and hands
``q[Nx,Ny,9] = [rho,Mx,My,Mz,Bx,By,Bz,S,E]`` plus the face fields ``bxb``/``byb`` to the
time-step call.  This module does the same for the two problems the reference ships
(RJ95-2a shock tube, shock-cloud) and for the problems named in BASELINE.json
(Brio-Wu, circularly polarised Alfven wave, Orszag-Tang, blast), which reuse the
reference numerics unchanged.

All 2-D arrays are Fortran-ordered numpy arrays so that their memory is exactly what
Julia's ``Array{Float64,3}(Nx,Ny,9)`` holds (i fastest).  NBnd = 3.

For problems with a non-trivial in-plane field the face fields come from a vector
potential (div B = 0 to round-off) and the cell-centred Bx,By are derived from them with
the reference's own 4th-order stencil (Weno5_2D.jl:350-351), cf. SURVEY.md Q14.
"""
from dataclasses import dataclass
import numpy as np

NBND = 3
OUTFLOW, PERIODIC = 0, 1


@dataclass
class Grid:
    nx: int
    ny: int = 1
    xsize: float = 1.0
    ysize: float = 1.0
    gam: float = 5.0 / 3.0
    cfl: float = 0.8
    eps: float = 1e-6
    bc_x: int = OUTFLOW
    bc_y: int = OUTFLOW

    @property
    def Nx(self):
        return self.nx + 2 * NBND

    @property
    def Ny(self):
        return self.ny + 2 * NBND

    @property
    def dx(self):
        return self.xsize / self.nx

    @property
    def dy(self):
        return self.ysize / self.ny

    def xc(self):
        """cell-centre x of public index i=1..Nx (Weno5_2D.jl:2016)."""
        i = np.arange(1, self.Nx + 1, dtype=np.float64)
        return (i - 1.0 + 0.5) * self.dx - NBND * self.dx

    def yc(self):
        j = np.arange(1, self.Ny + 1, dtype=np.float64)
        return (j - 1.0 + 0.5) * self.dy - NBND * self.dy

    def xf(self):
        """x of the face i+1/2 for public index i=1..Nx."""
        return self.xc() + 0.5 * self.dx

    def yf(self):
        return self.yc() + 0.5 * self.dy


def state2conserved(rho, vx, vy, vz, Bx, By, Bz, P, gam=5.0 / 3.0):
    """Weno5_2D.jl:2043-2053 -- works on scalars or arrays; returns a list of 9 fields."""
    v2 = vx ** 2 + vy ** 2 + vz ** 2
    B2 = Bx ** 2 + By ** 2 + Bz ** 2
    E = P / (gam - 1) + 0.5 * (rho * v2 + B2)
    S = P / rho ** (gam - 1)
    return [rho, rho * vx, rho * vy, rho * vz, Bx, By, Bz, S, E]


def _wrap_fill(a, axis, bc):
    N = a.shape[axis]
    n = N - 2 * NBND
    v = np.moveaxis(a, axis, 0)
    if bc == PERIODIC:
        v[0:NBND] = v[n:n + NBND]
        v[N - NBND:N] = v[NBND:2 * NBND]
    else:
        v[0:NBND] = v[NBND]
        v[N - NBND:N] = v[N - NBND - 1]


def face2center(bxb, byb):
    """Weno5_2D.jl:344-356 evaluated wherever the stencil fits."""
    Nx, Ny = bxb.shape
    Bx = np.zeros((Nx, Ny), order="F")
    By = np.zeros((Nx, Ny), order="F")
    Bx[2:Nx - 1, :] = (-bxb[0:Nx - 3, :] + 9.0 * bxb[1:Nx - 2, :] + 9.0 * bxb[2:Nx - 1, :] - bxb[3:Nx, :]) / 16.0
    By[:, 2:Ny - 1] = (-byb[:, 0:Ny - 3] + 9.0 * byb[:, 1:Ny - 2] + 9.0 * byb[:, 2:Ny - 1] - byb[:, 3:Ny]) / 16.0
    return Bx, By


def center2face(q):
    """Weno5_2D.jl:1669-1682 (interpolateB_center2face)."""
    Nx, Ny = q.shape[0], q.shape[1]
    bxb = np.zeros((Nx, Ny), order="F")
    byb = np.zeros((Nx, Ny), order="F")
    i, j = slice(1, Nx - 2), slice(1, Ny - 2)
    bxb[i, j] = (-q[0:Nx - 3, j, 4] + 9 * q[i, j, 4] + 9 * q[2:Nx - 1, j, 4] - q[3:Nx, j, 4]) / 16
    byb[i, j] = (-q[i, 0:Ny - 3, 5] + 9 * q[i, j, 5] + 9 * q[i, 2:Ny - 1, 5] - q[i, 3:Ny, 5]) / 16
    return bxb, byb


def _assemble(g, rho, vx, vy, vz, Bz, P, Az_corner=None, Bx_uniform=0.0, By_uniform=0.0):
    """Build (q, bxb, byb) from primitive fields on the full (Nx,Ny) grid.

    Az_corner(xf, yf) gives the vector potential at cell corners (i+1/2, j+1/2); uniform
    in-plane components are added on top.
    """
    Nx, Ny = g.Nx, g.Ny
    shape = (Nx, Ny)
    bxb = np.full(shape, Bx_uniform, order="F")
    byb = np.full(shape, By_uniform, order="F")
    if Az_corner is not None:
        XF, YF = np.meshgrid(g.xf(), g.yf(), indexing="ij")
        Az = Az_corner(XF, YF)                       # Az[i,j] at corner (i+1/2, j+1/2)
        bxb[:, 1:] += (Az[:, 1:] - Az[:, :-1]) / g.dy         # Bx =  dAz/dy on x-face i+1/2
        byb[1:, :] += -(Az[1:, :] - Az[:-1, :]) / g.dx        # By = -dAz/dx on y-face j+1/2
        bxb[:, 0] = bxb[:, 1]
        byb[0, :] = byb[1, :]
    for f in (bxb, byb):
        _wrap_fill(f, 0, g.bc_x)
        _wrap_fill(f, 1, g.bc_y)
    Bx, By = face2center(bxb, byb)
    for f in (Bx, By):
        _wrap_fill(f, 0, g.bc_x)
        _wrap_fill(f, 1, g.bc_y)
    fields = state2conserved(rho, vx, vy, vz, Bx, By, Bz, P, g.gam)
    q = np.zeros((Nx, Ny, 9), order="F")
    for c, f in enumerate(fields):
        q[:, :, c] = f
    # ghost cells are exact copies of the cells the boundary condition maps them to
    # (what boundaries! would produce), not the formula evaluated outside the domain
    for c in range(9):
        _wrap_fill(q[:, :, c], 0, g.bc_x)
        _wrap_fill(q[:, :, c], 1, g.bc_y)
    return q, bxb, byb


# ----------------------------------------------------------------------------- reference problems
def rj95_2a(nx=1024, gam=5.0 / 3.0):
    """1-D RJ95-2a shock tube exactly as Weno5_1D.jl:1092-1113.  Returns (grid, q[N,9])."""
    g = Grid(nx=nx, ny=1, xsize=1.0, gam=gam, bc_x=OUTFLOW)
    N = g.Nx
    s4p = np.sqrt(4.0 * np.pi)
    ql = state2conserved(1.08, 1.2, 0.01, 0.5, 2 / s4p, 3.6 / s4p, 2 / s4p, 0.95, gam)
    qr = state2conserved(1.0, 0.0, 0.0, 0.0, 2 / s4p, 4 / s4p, 2 / s4p, 1.0, gam)
    q = np.zeros((N, 9), order="F")
    for i in range(1, N + 1):
        q[i - 1, :] = ql if i < N / 2 + 1 else qr
    return g, q


def brio_wu(nx=1024, gam=2.0):
    """Classic Brio & Wu (1988) tube: (rho,P,By) = (1,1,1 | 0.125,0.1,-1), Bx = 0.75."""
    g = Grid(nx=nx, ny=1, xsize=1.0, gam=gam, bc_x=OUTFLOW)
    N = g.Nx
    ql = state2conserved(1.0, 0.0, 0.0, 0.0, 0.75, 1.0, 0.0, 1.0, gam)
    qr = state2conserved(0.125, 0.0, 0.0, 0.0, 0.75, -1.0, 0.0, 0.1, gam)
    q = np.zeros((N, 9), order="F")
    for i in range(1, N + 1):
        q[i - 1, :] = ql if i < N / 2 + 1 else qr
    return g, q


def shock_cloud(nx=16, ny=4, xsize=4.0, ysize=1.0, gam=5.0 / 3.0):
    """2-D shock-cloud interaction, Weno5_2D.jl:2001-2041, with the reference's start-up
    sequence initial_conditions -> interpolateB_center2face (driver :35-37)."""
    g = Grid(nx=nx, ny=ny, xsize=xsize, ysize=ysize, gam=gam, bc_x=OUTFLOW, bc_y=OUTFLOW)
    R0, R1 = 0.12, 0.19
    X, Y = np.meshgrid(g.xc(), g.yc(), indexing="ij")
    rr = np.sqrt((X - 1.5) ** 2 + (Y - 0.5) ** 2)
    left = X <= 1.2
    rho = np.where(left, 2.11820, 1.0)
    rho = np.where(rr <= R0, 10.0, rho)
    ring = (rr > R0) & (rr <= R1)
    rho = np.where(ring, 1 + 9 * (R1 - rr) / (R1 - R0), rho)
    vx = np.where(left, 13.69836, 0.0)
    zero = np.zeros_like(X)
    By = np.where(left, 19.33647, 9.12871)
    Bz = np.where(left, 19.33647, 9.12871)
    P = np.where(left, 65.88816, 1.0)
    fields = state2conserved(rho, vx, zero, zero, zero, By, Bz, P, gam)
    q = np.zeros((g.Nx, g.Ny, 9), order="F")
    for c, f in enumerate(fields):
        q[:, :, c] = f
    bxb, byb = center2face(q)
    return g, q, bxb, byb


# ----------------------------------------------------------------------------- BASELINE.json problems
def orszag_tang(n=2048, ny=None, ysize=1.0, y0=0.0, gam=5.0 / 3.0):
    """Orszag-Tang vortex on [0,1]x[y0, y0+ysize], periodic.  ``ny``/``ysize``/``y0`` allow
    the slab of a y-decomposed (or weak-scaled, ysize = #ranks) domain."""
    ny = n if ny is None else ny
    g = Grid(nx=n, ny=ny, xsize=1.0, ysize=ysize, gam=gam, bc_x=PERIODIC, bc_y=PERIODIC)
    X, Y = np.meshgrid(g.xc(), g.yc() + y0, indexing="ij")
    B0 = 1.0 / np.sqrt(4.0 * np.pi)
    rho = np.full_like(X, 25.0 / (36.0 * np.pi))
    P = np.full_like(X, 5.0 / (12.0 * np.pi))
    vx, vy, z = -np.sin(2 * np.pi * Y), np.sin(2 * np.pi * X), np.zeros_like(X)
    Az = lambda xf, yf: B0 * (np.cos(4 * np.pi * xf) / (4 * np.pi) + np.cos(2 * np.pi * (yf + y0)) / (2 * np.pi))
    q, bxb, byb = _assemble(g, rho, vx, vy, z, z, P, Az)
    return g, q, bxb, byb


def alfven_wave(n=512, angle_deg=30.0, amp=0.1, gam=5.0 / 3.0):
    """Circularly polarised Alfven wave (Toth 2000) at ``angle_deg`` to the x axis,
    domain [0,1/cos a]x[0,1/sin a], periodic; smooth problem for the 1e-10 parity check."""
    al = np.deg2rad(angle_deg)
    ca, sa = np.cos(al), np.sin(al)
    g = Grid(nx=n, ny=n, xsize=1.0 / ca, ysize=1.0 / sa, gam=gam, bc_x=PERIODIC, bc_y=PERIODIC)
    X, Y = np.meshgrid(g.xc(), g.yc(), indexing="ij")
    xpar = X * ca + Y * sa
    vperp = amp * np.sin(2 * np.pi * xpar)
    vz = amp * np.cos(2 * np.pi * xpar)
    rho = np.ones_like(X)
    P = np.full_like(X, 0.1)
    vx, vy = -vperp * sa, vperp * ca
    Az = lambda xf, yf: amp * np.cos(2 * np.pi * (xf * ca + yf * sa)) / (2 * np.pi)
    q, bxb, byb = _assemble(g, rho, vx, vy, vz, vz.copy(), P, Az, Bx_uniform=ca, By_uniform=sa)
    return g, q, bxb, byb


def blast(n=512, ny=None, ysize=1.0, y0=0.0, gam=5.0 / 3.0):
    """MHD blast wave on [0,1]x[y0,y0+ysize], periodic: rho=1, P=0.1 (10 for r<0.1 around
    (0.5,0.5)), B=(1,1,0)/sqrt(2), v=0."""
    ny = n if ny is None else ny
    g = Grid(nx=n, ny=ny, xsize=1.0, ysize=ysize, gam=gam, bc_x=PERIODIC, bc_y=PERIODIC)
    X, Y = np.meshgrid(g.xc(), g.yc() + y0, indexing="ij")
    r = np.sqrt((X - 0.5) ** 2 + (Y - 0.5) ** 2)
    P = np.where(r < 0.1, 10.0, 0.1)
    one, z = np.ones_like(X), np.zeros_like(X)
    b = 1.0 / np.sqrt(2.0)
    q, bxb, byb = _assemble(g, one, z, z, z, z, P, None, Bx_uniform=b, By_uniform=b)
    return g, q, bxb, byb


def random_state(nx, ny, seed=1234, bc=PERIODIC, smooth=True, gam=5.0 / 3.0):
    """Kernel-level unit input: rho in [0.5,2], v,B in [-1,1]^3, P in [0.1,2].
    With ``smooth`` the random field is a few low Fourier modes (so WENO weights stay sane)."""
    g = Grid(nx=nx, ny=ny, gam=gam, bc_x=bc, bc_y=bc)
    rng = np.random.default_rng(seed)
    X, Y = np.meshgrid(g.xc(), g.yc(), indexing="ij")

    def field(lo, hi):
        if not smooth:
            return rng.uniform(lo, hi, size=X.shape)
        f = np.zeros_like(X)
        for _ in range(4):
            kx, ky = rng.integers(0, 3, size=2)
            ph = rng.uniform(0, 2 * np.pi)
            f += rng.uniform(-1, 1) * np.cos(2 * np.pi * (kx * X + ky * Y) + ph)
        f = (f - f.min()) / max(f.max() - f.min(), 1e-300)
        return lo + (hi - lo) * f

    rho, P = field(0.5, 2.0), field(0.1, 2.0)
    vx, vy, vz, Bz = field(-1, 1), field(-1, 1), field(-1, 1), field(-1, 1)
    XF, YF = np.meshgrid(g.xf(), g.yf(), indexing="ij")
    kx, ky = 1, 2
    Az = lambda xf, yf: 0.15 * np.cos(2 * np.pi * kx * xf) * np.sin(2 * np.pi * ky * yf) + 0.1 * np.sin(2 * np.pi * xf)
    q, bxb, byb = _assemble(g, rho, vx, vy, vz, Bz, P, Az, Bx_uniform=0.3, By_uniform=-0.2)
    for c in range(9):
        _wrap_fill(q[:, :, c], 0, bc)
        _wrap_fill(q[:, :, c], 1, bc)
    return g, q, bxb, byb
