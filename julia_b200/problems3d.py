"""
Host-side initial conditions for the 3-D extension, in the public array layout continued to three
dimensions: ``q[Nx,Ny,Nz,9] = [rho,Mx,My,Mz,Bx,By,Bz,S,E]`` plus the face fields ``bxb`` (x-face i+1/2), ``byb``
(y-face j+1/2), ``bzb`` (z-face k+1/2), Fortran-ordered, NBnd = 3 (Weno5_2D.jl:9-17,35-37).

Face fields come from a vector potential sampled on the cell edges (div B = 0 to round-off); the cell-centred field
is derived from them with the reference's 4th-order stencil (Weno5_2D.jl:350-351) along each axis.
``embed_2d`` places a 2-D problem of ``problems.py`` into a volume that does not depend on the third coordinate,
in any of the three orientations -- the reduction that pins the 3-D step to the 2-D reference path.
"""
from dataclasses import dataclass
import numpy as np

from .problems import NBND, OUTFLOW, PERIODIC, state2conserved, _wrap_fill


@dataclass
class Grid3D:
    nx: int
    ny: int
    nz: int
    xsize: float = 1.0
    ysize: float = 1.0
    zsize: float = 1.0
    gam: float = 5.0 / 3.0
    cfl: float = 0.8
    eps: float = 1e-6
    bc_x: int = PERIODIC
    bc_y: int = PERIODIC
    bc_z: int = PERIODIC

    @property
    def Nx(self):
        return self.nx + 2 * NBND

    @property
    def Ny(self):
        return self.ny + 2 * NBND

    @property
    def Nz(self):
        return self.nz + 2 * NBND

    @property
    def dx(self):
        return self.xsize / self.nx

    @property
    def dy(self):
        return self.ysize / self.ny

    @property
    def dz(self):
        return self.zsize / self.nz

    @property
    def shape(self):
        return (self.Nx, self.Ny, self.Nz)

    def centres(self):
        """cell-centre coordinates along each axis (Weno5_2D.jl:2016 continued)."""
        return [(np.arange(N, dtype=np.float64) + 0.5 - NBND) * d
                for N, d in ((self.Nx, self.dx), (self.Ny, self.dy), (self.Nz, self.dz))]

    def faces(self):
        return [c + 0.5 * d for c, d in zip(self.centres(), (self.dx, self.dy, self.dz))]

    def bcs(self):
        return (self.bc_x, self.bc_y, self.bc_z)


def fill_ghosts(a, g):
    for ax, bc in enumerate(g.bcs()):
        _wrap_fill(a, ax, bc)


def fill_face_ghosts(a, g):
    """ghost layers of a face or flux array by the reference's rules for those (Weno5_2D.jl:1704-1729) along every
    axis: outflow copies index 3 / N-2 (1-based) outwards and zeroes the rim, periodic wraps."""
    for ax, bc in enumerate(g.bcs()):
        if bc == PERIODIC:
            _wrap_fill(a, ax, bc)
        else:
            v = np.moveaxis(a, ax, 0)
            N = v.shape[0]
            v[0:2] = v[2]
            v[N - 2:N] = v[N - 3]
    for ax, bc in enumerate(g.bcs()):
        if bc != PERIODIC:
            v = np.moveaxis(a, ax, 0)
            v[0] = 0.0
            v[-1] = 0.0


def face2center_3d(bxb, byb, bzb):
    """interpolateB_face2center (Weno5_2D.jl:344-356) along each axis, wherever the stencil fits."""
    out = []
    for ax, b in enumerate((bxb, byb, bzb)):
        v = np.moveaxis(b, ax, 0)
        N = v.shape[0]
        c = np.zeros_like(b, order="F")
        cv = np.moveaxis(c, ax, 0)
        cv[2:N - 1] = (-v[0:N - 3] + 9.0 * v[1:N - 2] + 9.0 * v[2:N - 1] - v[3:N]) / 16.0
        out.append(c)
    return out


def center2face_3d(q):
    """interpolateB_center2face (Weno5_2D.jl:1669-1682) with the z face field; zero where the reference leaves it."""
    Nx, Ny, Nz = q.shape[:3]
    out = []
    inner = (slice(1, Nx - 2), slice(1, Ny - 2), slice(1, Nz - 2))
    for ax in range(3):
        B = q[..., 4 + ax]
        f = np.zeros((Nx, Ny, Nz), order="F")

        def sh(o):
            s = list(inner)
            s[ax] = slice(inner[ax].start + o, inner[ax].stop + o)
            return tuple(s)
        f[inner] = (-B[sh(-1)] + 9 * B[sh(0)] + 9 * B[sh(1)] - B[sh(2)]) / 16
        out.append(f)
    return out


def assemble_3d(g, rho, vx, vy, vz, P, A=None, B_uniform=(0.0, 0.0, 0.0)):
    """(q, bxb, byb, bzb) from primitive fields on the full grid.  ``A = (Ax, Ay, Az)``: callables of (x, y, z)
    evaluated on the cell edges (Ax at (i, j+1/2, k+1/2), ...); uniform components are added on top."""
    shape = g.shape
    xc, yc, zc = g.centres()
    xf, yf, zf = g.faces()
    b = [np.full(shape, float(B_uniform[m]), order="F") for m in range(3)]
    if A is not None:
        def grid(x, y, z):
            return np.meshgrid(x, y, z, indexing="ij")
        Ax = A[0](*grid(xc, yf, zf)) if A[0] is not None else np.zeros(shape)
        Ay = A[1](*grid(xf, yc, zf)) if A[1] is not None else np.zeros(shape)
        Az = A[2](*grid(xf, yf, zc)) if A[2] is not None else np.zeros(shape)

        def d(a, ax, h):           # backward difference along ax (index 0 keeps the value of index 1)
            r = np.zeros(shape)
            v, rv = np.moveaxis(a, ax, 0), np.moveaxis(r, ax, 0)
            rv[1:] = (v[1:] - v[:-1]) / h
            rv[0] = rv[1]
            return r
        b[0] += d(Az, 1, g.dy) - d(Ay, 2, g.dz)
        b[1] += d(Ax, 2, g.dz) - d(Az, 0, g.dx)
        b[2] += d(Ay, 0, g.dx) - d(Ax, 1, g.dy)
    for f in b:
        fill_face_ghosts(f, g)
    Bc = face2center_3d(*b)
    for f in Bc:
        fill_ghosts(f, g)
    fields = state2conserved(rho, vx, vy, vz, Bc[0], Bc[1], Bc[2], P, g.gam)
    q = np.zeros(shape + (9,), order="F")
    for c, f in enumerate(fields):
        q[..., c] = f
        fill_ghosts(q[..., c], g)
    return q, b[0], b[1], b[2]


# ----------------------------------------------------------------------------- reduction to the 2-D reference path
_VEC = ([1, 2, 3], [4, 5, 6])          # momentum and field components of q


def embed_2d(g2, q, bxb, byb, orientation=0, n3=4, size3=1.0, bc3=PERIODIC):
    """Place the 2-D state (Nx,Ny,9) of grid ``g2`` into a volume that is uniform along the third coordinate.
    orientation 0: (x,y) plane, uniform in z; 1: the 2-D axes become (y,z), uniform in x; 2: they become (z,x).
    Vector components are permuted with the axes.  Returns (Grid3D, q, bxb, byb, bzb)."""
    o = orientation
    ax_a, ax_b, ax_c = o % 3, (o + 1) % 3, (o + 2) % 3       # 2-D x -> axis ax_a, 2-D y -> ax_b, uniform: ax_c
    n = [0, 0, 0]
    size = [0.0, 0.0, 0.0]
    bc = [0, 0, 0]
    n[ax_a], n[ax_b], n[ax_c] = g2.nx, g2.ny, n3
    size[ax_a], size[ax_b], size[ax_c] = g2.xsize, g2.ysize, size3
    bc[ax_a], bc[ax_b], bc[ax_c] = g2.bc_x, g2.bc_y, bc3
    g = Grid3D(nx=n[0], ny=n[1], nz=n[2], xsize=size[0], ysize=size[1], zsize=size[2], gam=g2.gam, cfl=g2.cfl,
               eps=g2.eps, bc_x=bc[0], bc_y=bc[1], bc_z=bc[2])
    N3 = n3 + 2 * NBND

    def lift(a2):
        v = np.repeat(a2[:, :, None], N3, axis=2)            # axes (a, b, c)
        return np.asfortranarray(np.moveaxis(v, (0, 1, 2), (ax_a, ax_b, ax_c)))

    q3 = np.zeros(g.shape + (9,), order="F")
    for c in (0, 7, 8):
        q3[..., c] = lift(q[:, :, c])
    for base in _VEC:
        # 2-D components (x,y,z) live along axes (ax_a, ax_b, ax_c)
        for m2, ax in enumerate((ax_a, ax_b, ax_c)):
            q3[..., base[ax]] = lift(q[:, :, base[m2]])
    faces = [None, None, None]
    faces[ax_a] = lift(bxb)
    faces[ax_b] = lift(byb)
    faces[ax_c] = lift(q[:, :, 6])                            # uniform along ax_c: the face value is the cell value
    return g, q3, faces[0], faces[1], faces[2]


def extract_2d(a3, orientation=0, index=NBND):
    """inverse of the ``lift`` of embed_2d for one scalar volume: the (a, b) plane at position ``index`` of the
    uniform axis."""
    o = orientation
    ax_a, ax_b, ax_c = o % 3, (o + 1) % 3, (o + 2) % 3
    v = np.moveaxis(a3, (ax_a, ax_b, ax_c), (0, 1, 2))
    return np.asfortranarray(v[:, :, index])


def vector_component(orientation, m2):
    """index 0..2 of the 3-D vector component that carries the 2-D component m2 (0 x, 1 y, 2 z)."""
    return ((orientation % 3), ((orientation + 1) % 3), ((orientation + 2) % 3))[m2]


# ----------------------------------------------------------------------------- genuinely 3-D problems
def orszag_tang_3d(n=64, nz=None, eps_z=0.2, gam=5.0 / 3.0):
    """Orszag-Tang vortex with a z-modulated velocity (Helzel, Rossmanith & Taetz 2011), periodic unit cube:
    rho = 25/(36 pi), P = 5/(12 pi), v = (-(1 + e sin 2 pi z) sin 2 pi y, (1 + e sin 2 pi z) sin 2 pi x, e sin 2 pi z),
    B = (-B0 sin 2 pi y, B0 sin 4 pi x, 0), B0 = 1/sqrt(4 pi), from Az = B0 (cos 4 pi x / 4 pi + cos 2 pi y / 2 pi)."""
    nz = n if nz is None else nz
    g = Grid3D(nx=n, ny=n, nz=nz, gam=gam)
    X, Y, Z = np.meshgrid(*g.centres(), indexing="ij")
    B0 = 1.0 / np.sqrt(4.0 * np.pi)
    rho = np.full_like(X, 25.0 / (36.0 * np.pi))
    P = np.full_like(X, 5.0 / (12.0 * np.pi))
    mod = 1.0 + eps_z * np.sin(2 * np.pi * Z)
    vx, vy, vz = -mod * np.sin(2 * np.pi * Y), mod * np.sin(2 * np.pi * X), eps_z * np.sin(2 * np.pi * Z)
    Az = lambda x, y, z: B0 * (np.cos(4 * np.pi * x) / (4 * np.pi) + np.cos(2 * np.pi * y) / (2 * np.pi))
    return (g,) + assemble_3d(g, rho, vx, vy, vz, P, (None, None, Az))


def blast_3d(n=64, bc=PERIODIC, gam=5.0 / 3.0):
    """MHD blast wave in the unit cube: rho = 1, P = 0.1 (10 for r < 0.1 around the centre), B = (1,1,0)/sqrt 2."""
    g = Grid3D(nx=n, ny=n, nz=n, gam=gam, bc_x=bc, bc_y=bc, bc_z=bc)
    X, Y, Z = np.meshgrid(*g.centres(), indexing="ij")
    r = np.sqrt((X - 0.5) ** 2 + (Y - 0.5) ** 2 + (Z - 0.5) ** 2)
    P = np.where(r < 0.1, 10.0, 0.1)
    one, z = np.ones_like(X), np.zeros_like(X)
    b = 1.0 / np.sqrt(2.0)
    return (g,) + assemble_3d(g, one, z, z, z, P, None, (b, b, 0.0))


def alfven_wave_3d(n=32, amp=0.1, gam=5.0 / 3.0):
    """Circularly polarised Alfven wave (an exact nonlinear solution) along the diagonal of the periodic unit cube:
    rho = 1, P = 0.1, B_par = 1 along e1 = (1,1,1)/sqrt 3, transverse field and velocity amp (sin phi e2 + cos phi e3),
    phi = 2 pi (x + y + z); wavelength 1/sqrt 3, Alfven speed 1, period 1/sqrt 3.  The 3-D counterpart of alfven_wave
    (BASELINE config 2); face fields from the vector potential amp/kappa (sin phi e2 + cos phi e3), kappa = 2 pi sqrt 3."""
    g = Grid3D(nx=n, ny=n, nz=n, gam=gam)
    e1 = np.array([1.0, 1.0, 1.0]) / np.sqrt(3.0)
    e2 = np.array([1.0, -1.0, 0.0]) / np.sqrt(2.0)
    e3 = np.cross(e1, e2)
    kap = 2 * np.pi * np.sqrt(3.0)

    def ph(x, y, z):
        return 2 * np.pi * (x + y + z)
    X, Y, Z = np.meshgrid(*g.centres(), indexing="ij")
    s, c = np.sin(ph(X, Y, Z)), np.cos(ph(X, Y, Z))
    v = [amp * (s * e2[m] + c * e3[m]) for m in range(3)]
    A = [(lambda x, y, z, m=m: amp / kap * (np.sin(ph(x, y, z)) * e2[m] + np.cos(ph(x, y, z)) * e3[m])) for m in range(3)]
    return (g,) + assemble_3d(g, np.ones_like(X), v[0], v[1], v[2], np.full_like(X, 0.1), A, tuple(e1))


def smooth_3d(nx=24, ny=20, nz=16, seed=7, bc=PERIODIC, gam=5.0 / 3.0):
    """A smooth state with every component of v and B depending on all three coordinates (low Fourier modes),
    rho in [0.6,1.6], P in [0.4,1.4]; the kernel-level input of the 3-D parity tests."""
    g = Grid3D(nx=nx, ny=ny, nz=nz, gam=gam, bc_x=bc, bc_y=bc, bc_z=bc)
    rng = np.random.default_rng(seed)
    X, Y, Z = np.meshgrid(*g.centres(), indexing="ij")

    def mode(x, y, z):
        f = np.zeros(np.broadcast(x, y, z).shape)
        for _ in range(3):
            k = rng.integers(0, 3, size=3)
            ph = rng.uniform(0, 2 * np.pi)
            f = f + rng.uniform(-1, 1) * np.cos(2 * np.pi * (k[0] * x + k[1] * y + k[2] * z) + ph)
        return f

    def unit(f):
        return (f - f.min()) / max(f.max() - f.min(), 1e-300)

    rho = 0.6 + unit(mode(X, Y, Z))
    P = 0.4 + unit(mode(X, Y, Z))
    vx, vy, vz = (0.5 * mode(X, Y, Z) for _ in range(3))
    coef = [[(rng.integers(1, 3, size=3), rng.uniform(0, 2 * np.pi), rng.uniform(-0.06, 0.06)) for _ in range(2)]
            for _ in range(3)]

    def pot(m):
        def f(x, y, z):
            s = 0.0
            for k, ph, amp in coef[m]:
                s = s + amp * np.cos(2 * np.pi * (k[0] * x + k[1] * y + k[2] * z) + ph)
            return s
        return f
    return (g,) + assemble_3d(g, rho, vx, vy, vz, P, (pot(0), pot(1), pot(2)), (0.3, -0.2, 0.25))


def rotate_problem(g, q, bxb, byb, bzb):
    """The same physical problem with the axes renamed cyclically (x,y,z) -> (y,z,x): what was along x now runs along y.
    A solver that treats the three directions alike must return the renamed result."""
    g2 = Grid3D(nx=g.nz, ny=g.nx, nz=g.ny, xsize=g.zsize, ysize=g.xsize, zsize=g.ysize, gam=g.gam, cfl=g.cfl, eps=g.eps,
                bc_x=g.bc_z, bc_y=g.bc_x, bc_z=g.bc_y)

    def mv(a):
        return np.asfortranarray(np.moveaxis(a, (0, 1, 2), (1, 2, 0)))
    q2 = np.zeros(g2.shape + (9,), order="F")
    for c in (0, 7, 8):
        q2[..., c] = mv(q[..., c])
    for base in _VEC:
        q2[..., base[1]] = mv(q[..., base[0]])
        q2[..., base[2]] = mv(q[..., base[1]])
        q2[..., base[0]] = mv(q[..., base[2]])
    return g2, q2, mv(bzb), mv(bxb), mv(byb)
