"""
Host-side mirror of the reference's operator interface for the WENO5 MHD path.

The reference exposes plain Julia functions; this module offers the same
names, argument meaning and array layout in Python on top of the C ABI, so that the parity
tests read like a driver of the reference:

    q = initial_conditions(); bxb, byb = interpolateB_center2face(q); boundaries!(q,bxb,byb)
    dt, vxmax, vymax = timestep(q)
    q, bxb, byb = classical_RK4_step(q, bxb, byb, dt)
    divb = compute_divB(bxb, byb)

Two styles:
  * ``Solver`` keeps the state resident in HBM (upload once, ``advance`` many steps) -- the
    B200-native way to drive the path;
  * the module-level functions take and return host arrays exactly like the reference
    (every call uploads and downloads).
All compute happens in libweno5mhd.so; nothing here falls back to the CPU.
"""
import ctypes as C
import numpy as np

from . import lib as L
from .problems import Grid, NBND


def _farr(a, shape=None):
    a = np.asfortranarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"expected array of shape {tuple(shape)}, got {a.shape}")
    return a


def make_params(g, es_roundtrip=True):
    return L.Params(g.nx, g.ny, g.dx, g.dy if g.ny > 1 else 1.0, g.gam, g.cfl, g.eps,
                    int(g.bc_x), int(g.bc_y), 1 if es_roundtrip else 0)


class Solver:
    """One GPU, one y-slab of the grid ``g`` (the whole grid by default)."""

    def __init__(self, g, device=0, ny_local=None, j_offset=0, es_roundtrip=True):
        self.lib = L.load()
        self.g = g
        self.is1d = g.ny == 1
        self.ny_local = g.ny if ny_local is None else ny_local
        self.j_offset = j_offset
        self.Nx = g.nx + 2 * NBND
        self.Ny = 1 if self.is1d else self.ny_local + 2 * NBND
        self.params = make_params(g, es_roundtrip)
        h = C.c_void_p()
        L.check(self.lib.weno5_create(C.byref(self.params), device, self.ny_local, j_offset, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.weno5_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- multi-GPU
    @staticmethod
    def comm_unique_id():
        buf = (C.c_ubyte * 128)()
        L.check(L.load().weno5_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, rank, nranks, uid):
        buf = (C.c_ubyte * 128).from_buffer_copy(uid)
        L.check(self.lib.weno5_comm_init(self.h, rank, nranks, buf))

    def halo_mode(self):
        """0: no communicator, 1: NCCL send/recv + NCCL all-reduce, 2: direct stores into the neighbours' ghost rows
        (peer memory), 3: that and the dt all-reduce over peer memory."""
        return int(self.lib.weno5_comm_halo_mode(self.h))

    # -- state
    def upload(self, q, bxb=None, byb=None):
        if self.is1d:
            q = _farr(q, (self.Nx, 9))
            L.check(self.lib.weno5_upload_1d(self.h, L.dptr(q)))
            return
        q = _farr(q, (self.Nx, self.Ny, 9))
        if bxb is not None:
            bxb, byb = _farr(bxb, (self.Nx, self.Ny)), _farr(byb, (self.Nx, self.Ny))
        L.check(self.lib.weno5_upload(self.h, L.dptr(q), L.dptr(bxb), L.dptr(byb)))

    def download(self, faces=True):
        if self.is1d:
            q = np.zeros((self.Nx, 9), order="F")
            L.check(self.lib.weno5_download_1d(self.h, L.dptr(q)))
            return q
        q = np.zeros((self.Nx, self.Ny, 9), order="F")
        bxb = np.zeros((self.Nx, self.Ny), order="F") if faces else None
        byb = np.zeros((self.Nx, self.Ny), order="F") if faces else None
        L.check(self.lib.weno5_download(self.h, L.dptr(q), L.dptr(bxb), L.dptr(byb)))
        return (q, bxb, byb) if faces else q

    # -- the hot path
    def boundaries(self):
        L.check(self.lib.weno5_boundaries(self.h))

    def timestep(self):
        dt, vx, vy = C.c_double(), C.c_double(), C.c_double()
        L.check(self.lib.weno5_timestep(self.h, C.byref(dt), C.byref(vx), C.byref(vy)))
        return dt.value, vx.value, vy.value

    def rk4_step(self, dt):
        L.check(self.lib.weno5_rk4_step(self.h, float(dt)))

    def advance(self, nsteps, tmax=0.0, t=0.0):
        tt, nd = C.c_double(t), C.c_int(0)
        L.check(self.lib.weno5_advance(self.h, int(nsteps), float(tmax), C.byref(tt), C.byref(nd)))
        return tt.value, nd.value

    def step_host(self, q, bxb, byb, dt, out=None):
        """classical_RK4_step through the handle: host arrays in, host arrays out."""
        if self.is1d:
            q = _farr(q, (self.Nx, 9))
            q4 = np.zeros_like(q, order="F") if out is None else out
            L.check(self.lib.weno5_step_host(self.h, L.dptr(q), None, None, float(dt), L.dptr(q4), None, None))
            return q4
        q, bxb, byb = _farr(q), _farr(bxb), _farr(byb)
        if out is None:
            out = (np.zeros_like(q, order="F"), np.zeros_like(bxb, order="F"), np.zeros_like(byb, order="F"))
        L.check(self.lib.weno5_step_host(self.h, L.dptr(q), L.dptr(bxb), L.dptr(byb), float(dt),
                                         L.dptr(out[0]), L.dptr(out[1]), L.dptr(out[2])))
        return out

    # -- dual-energy mode
    def set_es_switch(self, enable=True, threshold=1e-5):
        """Run both the E and the S sweeps every stage and let ES_Switch (Weno5_2D.jl:1734) choose per cell
        with the criterion the reference wrote and disabled (:1740): |S(E) - S| >= threshold -> E state.
        enable = 2 (1-D grids): the 1-D reference as shipped -- energy state everywhere, the flags are the output."""
        L.check(self.lib.weno5_set_es_switch(self.h, int(enable) if enable in (0, 1, 2) else (1 if enable else 0), float(threshold)))

    def es_switch_mask(self):
        """1.0 where |S(E) - S| >= threshold in the last stage (the reference's `idx`, Weno5_1D.jl:245, as a mask)."""
        m = np.zeros((self.Nx,) if self.is1d else (self.Nx, self.Ny), order="F")
        L.check(self.lib.weno5_es_switch_mask(self.h, L.dptr(m)))
        return m

    def es_switch_count(self):
        n = C.c_ulonglong()
        L.check(self.lib.weno5_es_switch_count(self.h, C.byref(n)))
        return int(n.value)

    # -- snapshot hook
    def arr2img(self, component, cmap=None, range=(0.0, 0.0), log=False, ncolours=None):
        """``Arr2Img(q[:,:,component+1], cmap; range, log)`` (JColorMaps.jl:92) on the resident state: returns
        ``cmap[rimg]`` (shape (Ny, Nx) + cmap.shape[1:]) -- or the 1-based Int16 colour indices if ``cmap`` is
        None -- and the range used.  Components 9, 10, 11 are bxb, byb and compute_divB."""
        n = int(ncolours if cmap is None else len(cmap))
        rimg = np.zeros((self.Ny, self.Nx), dtype=np.int16, order="F")
        rng = np.zeros(2)
        L.check(self.lib.weno5_arr2img(self.h, int(component), float(range[0]), float(range[1]), 1 if log else 0, n,
                                       rimg.ctypes.data_as(C.POINTER(C.c_short)), L.dptr(rng)))
        if cmap is None:
            return rimg, tuple(rng)
        return np.asarray(cmap)[rimg.astype(np.int64) - 1], tuple(rng)

    # -- diagnostics
    def divb(self):
        d = np.zeros((self.Nx, self.Ny), order="F")
        L.check(self.lib.weno5_divb(self.h, L.dptr(d)))
        return d

    def totals(self):
        s = np.zeros(11)
        L.check(self.lib.weno5_totals(self.h, L.dptr(s)))
        return s

    def launch_count(self):
        return int(self.lib.weno5_launch_count(self.h))

    def profile(self, enable=True):
        L.check(self.lib.weno5_profile(self.h, 1 if enable else 0))

    def profile_read(self, reset=True):
        ms, n = C.c_double(), C.c_ulonglong()
        L.check(self.lib.weno5_profile_read(self.h, C.byref(ms), C.byref(n), 1 if reset else 0))
        return ms.value, int(n.value)

    def debug_fetch(self, name):
        d = np.zeros((self.Nx, self.Ny), order="F")
        L.check(self.lib.weno5_debug_fetch(self.h, name.encode(), L.dptr(d)))
        return d


class Pipe:
    """Pipelined host -> host classical_RK4_step: y-slabs with overlap rows stream through the GPU so
    that upload, compute and download overlap (weno5_pipe_* in include/weno5mhd.h)."""

    def __init__(self, g, device=0, rows_per_slab=128, es_roundtrip=True):
        self.lib = L.load()
        self.g = g
        self.params = make_params(g, es_roundtrip)
        h = C.c_void_p()
        L.check(self.lib.weno5_pipe_create(C.byref(self.params), device, int(rows_per_slab), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.weno5_pipe_destroy(self.h)
            self.h = None
            for p in getattr(self, "_pinned", []):
                self.lib.weno5_host_free(p)
            self._pinned = []

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def pinned(self, shape):
        """Page-locked Fortran-ordered float64 array (weno5_host_alloc): pass such arrays as inputs and `out` to let
        the copies overlap the compute.  Freed with the Pipe."""
        n = int(np.prod(shape))
        p = C.c_void_p()
        L.check(self.lib.weno5_host_alloc(C.byref(p), n * 8))
        self._pinned = getattr(self, "_pinned", []) + [p]
        buf = (C.c_double * n).from_address(p.value)
        return np.frombuffer(buf, dtype=np.float64).reshape(shape, order="F")

    def overlapped(self):
        """True if the last step() was given page-locked arrays (only then do copies and compute overlap)."""
        return bool(self.lib.weno5_pipe_overlapped(self.h))

    def step(self, q, bxb, byb, dt, out=None):
        """(q4, bxb4, byb4, dt_next); `out` = preallocated (q4, bxb4, byb4), no byte of which may lie in an input.
        Pageable arrays (the default `out`) work but serialise the copies: see pinned() / overlapped()."""
        g = self.g
        q, bxb, byb = _farr(q, (g.Nx, g.Ny, 9)), _farr(bxb, (g.Nx, g.Ny)), _farr(byb, (g.Nx, g.Ny))
        if out is None:
            out = (np.zeros_like(q, order="F"), np.zeros_like(bxb, order="F"), np.zeros_like(byb, order="F"))
        dtn = C.c_double()
        L.check(self.lib.weno5_pipe_step(self.h, L.dptr(q), L.dptr(bxb), L.dptr(byb), float(dt),
                                         L.dptr(out[0]), L.dptr(out[1]), L.dptr(out[2]), C.byref(dtn)))
        return out[0], out[1], out[2], dtn.value


# ----------------------------------------------------------------------------- reference-shaped API
def classical_RK4_step(g, q, bxb=None, byb=None, dt=None, es_roundtrip=True):
    """Weno5_2D.jl:125 ``classical_RK4_step(q,bxb,byb,dt) -> (q4,bxb4,byb4)``;
    1-D (Weno5_1D.jl:152): ``classical_RK4_step(g, q, dt=dt) -> q4``."""
    lib = L.load()
    P = make_params(g, es_roundtrip)
    if g.ny == 1:
        q = _farr(q, (g.Nx, 9))
        q4 = np.zeros_like(q, order="F")
        L.check(lib.weno5_classical_rk4_step_1d(C.byref(P), L.dptr(q), float(dt), L.dptr(q4)))
        return q4
    q, bxb, byb = _farr(q, (g.Nx, g.Ny, 9)), _farr(bxb, (g.Nx, g.Ny)), _farr(byb, (g.Nx, g.Ny))
    q4, bx4, by4 = np.zeros_like(q, order="F"), np.zeros_like(bxb, order="F"), np.zeros_like(byb, order="F")
    L.check(lib.weno5_classical_rk4_step_2d(C.byref(P), L.dptr(q), L.dptr(bxb), L.dptr(byb), float(dt),
                                            L.dptr(q4), L.dptr(bx4), L.dptr(by4)))
    return q4, bx4, by4


def timestep(g, q):
    """Weno5_2D.jl:1930 ``timestep(q) -> (dt, vxmax, vymax)`` (1-D: dt = courFac*dx/vmax)."""
    with Solver(g) as s:
        if g.ny == 1:
            s.upload(q)
        else:
            # the face fields do not enter the time step
            z = np.zeros((g.Nx, g.Ny), order="F")
            s.upload(q, z, z)
        return s.timestep()


def compute_global_vmax(g, q):
    """Weno5_2D.jl:1942 / Weno5_1D.jl:1038."""
    dt, vx, vy = timestep(g, q)
    return vx if g.ny == 1 else (vx, vy)


def interpolateB_center2face(g, q):
    """Weno5_2D.jl:1669 ``interpolateB_center2face(q) -> (bxb, byb)`` followed by the
    driver's boundaries!(q,bxb,byb) (:39)."""
    with Solver(g) as s:
        s.upload(q, None, None)
        return s.download()


def compute_divB(g, bxb, byb):
    """Weno5_2D.jl:1986."""
    with Solver(g) as s:
        q = np.ones((g.Nx, g.Ny, 9), order="F")
        s.upload(q, bxb, byb)
        return s.divb()


def WENO5_2D(g, q, bxb, byb, nstep=None, tmax=None, log=None):
    """The driver loop of Weno5_2D.jl:33-110 with the state resident on the GPU.
    Returns (q, bxb, byb, t, steps)."""
    with Solver(g) as s:
        s.upload(q, bxb, byb)
        t, n = s.advance(nstep if nstep is not None else 10 ** 9, tmax or 0.0, 0.0)
        if log is not None:
            log(f"t = {t:g} after {n} steps; max |divB| = {np.abs(s.divb()).max():g}")
        q, bxb, byb = s.download()
        return q, bxb, byb, t, n


def WENO5_1D(g, q, tmax=0.2, nstep=None):
    """The driver loop of Weno5_1D.jl:33-148 (without plotting / the dt = 0 first step, Q12)."""
    with Solver(g) as s:
        s.upload(q)
        t, n = s.advance(nstep if nstep is not None else 10 ** 9, tmax or 0.0, 0.0)
        return s.download(), t, n
