"""
Host-side mirror of the reference's operator interface for the 3-D extension: the same names and
argument meaning as ``solver.py`` with one more axis --

    dt, vmax = timestep(q);  q, bxb, byb, bzb = classical_RK4_step(q, bxb, byb, bzb, dt);  divb = compute_divB(...)

on arrays ``q[Nx,Ny,Nz,9]``, ``bxb/byb/bzb[Nx,Ny,Nz]`` (Fortran order, NBnd = 3).  All compute happens in
libweno5mhd.so (weno5_*_3d); nothing here falls back to the CPU.
"""
import ctypes as C
import numpy as np

from . import lib as L
from .problems import NBND


def _farr(a, shape):
    a = np.asfortranarray(a, dtype=np.float64)
    if a.shape != tuple(shape):
        raise ValueError(f"expected array of shape {tuple(shape)}, got {a.shape}")
    return a


def make_params3d(g, es_roundtrip=True):
    return L.Params3D(g.nx, g.ny, g.nz, g.dx, g.dy, g.dz, g.gam, g.cfl, g.eps, int(g.bc_x), int(g.bc_y), int(g.bc_z),
                      1 if es_roundtrip else 0)


class Solver3D:
    """One GPU, the whole grid ``g`` (problems3d.Grid3D) resident in HBM."""

    def __init__(self, g, device=0, es_roundtrip=True):
        self.lib = L.load()
        self.g = g
        self.shape = (g.nx + 2 * NBND, g.ny + 2 * NBND, g.nz + 2 * NBND)
        self.params = make_params3d(g, es_roundtrip)
        h = C.c_void_p()
        L.check(self.lib.weno5_create_3d(C.byref(self.params), device, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.weno5_destroy_3d(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def upload(self, q, bxb=None, byb=None, bzb=None):
        q = _farr(q, self.shape + (9,))
        if bxb is not None:
            bxb, byb, bzb = (_farr(a, self.shape) for a in (bxb, byb, bzb))
        L.check(self.lib.weno5_upload_3d(self.h, L.dptr(q), L.dptr(bxb), L.dptr(byb), L.dptr(bzb)))

    def download(self):
        q = np.zeros(self.shape + (9,), order="F")
        b = [np.zeros(self.shape, order="F") for _ in range(3)]
        L.check(self.lib.weno5_download_3d(self.h, L.dptr(q), L.dptr(b[0]), L.dptr(b[1]), L.dptr(b[2])))
        return q, b[0], b[1], b[2]

    def timestep(self):
        dt, vm = C.c_double(), (C.c_double * 3)()
        L.check(self.lib.weno5_timestep_3d(self.h, C.byref(dt), vm))
        return dt.value, [float(v) for v in vm]

    def rk4_step(self, dt):
        L.check(self.lib.weno5_rk4_step_3d(self.h, float(dt)))

    def advance(self, nsteps, tmax=0.0, t=0.0):
        tt, nd = C.c_double(t), C.c_int(0)
        L.check(self.lib.weno5_advance_3d(self.h, int(nsteps), float(tmax), C.byref(tt), C.byref(nd)))
        return tt.value, nd.value

    def divb(self):
        d = np.zeros(self.shape, order="F")
        L.check(self.lib.weno5_divb_3d(self.h, L.dptr(d)))
        return d

    def totals(self):
        s = np.zeros(12)
        L.check(self.lib.weno5_totals_3d(self.h, L.dptr(s)))
        return s

    def launch_count(self):
        return int(self.lib.weno5_launch_count_3d(self.h))

    def profile(self, enable=True):
        L.check(self.lib.weno5_profile_3d(self.h, 1 if enable else 0))

    def profile_read(self):
        """(ms of the x, y, z sweep kernels, total launches) since profile(True) / the last read"""
        ms, n = (C.c_double * 3)(), C.c_ulonglong()
        L.check(self.lib.weno5_profile_read_3d(self.h, ms, C.byref(n)))
        return [float(x) for x in ms], int(n.value)


def classical_RK4_step_3d(g, q0, bxb0, byb0, bzb0, dt, es_roundtrip=True):
    """(q4, bxb4, byb4, bzb4) = classical_RK4_step(q0, bxb0, byb0, bzb0, dt): host arrays in, host arrays out."""
    shape = (g.nx + 2 * NBND, g.ny + 2 * NBND, g.nz + 2 * NBND)
    q0 = _farr(q0, shape + (9,))
    b0 = [_farr(a, shape) for a in (bxb0, byb0, bzb0)]
    q4 = np.zeros_like(q0, order="F")
    b4 = [np.zeros(shape, order="F") for _ in range(3)]
    P = make_params3d(g, es_roundtrip)
    L.check(L.load().weno5_classical_rk4_step_3d(C.byref(P), L.dptr(q0), L.dptr(b0[0]), L.dptr(b0[1]), L.dptr(b0[2]),
                                                 float(dt), L.dptr(q4), L.dptr(b4[0]), L.dptr(b4[1]), L.dptr(b4[2])))
    return q4, b4[0], b4[1], b4[2]
