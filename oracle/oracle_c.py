"""
TEST INFRASTRUCTURE ONLY -- ctypes binding of the C oracle (oracle/weno5_oracle.c).

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs only.  The product package (julia_b200) must never import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
OUTFLOW, PERIODIC = 0, 1


class OrcParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("dx", C.c_double), ("dy", C.c_double),
                ("gam", C.c_double), ("cfl", C.c_double), ("eps", C.c_double),
                ("bc_x", C.c_int), ("bc_y", C.c_int), ("variant", C.c_int)]


class OrcParams3(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("dx", C.c_double), ("dy", C.c_double),
                ("dz", C.c_double), ("gam", C.c_double), ("cfl", C.c_double), ("eps", C.c_double),
                ("bc_x", C.c_int), ("bc_y", C.c_int), ("bc_z", C.c_int)]


def build(force=False):
    libs = [os.path.join(_HERE, n) for n in ("liboracle.so", "liboracle_fast.so")]
    src = os.path.join(_HERE, "weno5_oracle.c")
    if force or any((not os.path.exists(l)) or os.path.getmtime(l) < os.path.getmtime(src) for l in libs):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return libs


_libs = {}


def lib(fast=False):
    key = "fast" if fast else "strict"
    if key not in _libs:
        paths = build()
        L = C.CDLL(paths[1 if fast else 0])
        dp = C.POINTER(C.c_double)
        pp = C.POINTER(OrcParams)
        L.orc_rk4_step_2d.argtypes = [pp, dp, dp, dp, C.c_double, dp, dp, dp]
        L.orc_timestep_2d.argtypes = [pp, dp, dp, dp, dp]
        L.orc_divb.argtypes = [pp, dp, dp, dp]
        L.orc_center2face.argtypes = [pp, dp, dp, dp]
        L.orc_boundaries_2d.argtypes = [pp, dp, dp, dp]
        L.orc_sweep_2d.argtypes = [pp, dp, C.c_int, dp, dp]
        L.orc_rk4_step_1d.argtypes = [pp, dp, C.c_double, dp]
        L.orc_advance_2d.argtypes = [pp, dp, dp, dp, C.c_int, dp]
        L.orc_sweep_2d_S.argtypes = [pp, dp, C.c_int, dp, dp]
        L.orc_rk4_step_2d_dual.argtypes = [pp, dp, dp, dp, C.c_double, C.c_double, dp, dp, dp, C.POINTER(C.c_long)]
        L.orc_num_threads.restype = C.c_int
        p3 = C.POINTER(OrcParams3)
        L.orc_rk4_step_3d.argtypes = [p3, dp, dp, dp, dp, C.c_double, C.c_int, dp, dp, dp, dp]
        L.orc_timestep_3d.argtypes = [p3, dp, dp, dp]
        L.orc_divb_3d.argtypes = [p3, dp, dp, dp, dp]
        L.orc_center2face_3d.argtypes = [p3, dp, dp, dp, dp]
        L.orc_boundaries_3d.argtypes = [p3, dp, dp, dp, dp]
        L.orc_advance_3d.argtypes = [p3, dp, dp, dp, dp, C.c_int, C.c_int, dp]
        _libs[key] = L
    return _libs[key]


def params(g, variant=0):
    """g: julia_b200.problems.Grid (or anything with the same attributes)."""
    return OrcParams(g.nx, g.ny, g.dx, g.dy, g.gam, g.cfl, g.eps, int(g.bc_x), int(g.bc_y), variant)


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f(a):
    a = np.asfortranarray(a, dtype=np.float64)
    return a


def rk4_step_2d(g, q, bxb, byb, dt, fast=False):
    q, bxb, byb = _f(q), _f(bxb), _f(byb)
    q4, bx4, by4 = np.zeros_like(q, order="F"), np.zeros_like(bxb, order="F"), np.zeros_like(byb, order="F")
    P = params(g)
    rc = lib(fast).orc_rk4_step_2d(C.byref(P), _p(q), _p(bxb), _p(byb), float(dt), _p(q4), _p(bx4), _p(by4))
    assert rc == 0
    return q4, bx4, by4


def timestep_2d(g, q, fast=False):
    q = _f(q)
    dt, vx, vy = C.c_double(), C.c_double(), C.c_double()
    P = params(g)
    lib(fast).orc_timestep_2d(C.byref(P), _p(q), C.byref(dt), C.byref(vx), C.byref(vy))
    return dt.value, vx.value, vy.value


def advance_2d(g, q, bxb, byb, nsteps, fast=False):
    q, bxb, byb = _f(q).copy(order="F"), _f(bxb).copy(order="F"), _f(byb).copy(order="F")
    t = C.c_double(0.0)
    P = params(g)
    rc = lib(fast).orc_advance_2d(C.byref(P), _p(q), _p(bxb), _p(byb), int(nsteps), C.byref(t))
    assert rc == 0
    return q, bxb, byb, t.value


def divb(g, bxb, byb):
    bxb, byb = _f(bxb), _f(byb)
    d = np.zeros_like(bxb, order="F")
    P = params(g)
    lib().orc_divb(C.byref(P), _p(bxb), _p(byb), _p(d))
    return d


def center2face(g, q):
    q = _f(q)
    bxb = np.zeros(q.shape[:2], order="F")
    byb = np.zeros(q.shape[:2], order="F")
    P = params(g)
    lib().orc_center2face(C.byref(P), _p(q), _p(bxb), _p(byb))
    return bxb, byb


def boundaries_2d(g, q=None, a=None, b=None):
    P = params(g)
    for x in (q, a, b):
        assert x is None or (x.flags.f_contiguous and x.dtype == np.float64)
    nul = C.POINTER(C.c_double)()
    lib().orc_boundaries_2d(C.byref(P), _p(q) if q is not None else nul,
                            _p(a) if a is not None else nul, _p(b) if b is not None else nul)


def sweep_2d(g, q8, direction):
    """q8: (Nx,Ny,8) [rho,M,B,E].  Returns dq (Nx,Ny,8) and fsy (dir 0) / gsx (dir 1)."""
    q8 = _f(q8)
    dq = np.zeros_like(q8, order="F")
    bs = np.zeros(q8.shape[:2], order="F")
    P = params(g)
    lib().orc_sweep_2d(C.byref(P), _p(q8), int(direction), _p(dq), _p(bs))
    return dq, bs


def sweep_2d_S(g, q8, direction):
    """weno5_flux_difference_S (Weno5_2D.jl:361): q8 = (Nx,Ny,8) [rho,M,B,S] -> dq, fsy (dir 0) / gsx (dir 1)."""
    q8 = _f(q8)
    dq = np.zeros_like(q8, order="F")
    bs = np.zeros(q8.shape[:2], order="F")
    P = params(g)
    lib().orc_sweep_2d_S(C.byref(P), _p(q8), int(direction), _p(dq), _p(bs))
    return dq, bs


def rk4_step_2d_dual(g, q, bxb, byb, dt, threshold=1e-5, fast=False):
    """classical_RK4_step with both branches and the live ES_Switch.  Returns (q4, bxb4, byb4, nbad)."""
    q, bxb, byb = _f(q), _f(bxb), _f(byb)
    q4, bx4, by4 = np.zeros_like(q, order="F"), np.zeros_like(bxb, order="F"), np.zeros_like(byb, order="F")
    nbad = (C.c_long * 4)()
    P = params(g)
    rc = lib(fast).orc_rk4_step_2d_dual(C.byref(P), _p(q), _p(bxb), _p(byb), float(dt), float(threshold),
                                        _p(q4), _p(bx4), _p(by4), nbad)
    assert rc == 0
    return q4, bx4, by4, [int(x) for x in nbad]


def rk4_step_1d(g, q, dt):
    q = _f(q)
    q4 = np.zeros_like(q, order="F")
    P = params(g, variant=1)
    lib().orc_rk4_step_1d(C.byref(P), _p(q), float(dt), _p(q4))
    return q4


def vmax_1d(g, q):
    q = _f(q)
    dt, vx, vy = C.c_double(), C.c_double(), C.c_double()
    P = params(g, variant=1)
    P.ny = 1
    lib().orc_timestep_2d(C.byref(P), _p(q), C.byref(dt), C.byref(vx), C.byref(vy))
    return vx.value, dt.value


def num_threads(fast=True):
    return lib(fast).orc_num_threads()


# ----------------------------------------------------------------------------- 3-D extension
def params3(g):
    """g: julia_b200.problems3d.Grid3D."""
    return OrcParams3(g.nx, g.ny, g.nz, g.dx, g.dy, g.dz, g.gam, g.cfl, g.eps, int(g.bc_x), int(g.bc_y), int(g.bc_z))


def rk4_step_3d(g, q, bxb, byb, bzb, dt, roundtrip=True, fast=False):
    q, bxb, byb, bzb = _f(q), _f(bxb), _f(byb), _f(bzb)
    q4 = np.zeros_like(q, order="F")
    b4 = [np.zeros_like(bxb, order="F") for _ in range(3)]
    P = params3(g)
    rc = lib(fast).orc_rk4_step_3d(C.byref(P), _p(q), _p(bxb), _p(byb), _p(bzb), float(dt), int(bool(roundtrip)),
                                   _p(q4), _p(b4[0]), _p(b4[1]), _p(b4[2]))
    assert rc == 0
    return q4, b4[0], b4[1], b4[2]


def timestep_3d(g, q, fast=False):
    q = _f(q)
    dt = C.c_double()
    vm = (C.c_double * 3)()
    P = params3(g)
    lib(fast).orc_timestep_3d(C.byref(P), _p(q), C.byref(dt), vm)
    return dt.value, [float(v) for v in vm]


def advance_3d(g, q, bxb, byb, bzb, nsteps, roundtrip=True, fast=False):
    q, bxb, byb, bzb = (_f(a).copy(order="F") for a in (q, bxb, byb, bzb))
    t = C.c_double(0.0)
    P = params3(g)
    rc = lib(fast).orc_advance_3d(C.byref(P), _p(q), _p(bxb), _p(byb), _p(bzb), int(nsteps), int(bool(roundtrip)),
                                  C.byref(t))
    assert rc == 0
    return q, bxb, byb, bzb, t.value


def divb_3d(g, bxb, byb, bzb):
    bxb, byb, bzb = _f(bxb), _f(byb), _f(bzb)
    d = np.zeros_like(bxb, order="F")
    P = params3(g)
    lib().orc_divb_3d(C.byref(P), _p(bxb), _p(byb), _p(bzb), _p(d))
    return d


def center2face_3d(g, q):
    q = _f(q)
    b = [np.zeros(q.shape[:3], order="F") for _ in range(3)]
    P = params3(g)
    lib().orc_center2face_3d(C.byref(P), _p(q), _p(b[0]), _p(b[1]), _p(b[2]))
    return b


def boundaries_3d(g, q=None, a=None, b=None, c=None):
    P = params3(g)
    nul = C.POINTER(C.c_double)()
    for x in (q, a, b, c):
        assert x is None or (x.flags.f_contiguous and x.dtype == np.float64)
    lib().orc_boundaries_3d(C.byref(P), *[(_p(x) if x is not None else nul) for x in (q, a, b, c)])
