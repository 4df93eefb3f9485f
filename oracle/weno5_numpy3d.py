"""
TEST INFRASTRUCTURE ONLY -- second, independent restatement of the 3-D extension of classical_RK4_step
(oracle/weno5_oracle.c, orc_rk4_step_3d, is the first).  Array-at-a-time numpy on top of the 2-D restatement's
directional sweep (oracle/weno5_numpy.py flux_difference_E, pinned against the reference): every direction is swept by
that same function on a rotated copy of the state -- literally what the reference's docstring describes
(Weno5_2D.jl:113-123, rotate_state :1897-1927) -- and the constrained-transport part is written with whole-array
slices instead of the C oracle's index loops.  Shares no 3-D code with the C oracle.
Layout: q (Nx,Ny,Nz,9) [rho,Mx,My,Mz,Bx,By,Bz,S,E], faces (Nx,Ny,Nz); NBnd = 3.
"""
import numpy as np

from . import weno5_numpy as M

NB = M.NB
# sweep frame for axis d: components (rho, M_d, M_d+1, M_d+2, B_d, B_d+1, B_d+2, E) as indices into q8 = [rho,M,B,E]
_PERM = ([0, 1, 2, 3, 4, 5, 6, 7], [0, 2, 3, 1, 5, 6, 4, 7], [0, 3, 1, 2, 6, 4, 5, 7])


def _fill_cells(a, bcs):
    for ax, bc in enumerate(bcs):
        v = np.moveaxis(a, ax, 0)
        N = v.shape[0]
        n = N - 2 * NB
        if bc == 1:
            v[0:NB] = v[n:n + NB]
            v[N - NB:N] = v[NB:2 * NB]
        else:
            v[0:NB] = v[NB]
            v[N - NB:N] = v[N - NB - 1]


def _fill_flux(a, bcs):
    """boundaries! on a flux / face array (Weno5_2D.jl:1704-1729), per axis: low sides, high sides, rims"""
    for ax, bc in enumerate(bcs):
        v = np.moveaxis(a, ax, 0)
        N = v.shape[0]
        n = N - 2 * NB
        if bc == 1:
            v[0:NB] = v[n:n + NB]
            v[N - NB:N] = v[NB:2 * NB]
        else:
            v[0:2] = v[2]
    for ax, bc in enumerate(bcs):
        if bc != 1:
            v = np.moveaxis(a, ax, 0)
            N = v.shape[0]
            v[N - 2:N] = v[N - 3]
    for ax, bc in enumerate(bcs):
        if bc != 1:
            v = np.moveaxis(a, ax, 0)
            v[0] = 0.0
            v[-1] = 0.0


def sweep(q8, axis, gam, eps, bc):
    """weno5_flux_difference_E along `axis`: returns the fhat differences of (rho, Mx, My, Mz, E) and the two
    transport fluxes (of B_axis+1 and B_axis+2), all in the volume's own orientation."""
    rot = np.moveaxis(q8[..., _PERM[axis]], axis, 0)               # sweep axis first, components rotated
    shp = rot.shape
    dq, b2, b3 = M.flux_difference_E(np.ascontiguousarray(rot.reshape(shp[0], -1, 8)), gam, eps,
                                     "periodic" if bc == 1 else "outflow")
    dq = np.moveaxis(dq.reshape(shp), 0, axis)
    b2 = np.moveaxis(b2.reshape(shp[:3]), 0, axis)
    b3 = np.moveaxis(b3.reshape(shp[:3]), 0, axis)
    out = np.zeros(q8.shape[:3] + (5,))
    out[..., 0] = dq[..., 0]
    for m in range(3):                                             # rotated momentum m is physical component (axis + m) % 3
        out[..., 1 + (axis + m) % 3] = dq[..., 1 + m]
    out[..., 4] = dq[..., 7]
    return out, b2, b3


def stage(qk, base8, bb, c, dt, h, gam, eps, bcs, roundtrip):
    q8 = np.concatenate([qk[..., 0:7], qk[..., 8:9]], axis=-1)
    dFx, fy, fz = sweep(q8, 0, gam, eps, bcs[0])
    dFy, gz, gx = sweep(q8, 1, gam, eps, bcs[1])
    dFz, hx, hy = sweep(q8, 2, gam, eps, bcs[2])
    for f in (fy, fz, gz, gx, hx, hy):
        _fill_flux(f, bcs)
    qn = np.zeros_like(q8)
    rhs = (dFx / h[0] + dFy / h[1]) + dFz / h[2]
    for m, comp in enumerate((0, 1, 2, 3, 7)):
        qn[..., comp] = base8[..., comp] - c * dt * rhs[..., m]
    I = slice(1, -1)
    Ox, Oy, Oz = (np.zeros(q8.shape[:3]) for _ in range(3))
    Ox[I, I, I] = 0.5 * (hy[I, I, I] + hy[I, 2:, I] - gz[I, I, I] - gz[I, I, 2:])
    Oy[I, I, I] = 0.5 * (fz[I, I, I] + fz[I, I, 2:] - hx[I, I, I] - hx[2:, I, I])
    Oz[I, I, I] = 0.5 * (gx[I, I, I] + gx[2:, I, I] - fy[I, I, I] - fy[I, 2:, I])
    bn = [b.copy() for b in bb]
    L = slice(0, -2)
    bn[0][I, I, I] = bb[0][I, I, I] - c * dt / h[1] * (Oz[I, I, I] - Oz[I, L, I]) + c * dt / h[2] * (Oy[I, I, I] - Oy[I, I, L])
    bn[1][I, I, I] = bb[1][I, I, I] - c * dt / h[2] * (Ox[I, I, I] - Ox[I, I, L]) + c * dt / h[0] * (Oz[I, I, I] - Oz[L, I, I])
    bn[2][I, I, I] = bb[2][I, I, I] - c * dt / h[0] * (Oy[I, I, I] - Oy[L, I, I]) + c * dt / h[1] * (Ox[I, I, I] - Ox[I, L, I])
    for ax in range(3):                                            # interpolateB_face2center, Weno5_2D.jl:344-356
        v = np.moveaxis(bn[ax], ax, 0)
        N = v.shape[0]
        cen = np.moveaxis(qn[..., 4 + ax], ax, 0)
        cen[2:N - 1] = (-v[0:N - 3] + 9.0 * v[1:N - 2] + 9.0 * v[2:N - 1] - v[3:N]) / 16.0
    for comp in range(8):
        _fill_cells(qn[..., comp], bcs)
    out = np.zeros(qk.shape)
    out[..., 0:7] = qn[..., 0:7]
    rho, B2 = qn[..., 0], (qn[..., 4:7] ** 2).sum(axis=-1)
    mom2 = (qn[..., 1:4] ** 2).sum(axis=-1)
    S = (qn[..., 7] - 0.5 * mom2 / rho - B2 / 2) * (gam - 1) / rho ** (gam - 1)          # E2S, :1770
    out[..., 7] = S
    out[..., 8] = rho ** (gam - 1) * S / (gam - 1) + B2 / 2 + 0.5 * mom2 / rho if roundtrip else qn[..., 7]   # S2E, :1782
    return out, bn


def classical_RK4_step_3d(g, q0, bxb0, byb0, bzb0, dt, roundtrip=True):
    """g: julia_b200.problems3d.Grid3D (or anything with its attributes)"""
    h = (g.dx, g.dy, g.dz)
    bcs = (int(g.bc_x), int(g.bc_y), int(g.bc_z))
    b0 = [np.array(b, dtype=np.float64) for b in (bxb0, byb0, bzb0)]
    q0 = np.array(q0, dtype=np.float64)
    q0E = np.concatenate([q0[..., 0:7], q0[..., 8:9]], axis=-1)
    a = (g.gam, g.eps, bcs, roundtrip)
    q1, b1 = stage(q0, q0E, b0, 0.5, dt, h, *a)
    q2, b2 = stage(q1, q0E, b0, 0.5, dt, h, *a)
    q3, b3 = stage(q2, q0E, b0, 1.0, dt, h, *a)
    E = lambda q: np.concatenate([q[..., 0:7], q[..., 8:9]], axis=-1)
    base = 1.0 / 3 * (-q0E + E(q1) + 2 * E(q2) + E(q3))
    bbase = [1.0 / 3 * (-b0[m] + b1[m] + 2 * b2[m] + b3[m]) for m in range(3)]
    q4, b4 = stage(q3, base, bbase, 1.0 / 6, dt, h, *a)
    for b in b4:
        _fill_flux(b, bcs)
    return q4, b4[0], b4[1], b4[2]
