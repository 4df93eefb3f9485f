"""Where does the GPU differ from the oracle on Orszag-Tang / Brio-Wu?  Prints the maximum relative error of one RK4
step as a function of the distance (in cells) to the nearest degenerate face (tangential field ~ 0), to calibrate
tests/test_gpu_parity.py::test_error_is_confined_to_degenerate_faces."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from julia_b200 import problems as pr  # noqa: E402
from julia_b200.solver import Solver  # noqa: E402
from oracle import oracle_c as O  # noqa: E402


def degenerate_distance_2d(q, thr):
    """cells' distance to the nearest x-sweep face with By_f^2+Bz_f^2 <= thr B_f^2 (along x) and to the nearest y-sweep
    face with Bx_f^2+Bz_f^2 <= thr B_f^2 (along y); the minimum of both."""
    Bx, By, Bz = q[..., 4], q[..., 5], q[..., 6]
    nx, ny = Bx.shape
    big = 10 ** 6
    dist = np.full((nx, ny), big)
    # x faces between i, i+1
    fy, fz, fx = 0.5 * (By[:-1] + By[1:]), 0.5 * (Bz[:-1] + Bz[1:]), 0.5 * (Bx[:-1] + Bx[1:])
    deg = (fy ** 2 + fz ** 2) <= thr * (fx ** 2 + fy ** 2 + fz ** 2)
    ii = np.arange(nx)[:, None]
    for j in range(ny):
        f = np.nonzero(deg[:, j])[0]
        if f.size:
            d = np.minimum(np.abs(ii - f[None, :]), np.abs(ii - (f[None, :] + 1))).min(axis=1)
            dist[:, j] = np.minimum(dist[:, j], d)
    gx, gz, gy = 0.5 * (Bx[:, :-1] + Bx[:, 1:]), 0.5 * (Bz[:, :-1] + Bz[:, 1:]), 0.5 * (By[:, :-1] + By[:, 1:])
    deg = (gx ** 2 + gz ** 2) <= thr * (gx ** 2 + gy ** 2 + gz ** 2)
    jj = np.arange(ny)[:, None]
    for i in range(nx):
        f = np.nonzero(deg[i, :])[0]
        if f.size:
            d = np.minimum(np.abs(jj - f[None, :]), np.abs(jj - (f[None, :] + 1))).min(axis=1)
            dist[i, :] = np.minimum(dist[i, :], d)
    return dist


def main():
    for n in (128, 256):
        g, q, bx, by = pr.orszag_tang(n)
        with Solver(g) as s:
            s.upload(q, bx, by)
            dt = s.timestep()[0]
            s.rk4_step(dt)
            qg, bxg, byg = s.download()
        qo, bxo, byo = O.rk4_step_2d(g, q, bx, by, dt)
        err = np.zeros(q.shape[:2])
        for c in (0, 1, 2, 4, 5, 7, 8):
            err = np.maximum(err, np.abs(qg[..., c] - qo[..., c]) / np.abs(qo[..., c]).max())
        for thr in (1e-20, 1e-12, 1e-8, 1e-6):
            dist = degenerate_distance_2d(q, thr)
            prof = [float(err[dist == d].max()) if (dist == d).any() else 0.0 for d in range(0, 20)]
            far = float(err[dist >= 20].max()) if (dist >= 20).any() else 0.0
            print(f"OT {n}^2 thr={thr:g} total Linf={err.max():.2e} far(>=20)={far:.2e} frac within 13: {(dist <= 13).mean():.3f}")
            print("   by distance:", " ".join(f"{p:.1e}" for p in prof))
    for nx in (256, 1024):
        g, q = pr.brio_wu(nx)
        with Solver(g) as s:
            s.upload(q)
            dt = 0.4 * g.dx
            s.rk4_step(dt)
            qg = s.download()
        qo = O.rk4_step_1d(g, q, dt)
        err = np.zeros(q.shape[0])
        for c in (0, 1, 2, 5, 8):
            err = np.maximum(err, np.abs(qg[:, c] - qo[:, c]) / np.abs(qo[:, c]).max())
        mid = q.shape[0] // 2
        prof = [float(err[max(mid - d - 1, 0):mid + d + 1].max()) for d in range(0, 20)]
        print(f"BW {nx}: total Linf={err.max():.2e}; max error within d cells of the interface:", " ".join(f"{p:.1e}" for p in prof))
        print("   outside 13 cells:", float(np.delete(err, np.arange(mid - 14, mid + 14)).max()))


if __name__ == "__main__":
    main()
