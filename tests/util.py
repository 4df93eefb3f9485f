import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel_linf(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def rel_l1(a, b):
    return float(np.abs(a - b).sum() / max(np.abs(b).sum(), 1e-300))


def interior(a):
    return a[3:-3, 3:-3] if a.ndim >= 2 and a.shape[1] > 9 else a[3:-3]


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def degenerate_distance_2d(q, thr=1e-20):
    """Distance (cells, along the sweep) of every cell to the nearest DEGENERATE face of the state q[Nx,Ny,9]: an
    x-sweep face whose averaged tangential field (By,Bz) has |Bt|^2 <= thr |B|^2, or a y-sweep face with the same for
    (Bx,Bz).  There alpha_s = sqrt(max(0, cf^2-a^2)/(cf^2-cs^2)) is the square root of a round-off residual
    (Weno5_2D.jl:1355-1356) and any two correct implementations differ."""
    Bx, By, Bz = q[..., 4], q[..., 5], q[..., 6]
    nx, ny = Bx.shape
    dist = np.full((nx, ny), 10 ** 6)
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
