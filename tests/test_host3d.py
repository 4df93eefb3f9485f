"""
CPU check of the PRODUCT's 3-D code path: the kernel bodies and the stage sequence of julia_b200/csrc/weno5_3d_body.cuh
-- the very functions the sm_100a kernels of weno5_3d.cu wrap -- run thread by thread on the host
(tests/hostmath/host3d.cpp) and must reproduce the 3-D oracle, which is itself pinned to the reference's 2-D step by
reduction (tests/test_oracle3d.py).  Grids are chosen so that lines cross CTA-tile boundaries in every direction.
"""
import ctypes as C
import numpy as np
import pytest

from julia_b200 import problems as pr
from julia_b200 import problems3d as p3
from oracle import oracle_c as O

I = slice(3, -3)


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def host_step(lib, g, q, bx, by, bz, dt, roundtrip=True, march=0):
    """march = 0: tile kernel bodies in all three directions; > 0: the marching body with that many segments per
    line; 100 s + m: x by the warp-per-line body with s segments,
    y and z marching with m segments, + 10000 l: l lanes per line in the x body (default 32); negative: the same
    with the split EMF / face-update kernels instead of the fused one"""
    q4 = np.zeros_like(q, order="F")
    b4 = [np.zeros_like(bx, order="F") for _ in range(3)]
    h = (C.c_double * 3)(g.dx, g.dy, g.dz)
    bc = (C.c_int * 3)(int(g.bc_x), int(g.bc_y), int(g.bc_z))
    rc = lib.h3_rk4_step(g.nx, g.ny, g.nz, h, g.gam, g.eps, bc, int(roundtrip), int(march), dp(q), dp(bx), dp(by), dp(bz), float(dt),
                         dp(q4), dp(b4[0]), dp(b4[1]), dp(b4[2]))
    assert rc == 0
    return q4, b4[0], b4[1], b4[2]


def compare(got, want, tol):
    """whole arrays, ghosts included: the kernels' gather-style ghost handling against boundaries! applied in place"""
    assert np.isfinite(got[0]).all(), "the kernels read something they never wrote"
    for c in range(9):
        s = max(np.abs(want[0][..., c]).max(), 1.0)
        assert np.abs(got[0][..., c] - want[0][..., c]).max() <= tol * s, c
    for a, b in zip(got[1:], want[1:]):
        assert np.isfinite(a).all()
        assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1.0)


@pytest.mark.parametrize("march", [0, 1, 3, 102, 301, 80102, 160201, -80102])
@pytest.mark.parametrize("shape", [(130, 5, 4), (5, 36, 4), (4, 5, 36), (37, 19, 18)])
@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 0, 0), (1, 0, 1), (0, 1, 0)])
def test_kernel_bodies_reproduce_the_oracle_step(host3d, shape, bc, march):
    g, q, bx, by, bz = p3.smooth_3d(*shape, seed=sum(shape) + bc[0])
    g.bc_x, g.bc_y, g.bc_z = bc
    # re-derive ghosts for the boundary conditions under test, the way the driver would before its first step
    q, bx, by, bz = (np.asfortranarray(a.copy()) for a in (q, bx, by, bz))
    O.boundaries_3d(g, q, bx, by, bz)
    dt = 0.8 * O.timestep_3d(g, q)[0]
    want = O.rk4_step_3d(g, q, bx, by, bz, dt)
    got = host_step(host3d, g, q, bx, by, bz, dt, march=march)
    compare(got, want, 2e-13)


def test_two_steps_and_no_roundtrip(host3d):
    g, q, bx, by, bz = p3.orszag_tang_3d(12, nz=8)
    dt = O.timestep_3d(g, q)[0]
    want = got = (q, bx, by, bz)
    for _ in range(2):
        want = O.rk4_step_3d(g, *want, dt, roundtrip=False)
        got = host_step(host3d, g, *got, dt, roundtrip=False, march=2)
    # Orszag-Tang has degenerate faces (B_t = 0 lines), where alpha_s is the square root of a round-off residual and
    # any two correct implementations differ; measured 5e-11
    compare(got, want, 1e-9)


@pytest.mark.parametrize("march", [0, 2, 80101])
@pytest.mark.parametrize("orientation", [0, 1, 2])
def test_kernel_bodies_reduce_to_the_2d_reference_step(host3d, orientation, march):
    g2, q, bx, by = pr.orszag_tang(16, ny=12, ysize=0.75)
    dt = 0.6 * O.timestep_2d(g2, q)[0]
    q2, bx2, by2 = O.rk4_step_2d(g2, q, bx, by, dt)
    g, q3, b0, b1, b2 = p3.embed_2d(g2, q, bx, by, orientation, n3=4)
    r = host_step(host3d, g, q3, b0, b1, b2, dt, march=march)
    # the reduction needs vz = Bz = 0, i.e. Orszag-Tang, which has degenerate faces: the structured eigen-system of
    # the product and the dense one of the oracle differ there at the 1e-10 level (measured 5e-11)
    tol = 1e-9
    for c in (0, 7, 8):
        got = p3.extract_2d(r[0][..., c], orientation, 4)
        assert np.abs(got[I, I] - q2[I, I, c]).max() <= tol * np.abs(q2[..., c]).max()
    for base in ([1, 2, 3], [4, 5, 6]):
        for m2 in range(3):
            got = p3.extract_2d(r[0][..., base[p3.vector_component(orientation, m2)]], orientation, 4)
            assert np.abs(got[I, I] - q2[I, I, base[m2]]).max() <= tol
    faces = {orientation % 3: bx2, (orientation + 1) % 3: by2}
    for ax, want in faces.items():
        assert np.abs(p3.extract_2d(r[1 + ax], orientation, 4)[I, I] - want[I, I]).max() <= tol


def test_timestep_divb_center2face(host3d):
    g, q, bx, by, bz = p3.smooth_3d(9, 8, 7)
    h = (C.c_double * 3)(g.dx, g.dy, g.dz)
    dt, vm = C.c_double(), (C.c_double * 3)()
    host3d.h3_timestep(g.nx, g.ny, g.nz, h, g.gam, g.cfl, dp(q), C.byref(dt), vm)
    dto, vmo = O.timestep_3d(g, q)
    assert abs(dt.value - dto) <= 1e-13 * dto and np.allclose(list(vm), vmo, rtol=1e-13, atol=0)
    f = [np.zeros(g.shape, order="F") for _ in range(4)]
    host3d.h3_divb_center2face(g.nx, g.ny, g.nz, h, dp(q), dp(f[0]), dp(f[1]), dp(f[2]), dp(f[3]))
    for a, b in zip(f[:3], O.center2face_3d(g, q)):
        assert np.abs(a - b).max() <= 1e-15          # the host build contracts multiply-adds, the oracle does not
    assert np.abs(f[3] - O.divb_3d(g, f[0], f[1], f[2])).max() <= 1e-12


def test_xwarp_plan(host3d):
    ns, lpl = C.c_int(), C.c_int()
    host3d.h3_xwarp_plan(256, 256, 256, 1, 148, C.byref(ns), C.byref(lpl))       # 257 faces: 33 steps of 8 lanes
    assert (ns.value, lpl.value) == (1, 8)
    host3d.h3_xwarp_plan(255, 256, 256, 1, 148, C.byref(ns), C.byref(lpl))       # 256 faces: 8 full steps of 32
    assert (ns.value, lpl.value) == (1, 32)
    host3d.h3_xwarp_plan(256, 8, 8, 1, 148, C.byref(ns), C.byref(lpl))           # 64 lines only: cut them
    assert ns.value > 1 and lpl.value in (8, 16, 32)


def test_march_segment_choice(host3d):
    # 256^3 on 148 SMs: 8 x 32 CTAs per segment -> 4 segments of 64 cells fill 6.9 of 7 waves
    assert host3d.h3_march_segments(256, 256, 256, 1, 148) == 4
    for n in (8, 40, 100, 512):
        ns = host3d.h3_march_segments(n, n, n, 2, 148)
        assert 1 <= ns <= max(1, n // 8)
