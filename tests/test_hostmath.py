"""CPU check of the PRODUCT's per-face arithmetic (julia_b200/csrc/weno5_math.cuh compiled for the
host): the structured eigen-projection, the one-reciprocal WENO weights and the structured
back-projection must reproduce the oracle's dense L/R formulation to round-off."""
import ctypes as C
import numpy as np
import pytest

from julia_b200 import problems as pr
from oracle import oracle_c as O


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


CASES = {
    "alfven": lambda: pr.alfven_wave(40),
    "blast": lambda: pr.blast(40),
    "random_smooth": lambda: pr.random_state(40, 24),
    "random_rough": lambda: pr.random_state(40, 24, smooth=False),
    "shock_cloud": lambda: pr.shock_cloud(64, 16),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("direction", [0, 1])
def test_structured_sweep_matches_oracle(hostmath, name, direction):
    g, q, bx, by = CASES[name]()
    q8 = np.asfortranarray(q[..., [0, 1, 2, 3, 4, 5, 6, 8]])
    dq = np.zeros_like(q8, order="F")
    bs = np.zeros(q8.shape[:2], order="F")
    bc = g.bc_y if direction else g.bc_x
    hostmath.hm_sweep(g.Nx, g.Ny, dp(q8), direction, int(bc), g.gam, g.eps, dp(dq), dp(bs))
    dqo, bso = O.sweep_2d(g, q8, direction)
    sl = (slice(2, g.Nx - 2), slice(2, g.Ny - 2))
    # tolerance: 1e-12 of the largest flux difference (shock-cloud has |dq| ~ 1e4)
    assert np.abs(dq[sl] - dqo[sl]).max() <= 1e-12 * max(1.0, np.abs(dqo[sl]).max())
    assert np.abs(bs[sl] - bso[sl]).max() <= 1e-12 * max(1.0, np.abs(bso[sl]).max())


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("direction", [0, 1])
def test_structured_sweep_S_branch_matches_oracle(hostmath, name, direction):
    """The entropy-variable sweep against the numpy restatement of
    weno5_flux_difference_S, which tests/test_reference_pin.py pins against the reference."""
    from oracle import weno5_numpy as M
    g, q, bx, by = CASES[name]()
    qS = np.asfortranarray(q[..., 0:8])
    dq = np.zeros_like(qS, order="F")
    bs = np.zeros(qS.shape[:2], order="F")
    bc = g.bc_y if direction else g.bc_x
    hostmath.hm_sweep_S(g.Nx, g.Ny, dp(qS), direction, int(bc), g.gam, g.eps, dp(dq), dp(bs))
    bcname = {0: "outflow", 1: "periodic"}[int(bc)]
    if direction == 0:
        dqo, bso, _ = M.flux_difference_S(np.ascontiguousarray(qS), g.gam, g.eps, bcname)
    else:
        dqr, _, gr = M.flux_difference_S(M.rotate_fwd(np.ascontiguousarray(qS)), g.gam, g.eps, bcname)
        dqo, bso = M.rotate_bwd(dqr), gr.T
    sl = (slice(2, g.Nx - 2), slice(2, g.Ny - 2))
    assert np.abs(dq[sl] - dqo[sl]).max() <= 1e-12 * max(1.0, np.abs(dqo[sl]).max())
    assert np.abs(bs[sl] - bso[sl]).max() <= 1e-12 * max(1.0, np.abs(bso[sl]).max())


def test_degenerate_faces_are_ill_conditioned_but_bounded(hostmath):
    """Orszag-Tang has faces where By changes sign with Bz = 0: there alpha_s is the square
    root of a round-off residual (Weno5_2D.jl:1355-1356), so two correct implementations differ
    by ~1e-9 -- a property of the reference algorithm, see DESIGN.md."""
    g, q, bx, by = pr.orszag_tang(48)
    q8 = np.asfortranarray(q[..., [0, 1, 2, 3, 4, 5, 6, 8]])
    dq = np.zeros_like(q8, order="F")
    bs = np.zeros(q8.shape[:2], order="F")
    hostmath.hm_sweep(g.Nx, g.Ny, dp(q8), 0, 1, g.gam, g.eps, dp(dq), dp(bs))
    dqo, bso = O.sweep_2d(g, q8, 0)
    sl = (slice(2, g.Nx - 2), slice(2, g.Ny - 2))
    err = np.abs(dq[sl] - dqo[sl])
    assert err.max() < 1e-7
    rows = np.unique(np.where(err > 1e-12)[0])
    assert len(rows) <= 16          # confined to the few lines of degenerate faces


def test_entropy_conversion(hostmath):
    from oracle import weno5_numpy as M
    rng = np.random.default_rng(0)
    for _ in range(50):
        rho, P = rng.uniform(0.5, 2), rng.uniform(0.1, 2)
        m, b = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
        E = P / (5 / 3 - 1) + 0.5 * ((m ** 2).sum() / rho + (b ** 2).sum())
        q = np.array([[rho, *m, *b, E]])
        S = hostmath.hm_e2s(rho, *m, *b, E, 5 / 3)
        assert abs(S - M.E2S(q, 5 / 3)[0]) <= 1e-14 * abs(S)
        q[0, 7] = S
        E2 = hostmath.hm_s2e(rho, *m, *b, S, 5 / 3)
        assert abs(E2 - M.S2E(q, 5 / 3)[0]) <= 1e-14 * E
        # the fused round trip of the kernels is the same pair of numbers
        import ctypes as C
        out = (C.c_double * 2)()
        hostmath.hm_e2s_s2e(*(C.c_double(v) for v in (rho, *m, *b, E, 5 / 3)), out)
        assert abs(out[0] - S) <= 2e-16 * abs(S) and abs(out[1] - E2) <= 2e-16 * E
