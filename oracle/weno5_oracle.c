/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the WENO5 MHD hot path.
 *
 * A plain-C restatement of the algorithm of jdonnert/julia's Weno5_2D.jl /
 * Weno5_1D.jl (E branch = the shipped behaviour; S branch and live ES_Switch = orc_rk4_step_2d_dual) (reference file:line cited at every function).  It is the checker for the
 * CUDA path (tests/, __graft_entry__.smoke()) and the timed CPU baseline of bench.py;
 * nothing in the product (julia_b200/, libweno5mhd.so) may call, link or load it.
 *
 * PARITY PINNED against outputs of the reference itself: the reference (Julia 0.6 source;
 * no julia binary here) is translated mechanically by oracle/jl2py.py and its own functions
 * are executed unmodified by scripts/make_reference_golden.py; the resulting vectors are
 * committed as tests/golden/ref_*.npz and checked by tests/test_reference_pin.py: every
 * cell-local function, both sweeps (on all cells whose faces the reference computes in
 * bounds), the constrained-transport pieces, boundaries!, timestep, the full 2-D RK4 step
 * (periodic ghosts) and the 1-D RK4 step are BIT-IDENTICAL to the reference (E2S/S2E to
 * 2 ulp of libm pow); with the reference's outflow boundaries! the step differs by the
 * documented quirk Q3 (<= 1e-8 of the field scale on the reference's shock-cloud problem).
 * Further pins: (i) algebraic identities of the eigen-system, (ii) invariants
 * (conservation, div B, uniform-state preservation), (iii) agreement to round-off with
 * the independently written numpy restatement oracle/weno5_numpy.py, which is
 * array-at-a-time and uses the reference's transpose/rotate formulation, while this file
 * sweeps lines in place (x lines contiguous, y lines strided, components permuted).
 *
 * Layout is the reference's: q[i + Nx*(j + Ny*c)], c = rho,Mx,My,Mz,Bx,By,Bz,S,E;
 * bxb/byb[i + Nx*j]; NBnd = 3 ghosts.  Indices in this file are 0-based (reference i-1).
 *
 * Quirk decisions: Q1 E-only; Q2 2D round-trips E through S every
 * stage, 1D keeps E; Q3 the stencil cell past the high end of a line continues the
 * boundary condition (clamp for outflow, wrap for periodic); Q11 flux-term order per
 * stage as in the reference; vmax sqrt arguments are clamped at 0 (reference: bare sqrt).
 *
 * Build: see oracle/Makefile.  The parity build uses -ffp-contract=off; the timed CPU
 * baseline build (-O3 -march=x86-64-v3 -fopenmp) is the same source.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NB 3

typedef struct {
    int nx, ny;          /* interior cells; ny = 1 for 1-D */
    double dx, dy, gam, cfl, eps;
    int bc_x, bc_y;      /* 0 outflow (reference), 1 periodic (addition) */
    int variant;         /* 0 = 2-D formulas, 1 = 1-D formulas (round-off differences, Q6/Q7) */
} orc_params;

/* ------------------------------------------------------------------ cell-centred quantities */

/* Weno5_2D.jl:1573-1606  q8 = [rho,M1,M2,M3,B1,B2,B3,E] -> u = [rho,v1,v2,v3,B1,B2,B3,P] */
static void cell_primitives(const double q[8], double gam, double u[8])
{
    u[0] = q[0];
    u[1] = q[1] / q[0];
    u[2] = q[2] / q[0];
    u[3] = q[3] / q[0];
    u[4] = q[4];
    u[5] = q[5];
    u[6] = q[6];
    double v2 = u[1] * u[1] + u[2] * u[2] + u[3] * u[3];
    double B2 = u[4] * u[4] + u[5] * u[5] + u[6] * u[6];
    u[7] = (gam - 1.0) * (q[7] - 0.5 * (q[0] * v2 + B2));
}

/* Weno5_2D.jl:1608-1667 */
static void cell_eigenvalues(const double u[8], double gam, double a[7])
{
    double B2 = u[4] * u[4] + u[5] * u[5] + u[6] * u[6];
    double cs2 = fmax(0.0, gam * fabs(u[7] / u[0]));
    double bnx2 = fmax(0.0, u[4] * u[4] / u[0]);
    double bbn2 = B2 / u[0];
    double s = bbn2 + cs2;
    double root = sqrt(fmax(0.0, s * s - 4.0 * bnx2 * cs2));
    double lf = sqrt(fmax(0.0, 0.5 * (bbn2 + cs2 + root)));
    double ls = sqrt(fmax(0.0, 0.5 * (bbn2 + cs2 - root)));
    double la = sqrt(bnx2);
    a[0] = u[1] - lf; a[1] = u[1] - la; a[2] = u[1] - ls; a[3] = u[1];
    a[4] = u[1] + ls; a[5] = u[1] + la; a[6] = u[1] + lf;
}

/* Weno5_2D.jl:1533-1555 */
static void cell_fluxes(const double q[8], const double u[8], double F[7])
{
    double pt = u[7] + 0.5 * (u[4] * u[4] + u[5] * u[5] + u[6] * u[6]);
    double vB = u[1] * u[4] + u[2] * u[5] + u[3] * u[6];
    F[0] = u[0] * u[1];
    F[1] = q[1] * u[1] + pt - u[4] * u[4];
    F[2] = q[2] * u[1] - u[4] * u[5];
    F[3] = q[3] * u[1] - u[4] * u[6];
    F[4] = u[5] * u[1] - u[4] * u[2];
    F[5] = u[6] * u[1] - u[4] * u[3];
    F[6] = (q[7] + pt) * u[1] - u[4] * vB;
}

/* ------------------------------------------------------------------ face eigen-system */

static double sgn(double x) { return (x > 0.0) - (x < 0.0); }

/* Weno5_2D.jl:1250-1531 (variant 0) / Weno5_1D.jl:784-1003 (variant 1).
 * L[m][c], R[c][m]; component order [rho,M1,M2,M3,B2,B3,E]. */
static void face_eigenvectors(const double uL[8], const double uR[8], double EL, double ER,
                              double gam, int variant, double L[7][7], double R[7][7])
{
    const double g0 = 1.0 - gam, g1 = (gam - 1.0) / 2.0, g2 = (gam - 2.0) / (gam - 1.0);
    double rho = 0.5 * (uL[0] + uR[0]);
    double r = sqrt(rho);
    double vx = 0.5 * (uL[1] + uR[1]), vy = 0.5 * (uL[2] + uR[2]), vz = 0.5 * (uL[3] + uR[3]);
    double Bx = 0.5 * (uL[4] + uR[4]), By = 0.5 * (uL[5] + uR[5]), Bz = 0.5 * (uL[6] + uR[6]);
    double E = 0.5 * (EL + ER);
    double B2 = Bx * Bx + By * By + Bz * Bz;
    double v2 = vx * vx + vy * vy + vz * vz;
    double pg = (gam - 1.0) * (E - 0.5 * (rho * v2 + B2));
    double bbn2 = B2 / rho;
    double cs2 = fmax(0.0, gam * fabs(pg / rho));
    double bnx2 = Bx * Bx / rho;
    double a = sqrt(cs2);
    double s_ = bbn2 + cs2;
    double root = sqrt(fmax(0.0, s_ * s_ - 4.0 * bnx2 * cs2));
    double cf = sqrt(fmax(0.0, (bbn2 + cs2 + root) / 2.0));
    double ca = sqrt(fmax(0.0, bnx2));
    double cs = sqrt(fmax(0.0, (bbn2 + cs2 - root) / 2.0));
    double Bt2 = By * By + Bz * Bz;
    double s = sgn(Bx);
    if (s == 0.0) s = 1.0;
    double bty = 1.0 / sqrt(2.0), btz = 1.0 / sqrt(2.0);
    int degenerate = variant == 0 ? (Bt2 <= 1e-30) : !(Bt2 >= 1e-30);
    if (!degenerate) { bty = By / sqrt(Bt2); btz = Bz / sqrt(Bt2); }
    double dl2 = cf * cf - cs * cs;
    double af = 1.0, as = 1.0;
    if (dl2 >= 1e-30) {
        if (variant == 0) {
            af = sqrt(fmax(0.0, cs2 - cs * cs) / dl2);
            as = sqrt(fmax(0.0, cf * cf - cs2) / dl2);
        } else {
            af = sqrt(fmax(0.0, cs2 - cs * cs)) / sqrt(dl2);
            as = sqrt(fmax(0.0, cf * cf - cs2)) / sqrt(dl2);
        }
    }
    double sig = bty * vy + btz * vz;

    L[0][0] = af * (g1 * v2 + cf * vx) - as * cs * sig * s;
    L[0][1] = af * (g0 * vx - cf);
    L[0][2] = g0 * af * vy + as * cs * bty * s;
    L[0][3] = g0 * af * vz + as * cs * btz * s;
    L[0][4] = g0 * af * By + a * as * bty * r;
    L[0][5] = g0 * af * Bz + a * as * btz * r;
    L[0][6] = -g0 * af;

    L[1][0] = btz * vy - bty * vz; L[1][1] = 0.0; L[1][2] = -btz; L[1][3] = bty;
    L[1][4] = -btz * s * r; L[1][5] = bty * s * r; L[1][6] = 0.0;

    L[2][0] = as * (g1 * v2 + cs * vx) + af * cf * sig * s;
    L[2][1] = g0 * as * vx - as * cs;
    L[2][2] = g0 * as * vy - af * cf * bty * s;
    L[2][3] = g0 * as * vz - af * cf * btz * s;
    L[2][4] = g0 * as * By - a * af * bty * r;
    L[2][5] = g0 * as * Bz - a * af * btz * r;
    L[2][6] = -g0 * as;

    L[3][0] = -cs2 / g0 - 0.5 * v2; L[3][1] = vx; L[3][2] = vy; L[3][3] = vz;
    L[3][4] = By; L[3][5] = Bz; L[3][6] = -1.0;

    L[4][0] = as * (g1 * v2 - cs * vx) - af * cf * sig * s;
    L[4][1] = as * (g0 * vx + cs);
    L[4][2] = g0 * as * vy + af * cf * bty * s;
    L[4][3] = g0 * as * vz + af * cf * btz * s;
    L[4][4] = g0 * as * By - a * af * bty * r;
    L[4][5] = g0 * as * Bz - a * af * btz * r;
    L[4][6] = -g0 * as;

    L[5][0] = btz * vy - bty * vz; L[5][1] = 0.0; L[5][2] = -btz; L[5][3] = bty;
    L[5][4] = btz * s * r; L[5][5] = -bty * s * r; L[5][6] = 0.0;

    L[6][0] = af * (g1 * v2 - cf * vx) + as * cs * sig * s;
    L[6][1] = af * (g0 * vx + cf);
    L[6][2] = g0 * af * vy - as * cs * bty * s;
    L[6][3] = g0 * af * vz - as * cs * btz * s;
    L[6][4] = g0 * af * By + a * as * bty * r;
    L[6][5] = g0 * af * Bz + a * as * btz * r;
    L[6][6] = -g0 * af;

    R[0][0] = af; R[1][0] = af * (vx - cf);
    R[2][0] = af * vy + as * cs * bty * s; R[3][0] = af * vz + as * cs * btz * s;
    R[4][0] = a * as * bty / r; R[5][0] = a * as * btz / r;
    R[6][0] = af * (cf * cf - cf * vx + 0.5 * v2 - g2 * cs2) + as * cs * sig * s;

    R[0][1] = 0.0; R[1][1] = 0.0; R[2][1] = -btz; R[3][1] = bty;
    R[4][1] = -btz * s / r; R[5][1] = bty * s / r; R[6][1] = bty * vz - btz * vy;

    R[0][2] = as; R[1][2] = as * (vx - cs);
    R[2][2] = as * vy - af * cf * bty * s; R[3][2] = as * vz - af * cf * btz * s;
    R[4][2] = -a * af * bty / r; R[5][2] = -a * af * btz / r;
    R[6][2] = as * (cs * cs - cs * vx + 0.5 * v2 - g2 * cs2) - af * cf * sig * s;

    R[0][3] = 1.0; R[1][3] = vx; R[2][3] = vy; R[3][3] = vz; R[4][3] = 0.0; R[5][3] = 0.0;
    R[6][3] = 0.5 * v2;

    R[0][4] = as; R[1][4] = as * (vx + cs);
    R[2][4] = as * vy + af * cf * bty * s; R[3][4] = as * vz + af * cf * btz * s;
    R[4][4] = -a * af * bty / r; R[5][4] = -a * af * btz / r;
    R[6][4] = as * (cs * cs + cs * vx + 0.5 * v2 - g2 * cs2) + af * cf * sig * s;

    R[0][5] = 0.0; R[1][5] = 0.0; R[2][5] = -btz; R[3][5] = bty;
    R[4][5] = btz * s / r; R[5][5] = -bty * s / r; R[6][5] = bty * vz - btz * vy;

    R[0][6] = af; R[1][6] = af * (vx + cf);
    R[2][6] = af * vy - as * cs * bty * s; R[3][6] = af * vz - as * cs * btz * s;
    R[4][6] = a * as * bty / r; R[5][6] = a * as * btz / r;
    R[6][6] = af * (cf * cf + cf * vx + 0.5 * v2 - g2 * cs2) - as * cs * sig * s;

    const double sc[7] = {0.5 / cs2, 0.5, 0.5 / cs2, -g0 / cs2, 0.5 / cs2, 0.5, 0.5 / cs2};
    for (int m = 0; m < 7; ++m)
        for (int c = 0; c < 7; ++c) L[m][c] *= sc[m];

    double sgnBt;
    if (variant == 0) { sgnBt = sgn(By); if (By == 0.0) sgnBt = sgn(Bz); }
    else { sgnBt = sgn(Bz); if (By != 0.0) sgnBt = sgn(Bx); }
    if (sgnBt == 0.0) sgnBt = 1.0;
    int m1 = (a >= ca) ? 2 : 0, m2 = (a >= ca) ? 4 : 6;
    for (int c = 0; c < 7; ++c) {
        L[m1][c] *= sgnBt; L[m2][c] *= sgnBt;
        R[c][m1] *= sgnBt; R[c][m2] *= sgnBt;
    }
}

/* ------------------------------------------------------------------ S branch
 * The entropy-variable twins that the shipped classical_RK4_step evaluates and discards (ES_Switch has its
 * criterion commented out); live in orc_rk4_step_2d_dual. */

/* Weno5_2D.jl:953-969  q8 = [rho,M1,M2,M3,B1,B2,B3,S] -> u = [rho,v1,v2,v3,B1,B2,B3,P(S)] */
static void cell_primitives_S(const double q[8], double gam, double u[8])
{
    u[0] = q[0];
    u[1] = q[1] / q[0];
    u[2] = q[2] / q[0];
    u[3] = q[3] / q[0];
    u[4] = q[4];
    u[5] = q[5];
    u[6] = q[6];
    u[7] = q[7] * pow(q[0], gam - 1.0);
}

/* Weno5_2D.jl:927-950 */
static void cell_fluxes_S(const double q[8], const double u[8], double F[7])
{
    double pt = u[7] + 0.5 * (u[4] * u[4] + u[5] * u[5] + u[6] * u[6]);
    F[0] = u[0] * u[1];
    F[1] = q[1] * u[1] + pt - u[4] * u[4];
    F[2] = q[2] * u[1] - u[4] * u[5];
    F[3] = q[3] * u[1] - u[4] * u[6];
    F[4] = u[5] * u[1] - u[4] * u[2];
    F[5] = u[6] * u[1] - u[4] * u[3];
    F[6] = q[7] * u[1];
}

/* compute_eigenvectors_S_vec, Weno5_2D.jl:640-924.  L[m][c], R[c][m]; components [rho,M1,M2,M3,B2,B3,S]. */
static void face_eigenvectors_S(const double uL[8], const double uR[8], double gam, double L[7][7], double R[7][7])
{
    const double g0 = 1.0 - gam, gS = (gam - 1.0) / gam;
    double drho = fabs(uL[0] - uR[0]);
    double rho = 0.5 * (uL[0] + uR[0]);
    if (!(drho <= 1e-12)) rho = pow(fabs(g0 * drho / (pow(uR[0], g0) - pow(uL[0], g0))), 1.0 / gam);   /* :657-667 */
    double r = sqrt(rho);
    double vx = 0.5 * (uL[1] + uR[1]), vy = 0.5 * (uL[2] + uR[2]), vz = 0.5 * (uL[3] + uR[3]);
    double Bx = 0.5 * (uL[4] + uR[4]), By = 0.5 * (uL[5] + uR[5]), Bz = 0.5 * (uL[6] + uR[6]);
    double pg = 0.5 * (uL[7] + uR[7]);
    double S = pg / pow(rho, gam - 1.0);
    double B2 = Bx * Bx + By * By + Bz * Bz;
    double bbn2 = B2 / rho;
    double cs2 = fmax(0.0, gam * fabs(pg / rho));
    double bnx2 = Bx * Bx / rho;
    double a = sqrt(cs2);
    double s_ = bbn2 + cs2;
    double root = sqrt(fmax(0.0, s_ * s_ - 4.0 * bnx2 * cs2));
    double cf = sqrt(fmax(0.0, (bbn2 + cs2 + root) / 2.0));
    double ca = sqrt(fmax(0.0, bnx2));
    double cs = sqrt(fmax(0.0, (bbn2 + cs2 - root) / 2.0));
    double Bt2 = By * By + Bz * Bz;
    double s = sgn(Bx);
    if (s == 0.0) s = 1.0;
    double bty = 1.0 / sqrt(2.0), btz = 1.0 / sqrt(2.0);
    if (!(Bt2 <= 1e-30)) { bty = By / sqrt(Bt2); btz = Bz / sqrt(Bt2); }
    double dl2 = cf * cf - cs * cs;
    double af = 1.0, as = 1.0;
    if (dl2 >= 1e-30) {
        af = sqrt(fmax(0.0, cs2 - cs * cs) / dl2);
        as = sqrt(fmax(0.0, cf * cf - cs2) / dl2);
    }
    double sig = bty * vy + btz * vz;
    double kap = rho / (gam * S), Sr = S / rho;

    L[0][0] = af * (gS * cs2 + cf * vx) - as * cs * sig * s; L[0][1] = -af * cf;
    L[0][2] = as * cs * bty * s; L[0][3] = as * cs * btz * s;
    L[0][4] = a * as * bty * r; L[0][5] = a * as * btz * r; L[0][6] = af * cs2 * kap;
    L[1][0] = btz * vy - bty * vz; L[1][1] = 0.0; L[1][2] = -btz; L[1][3] = bty;
    L[1][4] = -btz * s * r; L[1][5] = bty * s * r; L[1][6] = 0.0;
    L[2][0] = as * (gS * cs2 + cs * vx) + af * cf * sig * s; L[2][1] = -as * cs;
    L[2][2] = -af * cf * bty * s; L[2][3] = -af * cf * btz * s;
    L[2][4] = -a * af * bty * r; L[2][5] = -a * af * btz * r; L[2][6] = as * cs2 * kap;
    L[3][0] = 1.0 / gam; L[3][1] = 0.0; L[3][2] = 0.0; L[3][3] = 0.0; L[3][4] = 0.0; L[3][5] = 0.0; L[3][6] = -kap;
    L[4][0] = as * (gS * cs2 - cs * vx) - af * cf * sig * s; L[4][1] = as * cs;
    L[4][2] = af * cf * bty * s; L[4][3] = af * cf * btz * s;
    L[4][4] = -a * af * bty * r; L[4][5] = -a * af * btz * r; L[4][6] = as * cs2 * kap;
    L[5][0] = btz * vy - bty * vz; L[5][1] = 0.0; L[5][2] = -btz; L[5][3] = bty;
    L[5][4] = btz * s * r; L[5][5] = -bty * s * r; L[5][6] = 0.0;
    L[6][0] = af * (gS * cs2 - cf * vx) + as * cs * sig * s; L[6][1] = af * cf;
    L[6][2] = -as * cs * bty * s; L[6][3] = -as * cs * btz * s;
    L[6][4] = a * as * bty * r; L[6][5] = a * as * btz * r; L[6][6] = af * cs2 * kap;

    R[0][0] = af; R[1][0] = af * (vx - cf);
    R[2][0] = af * vy + as * cs * bty * s; R[3][0] = af * vz + as * cs * btz * s;
    R[4][0] = a * as * bty / r; R[5][0] = a * as * btz / r; R[6][0] = af * Sr;
    R[0][1] = 0.0; R[1][1] = 0.0; R[2][1] = -btz; R[3][1] = bty;
    R[4][1] = -btz * s / r; R[5][1] = bty * s / r; R[6][1] = 0.0;
    R[0][2] = as; R[1][2] = as * (vx - cs);
    R[2][2] = as * vy - af * cf * bty * s; R[3][2] = as * vz - af * cf * btz * s;
    R[4][2] = -a * af * bty / r; R[5][2] = -a * af * btz / r; R[6][2] = as * Sr;
    R[0][3] = 1.0; R[1][3] = vx; R[2][3] = vy; R[3][3] = vz; R[4][3] = 0.0; R[5][3] = 0.0; R[6][3] = g0 * Sr;
    R[0][4] = as; R[1][4] = as * (vx + cs);
    R[2][4] = as * vy + af * cf * bty * s; R[3][4] = as * vz + af * cf * btz * s;
    R[4][4] = -a * af * bty / r; R[5][4] = -a * af * btz / r; R[6][4] = as * Sr;
    R[0][5] = 0.0; R[1][5] = 0.0; R[2][5] = -btz; R[3][5] = bty;
    R[4][5] = btz * s / r; R[5][5] = -bty * s / r; R[6][5] = 0.0;
    R[0][6] = af; R[1][6] = af * (vx + cf);
    R[2][6] = af * vy - as * cs * bty * s; R[3][6] = af * vz - as * cs * btz * s;
    R[4][6] = a * as * bty / r; R[5][6] = a * as * btz / r; R[6][6] = af * Sr;

    const double sc[7] = {0.5 / cs2, 0.5, 0.5 / cs2, 1.0, 0.5 / cs2, 0.5, 0.5 / cs2};   /* row 4 unscaled, :872 */
    for (int m = 0; m < 7; ++m)
        for (int c = 0; c < 7; ++c) L[m][c] *= sc[m];
    double sgnBt = sgn(By);
    if (By == 0.0) sgnBt = sgn(Bz);
    if (sgnBt == 0.0) sgnBt = 1.0;
    int m1 = (a >= ca) ? 2 : 0, m2 = (a >= ca) ? 4 : 6;
    for (int c = 0; c < 7; ++c) {
        L[m1][c] *= sgnBt; L[m2][c] *= sgnBt;
        R[c][m1] *= sgnBt; R[c][m2] *= sgnBt;
    }
}

/* one side of the Jiang & Wu WENO correction, Weno5_2D.jl:1841-1853 */
static double weno_phi(double a, double b, double c, double d, double eps)
{
    double t;
    t = a - b;       double IS0 = 13.0 * (t * t);
    t = a - 3.0 * b; IS0 += 3.0 * (t * t);
    t = b - c;       double IS1 = 13.0 * (t * t);
    t = b + c;       IS1 += 3.0 * (t * t);
    t = c - d;       double IS2 = 13.0 * (t * t);
    t = 3.0 * c - d; IS2 += 3.0 * (t * t);
    double e0 = eps + IS0, e1 = eps + IS1, e2 = eps + IS2;
    double al0 = 1.0 / (e0 * e0), al1 = 6.0 / (e1 * e1), al2 = 3.0 / (e2 * e2);
    double om0 = al0 / (al0 + al1 + al2);
    double om2 = al2 / (al0 + al1 + al2);
    return om0 * (a - 2.0 * b + c) / 3.0 + (om2 - 0.5) * (b - 2.0 * c + d) / 6.0;
}

/* Numerical flux of one face from its 6-cell stencil, Weno5_2D.jl:1810-1888.
 * Fk[k], qk[k] (k = 0..5) are the flux / conserved 7-vectors of cells f-2 .. f+3;
 * aL, aR the eigenvalues of cells f and f+1. */
static void face_flux(const double *const Fk[6], const double *const qk[6],
                      const double aL[7], const double aR[7],
                      double L[7][7], double R[7][7], double eps, double fhat[7])
{
    double Fs[7];
    for (int m = 0; m < 7; ++m) {
        double w[6], v[6];
        for (int k = 0; k < 6; ++k) {
            double sw = L[m][0] * Fk[k][0], sv = L[m][0] * qk[k][0];
            for (int c = 1; c < 7; ++c) { sw += L[m][c] * Fk[k][c]; sv += L[m][c] * qk[k][c]; }
            w[k] = sw; v[k] = sv;
        }
        double first = (-w[1] + 7.0 * w[2] + 7.0 * w[3] - w[4]) / 12.0;
        double dw[5], dv[5];
        for (int k = 0; k < 5; ++k) { dw[k] = w[k + 1] - w[k]; dv[k] = v[k + 1] - v[k]; }
        double amax = fmax(fabs(aL[m]), fabs(aR[m]));
        double second = weno_phi(0.5 * (dw[0] + amax * dv[0]), 0.5 * (dw[1] + amax * dv[1]),
                                 0.5 * (dw[2] + amax * dv[2]), 0.5 * (dw[3] + amax * dv[3]), eps);
        double third = weno_phi(0.5 * (dw[4] - amax * dv[4]), 0.5 * (dw[3] - amax * dv[3]),
                                0.5 * (dw[2] - amax * dv[2]), 0.5 * (dw[1] - amax * dv[1]), eps);
        Fs[m] = first - second + third;
    }
    for (int c = 0; c < 7; ++c) {
        double s = Fs[0] * R[c][0];
        for (int m = 1; m < 7; ++m) s += Fs[m] * R[c][m];
        fhat[c] = s;
    }
}

/* ------------------------------------------------------------------ one line of a sweep
 *
 * line[n][8] holds [rho,M1,M2,M3,B1,B2,B3,E] of the n cells of a line, with direction 1 the
 * sweep direction.  Computes fhat[f][7] for faces f = NB-1 .. n-NB (0-based; the reference's
 * 1-based NBnd .. Nq-NBnd+1, Weno5_2D.jl:993-1008) and the CT seeds
 *   bs2[f] = fhat[4] + 0.5 (B1 v2|f + B1 v2|f+1),  bs3[f] = fhat[5] + 0.5 (B1 v3|f + B1 v3|f+1).
 * The stencil cell f+3 = n of the last face continues the boundary condition (Q3).
 * work: 4 arrays of n*8 doubles.
 */
static void sweep_line(int n, const double (*line)[8], double gam, double eps, int bc, int variant, int branch,
                       double (*fhat)[7], double *bs2, double *bs3, double *work)
{
    double (*u)[8] = (double (*)[8])work;
    double (*a)[8] = (double (*)[8])(work + 8 * (size_t)n);
    double (*F)[8] = (double (*)[8])(work + 16 * (size_t)n);
    double (*q7)[8] = (double (*)[8])(work + 24 * (size_t)n);
    for (int i = 0; i < n; ++i) {
        if (branch == 0) { cell_primitives(line[i], gam, u[i]); cell_fluxes(line[i], u[i], F[i]); }
        else { cell_primitives_S(line[i], gam, u[i]); cell_fluxes_S(line[i], u[i], F[i]); }
        cell_eigenvalues(u[i], gam, a[i]);
        q7[i][0] = line[i][0]; q7[i][1] = line[i][1]; q7[i][2] = line[i][2]; q7[i][3] = line[i][3];
        q7[i][4] = line[i][5]; q7[i][5] = line[i][6]; q7[i][6] = line[i][7];
    }
    const int ni = n - 2 * NB;
    for (int f = NB - 1; f <= n - NB; ++f) {
        double L[7][7], R[7][7];
        if (branch == 0) face_eigenvectors(u[f], u[f + 1], line[f][7], line[f + 1][7], gam, variant, L, R);
        else face_eigenvectors_S(u[f], u[f + 1], gam, L, R);
        const double *Fk[6], *qk[6];
        for (int k = 0; k < 6; ++k) {
            int c = f - 2 + k;
            if (c >= n) c = (bc == 1) ? c - ni : n - 1;
            Fk[k] = F[c]; qk[k] = q7[c];
        }
        face_flux(Fk, qk, a[f], a[f + 1], L, R, eps, fhat[f]);
        bs2[f] = fhat[f][4] + 0.5 * (u[f][4] * u[f][2] + u[f + 1][4] * u[f + 1][2]);
        bs3[f] = fhat[f][5] + 0.5 * (u[f][4] * u[f][3] + u[f + 1][4] * u[f + 1][3]);
    }
}

/* ------------------------------------------------------------------ 2-D grid services */

#define IDX(i, j) ((size_t)(i) + (size_t)Nx * (size_t)(j))

static void fill_axis_x(double *f, int Nx, int Ny, int bc)
{
    int n = Nx - 2 * NB;
    for (int j = 0; j < Ny; ++j)
        for (int g = 0; g < NB; ++g) {
            if (bc == 1) {
                f[IDX(g, j)] = f[IDX(n + g, j)];
                f[IDX(Nx - NB + g, j)] = f[IDX(NB + g, j)];
            } else {
                f[IDX(g, j)] = f[IDX(NB, j)];
                f[IDX(Nx - NB + g, j)] = f[IDX(Nx - NB - 1, j)];
            }
        }
}

static void fill_axis_y(double *f, int Nx, int Ny, int bc)
{
    int n = Ny - 2 * NB;
    for (int g = 0; g < NB; ++g)
        for (int i = 0; i < Nx; ++i) {
            if (bc == 1) {
                f[IDX(i, g)] = f[IDX(i, n + g)];
                f[IDX(i, Ny - NB + g)] = f[IDX(i, NB + g)];
            } else {
                f[IDX(i, g)] = f[IDX(i, NB)];
                f[IDX(i, Ny - NB + g)] = f[IDX(i, Ny - NB - 1)];
            }
        }
}

/* Weno5_2D.jl:1686-1702: x first, then y (so corners take the y copy of x-filled rows) */
static void boundaries_q(double *q, int ncomp, int Nx, int Ny, const orc_params *P)
{
    for (int c = 0; c < ncomp; ++c) {
        double *f = q + (size_t)c * Nx * Ny;
        fill_axis_x(f, Nx, Ny, P->bc_x);
        if (Ny > 1) fill_axis_y(f, Nx, Ny, P->bc_y);
    }
}

/* Weno5_2D.jl:1704-1729 for one flux / face array */
static void boundaries_flux(double *f, int Nx, int Ny, const orc_params *P)
{
    if (P->bc_x == 1) fill_axis_x(f, Nx, Ny, 1);
    else for (int j = 0; j < Ny; ++j) { f[IDX(1, j)] = f[IDX(2, j)]; f[IDX(0, j)] = f[IDX(2, j)]; }
    if (P->bc_y == 1) fill_axis_y(f, Nx, Ny, 1);
    else for (int i = 0; i < Nx; ++i) { f[IDX(i, 1)] = f[IDX(i, 2)]; f[IDX(i, 0)] = f[IDX(i, 2)]; }
    if (P->bc_x != 1)
        for (int j = 0; j < Ny; ++j) { f[IDX(Nx - 2, j)] = f[IDX(Nx - 3, j)]; f[IDX(Nx - 1, j)] = f[IDX(Nx - 3, j)]; }
    if (P->bc_y != 1)
        for (int i = 0; i < Nx; ++i) { f[IDX(i, Ny - 2)] = f[IDX(i, Ny - 3)]; f[IDX(i, Ny - 1)] = f[IDX(i, Ny - 3)]; }
    if (P->bc_x != 1) for (int j = 0; j < Ny; ++j) { f[IDX(0, j)] = 0.0; f[IDX(Nx - 1, j)] = 0.0; }
    if (P->bc_y != 1) for (int i = 0; i < Nx; ++i) { f[IDX(i, 0)] = 0.0; f[IDX(i, Ny - 1)] = 0.0; }
}

/* Weno5_2D.jl:1770-1793 on planes; qE planes: rho,Mx,My,Mz,Bx,By,Bz,E */
static double e2s(double rho, double mx, double my, double mz, double bx, double by, double bz,
                  double E, double gam)
{
    double mom2 = mx * mx + my * my + mz * mz;
    double B2 = bx * bx + by * by + bz * bz;
    return (E - 0.5 * mom2 / rho - B2 / 2) * (gam - 1) / pow(rho, gam - 1);
}

static double s2e(double rho, double mx, double my, double mz, double bx, double by, double bz,
                  double S, double gam)
{
    double mom2 = mx * mx + my * my + mz * mz;
    double B2 = bx * bx + by * by + bz * bz;
    return pow(rho, gam - 1) * S / (gam - 1) + B2 / 2 + 0.5 * mom2 / rho;
}

/* ------------------------------------------------------------------ directional sweeps on the grid
 * q: 8 planes [rho,Mx,My,Mz,Bx,By,Bz,E]; dq: 8 planes (zeroed here); bs: one plane. */

static void sweep_x_br(const orc_params *P, const double *q, int branch, double *dq, double *fsy)
{
    const int Nx = P->nx + 2 * NB, Ny = (P->ny > 1) ? P->ny + 2 * NB : 1;
    const size_t pl = (size_t)Nx * Ny;
    memset(dq, 0, 8 * pl * sizeof(double));
    memset(fsy, 0, pl * sizeof(double));
    const int j0 = (Ny > 1) ? NB - 1 : 0, j1 = (Ny > 1) ? Ny - NB : 0;
#pragma omp parallel
    {
        double (*line)[8] = malloc(sizeof(double[8]) * Nx);
        double (*fh)[7] = malloc(sizeof(double[7]) * Nx);
        double *b2 = malloc(sizeof(double) * Nx), *b3 = malloc(sizeof(double) * Nx);
        double *work = malloc(sizeof(double) * 32 * Nx);
#pragma omp for schedule(static)
        for (int j = j0; j <= j1; ++j) {
            for (int i = 0; i < Nx; ++i)
                for (int c = 0; c < 8; ++c) line[i][c] = q[c * pl + IDX(i, j)];
            memset(fh, 0, sizeof(double[7]) * Nx);
            sweep_line(Nx, line, P->gam, P->eps, P->bc_x, P->variant, branch, fh, b2, b3, work);
            static const int out[7] = {0, 1, 2, 3, 5, 6, 7};
            for (int i = NB - 1; i <= Nx - NB; ++i) {
                for (int c = 0; c < 7; ++c) dq[out[c] * pl + IDX(i, j)] = fh[i][c] - fh[i - 1][c];
                fsy[IDX(i, j)] = b2[i];
            }
        }
        free(line); free(fh); free(b2); free(b3); free(work);
    }
}

/* y sweep without a transposed copy: the line runs along j and the vector components are
 * permuted (x,y,z) -> (y,z,x) as rotate_state does (Weno5_2D.jl:1897-1927). */
static void sweep_y_br(const orc_params *P, const double *q, int branch, double *dq, double *gsx)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB;
    const size_t pl = (size_t)Nx * Ny;
    memset(dq, 0, 8 * pl * sizeof(double));
    memset(gsx, 0, pl * sizeof(double));
    static const int in[8] = {0, 2, 3, 1, 5, 6, 4, 7};  /* line comp c reads plane in[c] */
    /* fhat comps [rho,M1,M2,M3,B2,B3,E] of the rotated frame -> planes My,Mz,Mx,Bz,Bx */
    static const int out[7] = {0, 2, 3, 1, 6, 4, 7};
#pragma omp parallel
    {
        double (*line)[8] = malloc(sizeof(double[8]) * Ny);
        double (*fh)[7] = malloc(sizeof(double[7]) * Ny);
        double *b2 = malloc(sizeof(double) * Ny), *b3 = malloc(sizeof(double) * Ny);
        double *work = malloc(sizeof(double) * 32 * Ny);
#pragma omp for schedule(static)
        for (int i = NB - 1; i <= Nx - NB; ++i) {
            for (int j = 0; j < Ny; ++j)
                for (int c = 0; c < 8; ++c) line[j][c] = q[in[c] * pl + IDX(i, j)];
            memset(fh, 0, sizeof(double[7]) * Ny);
            sweep_line(Ny, line, P->gam, P->eps, P->bc_y, P->variant, branch, fh, b2, b3, work);
            for (int j = NB - 1; j <= Ny - NB; ++j) {
                for (int c = 0; c < 7; ++c) dq[out[c] * pl + IDX(i, j)] = fh[j][c] - fh[j - 1][c];
                gsx[IDX(i, j)] = b3[j];
            }
        }
        free(line); free(fh); free(b2); free(b3); free(work);
    }
}

static void sweep_x(const orc_params *P, const double *q, double *dq, double *fsy) { sweep_x_br(P, q, 0, dq, fsy); }
static void sweep_y(const orc_params *P, const double *q, double *dq, double *gsx) { sweep_y_br(P, q, 0, dq, gsx); }

/* Weno5_2D.jl:319-342 + the face-B update lines (:150-151 etc.) + :344-356 */
static void ct_update(const orc_params *P, const double *fsy, const double *gsx,
                      const double *bbx, const double *bby, double c, double dt,
                      double *bxn, double *byn, double *Bx, double *By, double *Ox)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB;
    const size_t pl = (size_t)Nx * Ny;
    memset(Ox, 0, pl * sizeof(double));
#pragma omp parallel for schedule(static)
    for (int j = 1; j < Ny - 1; ++j)
        for (int i = 1; i < Nx - 1; ++i)
            Ox[IDX(i, j)] = 0.5 * (gsx[IDX(i, j)] + gsx[IDX(i + 1, j)] - fsy[IDX(i, j)] - fsy[IDX(i, j + 1)]);
    const double cy = c * dt / P->dy, cx = c * dt / P->dx;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < Ny; ++j)
        for (int i = 0; i < Nx; ++i) {
            int inner = (i >= 1 && i < Nx - 1 && j >= 1 && j < Ny - 1);
            double o = Ox[IDX(i, j)];
            double oi = inner ? Ox[IDX(i - 1, j)] : 0.0;
            double oj = inner ? Ox[IDX(i, j - 1)] : 0.0;
            bxn[IDX(i, j)] = bbx[IDX(i, j)] - cy * (o - oj);
            byn[IDX(i, j)] = bby[IDX(i, j)] - cx * (oi - o);
        }
    memset(Bx, 0, pl * sizeof(double));
    memset(By, 0, pl * sizeof(double));
#pragma omp parallel for schedule(static)
    for (int j = NB; j <= Ny - 3; ++j)
        for (int i = NB; i <= Nx - 3; ++i) {
            Bx[IDX(i, j)] = (-bxn[IDX(i - 2, j)] + 9.0 * bxn[IDX(i - 1, j)] + 9.0 * bxn[IDX(i, j)] - bxn[IDX(i + 1, j)]) / 16.0;
            By[IDX(i, j)] = (-byn[IDX(i, j - 2)] + 9.0 * byn[IDX(i, j - 1)] + 9.0 * byn[IDX(i, j)] - byn[IDX(i, j + 1)]) / 16.0;
        }
}

typedef struct { double *dFx, *dFy, *fsy, *gsx, *Ox, *qE; } stage_work;

/* One E-branch stage of classical_RK4_step (Weno5_2D.jl:134-180), including what survives
 * of ES_Switch (:1734-1768): S = E2S(state), E = S2E(state with S).
 * qk: 9 planes (input state); base: 8 planes [..,E]; out: 9 planes. */
static void stage_2d(const orc_params *P, const double *qk, const double *base, const double *bbx,
                     const double *bby, double c, double dt, int stage, double *out, double *bxn,
                     double *byn, stage_work *W)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB;
    const size_t pl = (size_t)Nx * Ny;
    memcpy(W->qE, qk, 7 * pl * sizeof(double));
    memcpy(W->qE + 7 * pl, qk + 8 * pl, pl * sizeof(double));
    sweep_x(P, W->qE, W->dFx, W->fsy);
    sweep_y(P, W->qE, W->dFy, W->gsx);
    double *qn = W->qE;   /* reuse: the state copy is no longer needed */
    const double hx = dt / P->dx, hy = dt / P->dy;
    for (int cmp = 0; cmp < 8; ++cmp) {
        const double *b = base + cmp * pl, *fx = W->dFx + cmp * pl, *fy = W->dFy + cmp * pl;
        double *o = qn + cmp * pl;
#pragma omp parallel for schedule(static)
        for (size_t k = 0; k < pl; ++k) {
            if (stage <= 2)      o[k] = b[k] - 1.0 / 2 * dt / P->dy * fy[k] - 1.0 / 2 * dt / P->dx * fx[k];
            else if (stage == 3) o[k] = b[k] - hx * fx[k] - hy * fy[k];
            else                 o[k] = b[k] + (-1.0 / 6 * dt / P->dx * fx[k] - 1.0 / 6 * dt / P->dy * fy[k]);
        }
    }
    boundaries_q(qn, 8, Nx, Ny, P);
    boundaries_flux(W->fsy, Nx, Ny, P);
    boundaries_flux(W->gsx, Nx, Ny, P);
    ct_update(P, W->fsy, W->gsx, bbx, bby, c, dt, bxn, byn, qn + 4 * pl, qn + 5 * pl, W->Ox);
    boundaries_q(qn, 8, Nx, Ny, P);
    memcpy(out, qn, 7 * pl * sizeof(double));
    double *S = out + 7 * pl, *E = out + 8 * pl;
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < pl; ++k) {
        S[k] = e2s(qn[k], qn[pl + k], qn[2 * pl + k], qn[3 * pl + k], qn[4 * pl + k], qn[5 * pl + k],
                   qn[6 * pl + k], qn[7 * pl + k], P->gam);
        E[k] = s2e(qn[k], qn[pl + k], qn[2 * pl + k], qn[3 * pl + k], qn[4 * pl + k], qn[5 * pl + k],
                   qn[6 * pl + k], S[k], P->gam);
    }
    boundaries_q(out, 9, Nx, Ny, P);
}

/* ------------------------------------------------------------------ exported entry points */

/* classical_RK4_step(q0,bxb0,byb0,dt) -> (q4,bxb4,byb4), Weno5_2D.jl:125-317 */
int orc_rk4_step_2d(const orc_params *P, const double *q0, const double *bxb0, const double *byb0,
                    double dt, double *q4, double *bxb4, double *byb4)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB;
    const size_t pl = (size_t)Nx * Ny;
    stage_work W;
    W.dFx = malloc(8 * pl * sizeof(double)); W.dFy = malloc(8 * pl * sizeof(double));
    W.fsy = malloc(pl * sizeof(double)); W.gsx = malloc(pl * sizeof(double));
    W.Ox = malloc(pl * sizeof(double)); W.qE = malloc(8 * pl * sizeof(double));
    double *q0E = malloc(8 * pl * sizeof(double)), *base = malloc(8 * pl * sizeof(double));
    double *q1 = malloc(9 * pl * sizeof(double)), *q2 = malloc(9 * pl * sizeof(double)),
           *q3 = malloc(9 * pl * sizeof(double));
    double *bx[4], *by[4];
    for (int k = 1; k < 4; ++k) { bx[k] = malloc(pl * sizeof(double)); by[k] = malloc(pl * sizeof(double)); }
    double *bbx = malloc(pl * sizeof(double)), *bby = malloc(pl * sizeof(double));
    if (!W.dFx || !W.dFy || !W.qE || !q0E || !base || !q1 || !q2 || !q3 || !bby) return -1;
    memcpy(q0E, q0, 7 * pl * sizeof(double));
    memcpy(q0E + 7 * pl, q0 + 8 * pl, pl * sizeof(double));

    stage_2d(P, q0, q0E, bxb0, byb0, 0.5, dt, 1, q1, bx[1], by[1], &W);
    stage_2d(P, q1, q0E, bxb0, byb0, 0.5, dt, 2, q2, bx[2], by[2], &W);
    stage_2d(P, q2, q0E, bxb0, byb0, 1.0, dt, 3, q3, bx[3], by[3], &W);
    for (int c = 0; c < 8; ++c) {
        int src = (c == 7) ? 8 : c;
        for (size_t k = 0; k < pl; ++k)
            base[c * pl + k] = 1.0 / 3 * (-q0E[c * pl + k] + q1[src * pl + k] + 2 * q2[src * pl + k] + q3[src * pl + k]);
    }
    for (size_t k = 0; k < pl; ++k) {
        bbx[k] = 1.0 / 3 * (-bxb0[k] + bx[1][k] + 2 * bx[2][k] + bx[3][k]);
        bby[k] = 1.0 / 3 * (-byb0[k] + by[1][k] + 2 * by[2][k] + by[3][k]);
    }
    stage_2d(P, q3, base, bbx, bby, 1.0 / 6, dt, 4, q4, bxb4, byb4, &W);
    boundaries_flux(bxb4, Nx, Ny, P);   /* Weno5_2D.jl:313 */
    boundaries_flux(byb4, Nx, Ny, P);

    free(W.dFx); free(W.dFy); free(W.fsy); free(W.gsx); free(W.Ox); free(W.qE);
    free(q0E); free(base); free(q1); free(q2); free(q3); free(bbx); free(bby);
    for (int k = 1; k < 4; ++k) { free(bx[k]); free(by[k]); }
    return 0;
}

/* ------------------------------------------------------------------ dual-energy step
 * classical_RK4_step with both branches live and ES_Switch (Weno5_2D.jl:1734-1768) using the criterion the
 * reference wrote and commented out (:1740): dS = |E2S(q_E) - S(q_S)| >= threshold -> the E state, and the CT seed
 * fluxes of that cell and of its low-side neighbour faces from the E sweep (:1754-1758); else the S state. */
typedef struct { double *dFx, *dFy, *fsy[2], *gsx[2], *Ox, *cand[2], *bxt, *byt; unsigned char *bad; } dual_work;

static void stage_2d_dual(const orc_params *P, const double *qk, const double *baseE, const double *baseS,
                          const double *bbx, const double *bby, double c, double dt, int stage, double threshold,
                          double *out, double *bxn, double *byn, long *nbad, dual_work *W)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB;
    const size_t pl = (size_t)Nx * Ny;
    const double hx = dt / P->dx, hy = dt / P->dy;
    for (int br = 0; br < 2; ++br) {
        double *qb = W->cand[br];
        memcpy(qb, qk, 7 * pl * sizeof(double));
        memcpy(qb + 7 * pl, qk + (br == 0 ? 8 : 7) * pl, pl * sizeof(double));
        sweep_x_br(P, qb, br, W->dFx, W->fsy[br]);
        sweep_y_br(P, qb, br, W->dFy, W->gsx[br]);
        const double *base = br == 0 ? baseE : baseS;
        for (int cmp = 0; cmp < 8; ++cmp) {
            const double *b = base + cmp * pl, *fx = W->dFx + cmp * pl, *fy = W->dFy + cmp * pl;
            double *o = qb + cmp * pl;
#pragma omp parallel for schedule(static)
            for (size_t k = 0; k < pl; ++k) {
                /* term order of the reference: E stages 1,2 take y first (:145,:190), everything else x first */
                if (stage <= 2 && br == 0) o[k] = b[k] - 1.0 / 2 * dt / P->dy * fy[k] - 1.0 / 2 * dt / P->dx * fx[k];
                else if (stage <= 2)       o[k] = b[k] - 1.0 / 2 * dt / P->dx * fx[k] - 1.0 / 2 * dt / P->dy * fy[k];
                else if (stage == 3)       o[k] = b[k] - hx * fx[k] - hy * fy[k];
                else                       o[k] = b[k] + (-1.0 / 6 * dt / P->dx * fx[k] - 1.0 / 6 * dt / P->dy * fy[k]);
            }
        }
        boundaries_q(qb, 8, Nx, Ny, P);
        boundaries_flux(W->fsy[br], Nx, Ny, P);
        boundaries_flux(W->gsx[br], Nx, Ny, P);
        ct_update(P, W->fsy[br], W->gsx[br], bbx, bby, c, dt, W->bxt, W->byt, qb + 4 * pl, qb + 5 * pl, W->Ox);
        boundaries_q(qb, 8, Nx, Ny, P);
    }
    const double *qE = W->cand[0], *qS = W->cand[1];
    long count = 0;
#pragma omp parallel for schedule(static) reduction(+ : count)
    for (int j = 0; j < Ny; ++j)
        for (int i = 0; i < Nx; ++i) {
            const size_t k = IDX(i, j);
            const double SE = e2s(qE[k], qE[pl + k], qE[2 * pl + k], qE[3 * pl + k], qE[4 * pl + k], qE[5 * pl + k],
                                  qE[6 * pl + k], qE[7 * pl + k], P->gam);
            const int b = i >= 1 && j >= 1 && fabs(SE - qS[7 * pl + k]) >= threshold;      /* the loop starts at 2 */
            W->bad[k] = (unsigned char)b;
            const double *src = b ? qE : qS;
            for (int cmp = 0; cmp < 7; ++cmp) out[cmp * pl + k] = src[cmp * pl + k];
            out[7 * pl + k] = b ? SE : qS[7 * pl + k];
            out[8 * pl + k] = s2e(out[k], out[pl + k], out[2 * pl + k], out[3 * pl + k], out[4 * pl + k],
                                  out[5 * pl + k], out[6 * pl + k], out[7 * pl + k], P->gam);       /* :1765 */
            if (b && i >= NB && i < Nx - NB && j >= NB && j < Ny - NB) ++count;
        }
    *nbad = count;
    double *fsy = W->fsy[1], *gsx = W->gsx[1];
#pragma omp parallel for schedule(static)
    for (int j = 0; j < Ny; ++j)
        for (int i = 0; i < Nx; ++i) {
            const size_t k = IDX(i, j);
            const int here = W->bad[k];
            if (here || (i + 1 < Nx && W->bad[IDX(i + 1, j)])) fsy[k] = W->fsy[0][k];
            if (here || (j + 1 < Ny && W->bad[IDX(i, j + 1)])) gsx[k] = W->gsx[0][k];
        }
    boundaries_q(out, 9, Nx, Ny, P);
    boundaries_flux(fsy, Nx, Ny, P);
    boundaries_flux(gsx, Nx, Ny, P);
    ct_update(P, fsy, gsx, bbx, bby, c, dt, bxn, byn, out + 4 * pl, out + 5 * pl, W->Ox);
    boundaries_q(out, 9, Nx, Ny, P);
}

/* classical_RK4_step with the live switch; nbad[4] = interior cells that took the E state, per stage */
int orc_rk4_step_2d_dual(const orc_params *P, const double *q0, const double *bxb0, const double *byb0, double dt,
                         double threshold, double *q4, double *bxb4, double *byb4, long nbad[4])
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB;
    const size_t pl = (size_t)Nx * Ny;
    dual_work W;
    W.dFx = malloc(8 * pl * sizeof(double)); W.dFy = malloc(8 * pl * sizeof(double));
    for (int b = 0; b < 2; ++b) {
        W.fsy[b] = malloc(pl * sizeof(double)); W.gsx[b] = malloc(pl * sizeof(double));
        W.cand[b] = malloc(8 * pl * sizeof(double));
    }
    W.Ox = malloc(pl * sizeof(double)); W.bxt = malloc(pl * sizeof(double)); W.byt = malloc(pl * sizeof(double));
    W.bad = malloc(pl);
    double *q0E = malloc(8 * pl * sizeof(double)), *q0S = malloc(8 * pl * sizeof(double));
    double *baseE = malloc(8 * pl * sizeof(double)), *baseS = malloc(8 * pl * sizeof(double));
    double *q[4];
    double *bx[4], *by[4];
    for (int k = 1; k < 4; ++k) {
        q[k] = malloc(9 * pl * sizeof(double)); bx[k] = malloc(pl * sizeof(double)); by[k] = malloc(pl * sizeof(double));
    }
    double *bbx = malloc(pl * sizeof(double)), *bby = malloc(pl * sizeof(double));
    memcpy(q0E, q0, 7 * pl * sizeof(double));
    memcpy(q0E + 7 * pl, q0 + 8 * pl, pl * sizeof(double));
    memcpy(q0S, q0, 8 * pl * sizeof(double));

    stage_2d_dual(P, q0, q0E, q0S, bxb0, byb0, 0.5, dt, 1, threshold, q[1], bx[1], by[1], &nbad[0], &W);
    stage_2d_dual(P, q[1], q0E, q0S, bxb0, byb0, 0.5, dt, 2, threshold, q[2], bx[2], by[2], &nbad[1], &W);
    stage_2d_dual(P, q[2], q0E, q0S, bxb0, byb0, 1.0, dt, 3, threshold, q[3], bx[3], by[3], &nbad[2], &W);
    for (int c = 0; c < 8; ++c) {
        const int srcE = (c == 7) ? 8 : c;
        for (size_t k = 0; k < pl; ++k) {
            baseE[c * pl + k] = 1.0 / 3 * (-q0E[c * pl + k] + q[1][srcE * pl + k] + 2 * q[2][srcE * pl + k] + q[3][srcE * pl + k]);
            baseS[c * pl + k] = 1.0 / 3 * (-q0S[c * pl + k] + q[1][c * pl + k] + 2 * q[2][c * pl + k] + q[3][c * pl + k]);
        }
    }
    for (size_t k = 0; k < pl; ++k) {
        bbx[k] = 1.0 / 3 * (-bxb0[k] + bx[1][k] + 2 * bx[2][k] + bx[3][k]);
        bby[k] = 1.0 / 3 * (-byb0[k] + by[1][k] + 2 * by[2][k] + by[3][k]);
    }
    stage_2d_dual(P, q[3], baseE, baseS, bbx, bby, 1.0 / 6, dt, 4, threshold, q4, bxb4, byb4, &nbad[3], &W);
    boundaries_flux(bxb4, Nx, Ny, P);
    boundaries_flux(byb4, Nx, Ny, P);

    free(W.dFx); free(W.dFy); free(W.Ox); free(W.bxt); free(W.byt); free(W.bad);
    for (int b = 0; b < 2; ++b) { free(W.fsy[b]); free(W.gsx[b]); free(W.cand[b]); }
    free(q0E); free(q0S); free(baseE); free(baseS); free(bbx); free(bby);
    for (int k = 1; k < 4; ++k) { free(q[k]); free(bx[k]); free(by[k]); }
    return 0;
}

/* weno5_flux_difference_S (Weno5_2D.jl:361), one direction: q8 = [rho,Mx,My,Mz,Bx,By,Bz,S] */
int orc_sweep_2d_S(const orc_params *P, const double *q8, int dir, double *dq, double *bs)
{
    if (dir == 0) sweep_x_br(P, q8, 1, dq, bs); else sweep_y_br(P, q8, 1, dq, bs);
    return 0;
}

/* timestep / compute_global_vmax, Weno5_2D.jl:1930-1984; q has 9 planes (S in plane 7) */
int orc_timestep_2d(const orc_params *P, const double *q, double *dt, double *vxmax, double *vymax)
{
    const int Nx = P->nx + 2 * NB, Ny = (P->ny > 1) ? P->ny + 2 * NB : 1;
    const size_t pl = (size_t)Nx * Ny;
    const int j0 = (Ny > 1) ? NB : 0, j1 = (Ny > 1) ? Ny - NB - 1 : 0;
    double mx = -INFINITY, my = -INFINITY;
#pragma omp parallel for reduction(max : mx, my) schedule(static)
    for (int j = j0; j <= j1; ++j)
        for (int i = NB; i <= Nx - NB - 1; ++i) {
            size_t a = IDX(i, j), b = IDX(i + 1, j);
            double rho = (q[a] + q[b]) / 2;
            double vx = (q[pl + a] + q[pl + b]) / 2 / rho;
            double vy = (q[2 * pl + a] + q[2 * pl + b]) / 2 / rho;
            double Bx = (q[4 * pl + a] + q[4 * pl + b]) / 2;
            double By = (q[5 * pl + a] + q[5 * pl + b]) / 2;
            double Bz = (q[6 * pl + a] + q[6 * pl + b]) / 2;
            double S = (q[7 * pl + a] + q[7 * pl + b]) / 2;
            double pres = S * pow(rho, P->gam - 1);
            double cs2 = P->gam * fabs(pres / rho);
            double B2 = Bx * Bx + By * By + Bz * Bz;
            double bbn2 = B2 / rho, bnx2 = Bx * Bx / rho, bny2 = By * By / rho;
            double s = bbn2 + cs2;
            double rootx = sqrt(fmax(0.0, s * s - 4 * bnx2 * cs2));
            double rooty = sqrt(fmax(0.0, s * s - 4 * bny2 * cs2));
            double lfx = sqrt(0.5 * (bbn2 + cs2 + rootx));
            double lfy = sqrt(0.5 * (bbn2 + cs2 + rooty));
            mx = fmax(mx, fabs(vx) + lfx);
            my = fmax(my, fabs(vy) + lfy);
        }
    *vxmax = mx; *vymax = my;
    if (Ny > 1) {
        double dtx = P->dx / mx, dty = P->dy / my;
        *dt = P->cfl / (1 / dtx + 1 / dty);
    } else {
        *dt = P->cfl * P->dx / mx;   /* Weno5_1D.jl:55 */
    }
    return 0;
}

/* compute_divB, Weno5_2D.jl:1986-1999 */
int orc_divb(const orc_params *P, const double *bxb, const double *byb, double *divb)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB;
    memset(divb, 0, sizeof(double) * (size_t)Nx * Ny);
    for (int j = NB - 1; j <= Ny - NB - 1; ++j)
        for (int i = NB - 1; i <= Nx - NB - 1; ++i)
            divb[IDX(i, j)] = (bxb[IDX(i, j)] - bxb[IDX(i - 1, j)]) / P->dx + (byb[IDX(i, j)] - byb[IDX(i, j - 1)]) / P->dy;
    return 0;
}

/* interpolateB_center2face, Weno5_2D.jl:1669-1682; q has >= 6 planes */
int orc_center2face(const orc_params *P, const double *q, double *bxb, double *byb)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB;
    const size_t pl = (size_t)Nx * Ny;
    memset(bxb, 0, pl * sizeof(double));
    memset(byb, 0, pl * sizeof(double));
    const double *Bx = q + 4 * pl, *By = q + 5 * pl;
    for (int j = 1; j <= Ny - 3; ++j)
        for (int i = 1; i <= Nx - 3; ++i) {
            bxb[IDX(i, j)] = (-Bx[IDX(i - 1, j)] + 9 * Bx[IDX(i, j)] + 9 * Bx[IDX(i + 1, j)] - Bx[IDX(i + 2, j)]) / 16;
            byb[IDX(i, j)] = (-By[IDX(i, j - 1)] + 9 * By[IDX(i, j)] + 9 * By[IDX(i, j + 1)] - By[IDX(i, j + 2)]) / 16;
        }
    return 0;
}

/* boundaries!(q, a, b), Weno5_2D.jl:1684-1732; any of the pointers may be NULL */
int orc_boundaries_2d(const orc_params *P, double *q9, double *a, double *b)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB;
    if (q9) boundaries_q(q9, 9, Nx, Ny, P);
    if (a) boundaries_flux(a, Nx, Ny, P);
    if (b) boundaries_flux(b, Nx, Ny, P);
    return 0;
}

/* one directional sweep for unit tests: weno5_flux_difference_E (Weno5_2D.jl:973-1018)
 * applied along x (dir 0) or along y with rotate_state folded in (dir 1).
 * q8: 8 planes [rho,Mx,My,Mz,Bx,By,Bz,E]; dq: 8 planes; bs: fsy (dir 0) or gsx (dir 1). */
int orc_sweep_2d(const orc_params *P, const double *q8, int dir, double *dq, double *bs)
{
    if (dir == 0) sweep_x(P, q8, dq, bs); else sweep_y(P, q8, dq, bs);
    return 0;
}

/* 1-D: classical_RK4_step(q,dt) -> q4, Weno5_1D.jl:152-236 with the E-only ES_Switch
 * (:238-260: q[1:7] = q_E, q[8] = E2S(q_E), q[9] = E).  q: (N,9) column-major. */
int orc_rk4_step_1d(const orc_params *Pin, const double *q0, double dt, double *q4)
{
    orc_params P = *Pin;
    P.ny = 1; P.variant = 1;
    const int N = P.nx + 2 * NB;
    const size_t pl = (size_t)N;
    double *qE[4], *Q = malloc(8 * pl * sizeof(double)), *fs = malloc(pl * sizeof(double));
    double *cur = malloc(9 * pl * sizeof(double)), *nxt = malloc(8 * pl * sizeof(double));
    for (int k = 0; k < 4; ++k) qE[k] = malloc(8 * pl * sizeof(double));
    memcpy(cur, q0, 9 * pl * sizeof(double));
    for (int st = 1; st <= 4; ++st) {
        double *e = qE[st - 1];
        memcpy(e, cur, 7 * pl * sizeof(double));
        memcpy(e + 7 * pl, cur + 8 * pl, pl * sizeof(double));
        sweep_x(&P, e, Q, fs);
        /* the reference forms Q only on iMin..iMax (Weno5_1D.jl:766-775) */
        for (int c = 0; c < 8; ++c) { Q[c * pl + NB - 1] = 0.0; Q[c * pl + N - NB] = 0.0; }
        for (int c = 0; c < 8; ++c)
            for (size_t i = 0; i < pl; ++i) {
                size_t k = c * pl + i;
                if (st <= 2)      nxt[k] = qE[0][k] - 1.0 / 2 * dt / P.dx * Q[k];
                else if (st == 3) nxt[k] = qE[0][k] - dt / P.dx * Q[k];
                else nxt[k] = 1.0 / 3 * (-qE[0][k] + qE[1][k] + 2 * qE[2][k] + qE[3][k] - 1.0 / 2 * dt / P.dx * Q[k]);
            }
        memcpy(cur, nxt, 7 * pl * sizeof(double));
        for (size_t i = 0; i < pl; ++i) {
            cur[7 * pl + i] = e2s(nxt[i], nxt[pl + i], nxt[2 * pl + i], nxt[3 * pl + i], nxt[4 * pl + i],
                                  nxt[5 * pl + i], nxt[6 * pl + i], nxt[7 * pl + i], P.gam);
            cur[8 * pl + i] = nxt[7 * pl + i];
        }
        for (int c = 0; c < 9; ++c) fill_axis_x(cur + c * pl, N, 1, P.bc_x);   /* Weno5_1D.jl:1076-1090 */
    }
    memcpy(q4, cur, 9 * pl * sizeof(double));
    for (int k = 0; k < 4; ++k) free(qE[k]);
    free(Q); free(fs); free(cur); free(nxt);
    return 0;
}

/* driver loop used for the CPU baseline: nsteps of (timestep, RK4 step) in place */
int orc_advance_2d(const orc_params *P, double *q, double *bxb, double *byb, int nsteps, double *t)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB;
    const size_t pl = (size_t)Nx * Ny;
    double *q4 = malloc(9 * pl * sizeof(double)), *bx4 = malloc(pl * sizeof(double)), *by4 = malloc(pl * sizeof(double));
    for (int s = 0; s < nsteps; ++s) {
        double dt, vx, vy;
        orc_timestep_2d(P, q, &dt, &vx, &vy);
        if (orc_rk4_step_2d(P, q, bxb, byb, dt, q4, bx4, by4)) return -1;
        memcpy(q, q4, 9 * pl * sizeof(double));
        memcpy(bxb, bx4, pl * sizeof(double));
        memcpy(byb, by4, pl * sizeof(double));
        *t += dt;
    }
    free(q4); free(bx4); free(by4);
    return 0;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ================================================================== 3-D extension
 * The reference has no 3-D solver; its classical_RK4_step docstring (Weno5_2D.jl:113-123) states the design:
 * the scheme is dimensionally unsplit, "the Roe solver / Weno functions are inherently the same for 1D, 2D and
 * 3D" and every direction is swept with the same functions on a rotated state.  This section extends
 * classical_RK4_step accordingly: a third sweep along z in the frame (z,x,y) (the cyclic continuation of
 * rotate_state, :1897-1927), flux differences of rho, M, E from all three sweeps, and constrained transport of
 * all three face fields from the edge-centred EMFs built of the six transport fluxes (the 3-D form of
 * corner_fluxes :319-342 and of the face updates :150-151, Ryu et al. 1998 eqs. 3.11-3.13):
 *   x sweep: f_y = bs2 (By vx), f_z = bs3 (Bz vx)   at faces (i+1/2,j,k)
 *   y sweep: g_z = bs2 (Bz vy), g_x = bs3 (Bx vy)   at faces (i,j+1/2,k)
 *   z sweep: h_x = bs2 (Bx vz), h_y = bs3 (By vz)   at faces (i,j,k+1/2)
 *   Ox[i,j,k] = (h_y[i,j,k] + h_y[i,j+1,k] - g_z[i,j,k] - g_z[i,j,k+1]) / 2      edge (i,j+1/2,k+1/2)
 *   Oy[i,j,k] = (f_z[i,j,k] + f_z[i,j,k+1] - h_x[i,j,k] - h_x[i+1,j,k]) / 2      edge (i+1/2,j,k+1/2)
 *   Oz[i,j,k] = (g_x[i,j,k] + g_x[i+1,j,k] - f_y[i,j,k] - f_y[i,j+1,k]) / 2      edge (i+1/2,j+1/2,k)  (= the 2-D Ox)
 *   bxb -= c dt ((Oz - Oz[j-1])/dy - (Oy - Oy[k-1])/dz),  byb -= c dt ((Ox - Ox[k-1])/dz - (Oz - Oz[i-1])/dx),
 *   bzb -= c dt ((Oy - Oy[i-1])/dx - (Ox - Ox[j-1])/dy).
 * Pinned by reduction: a state that does not depend on one coordinate (and has no field or velocity along it)
 * reproduces orc_rk4_step_2d in each of the three orientations (tests/test_oracle3d.py).
 * Layout: q[i + Nx*(j + Ny*(k + Nz*c))], c as in 2-D; bxb, byb, bzb [Nx,Ny,Nz]. */

typedef struct {
    int nx, ny, nz;
    double dx, dy, dz, gam, cfl, eps;
    int bc_x, bc_y, bc_z;
} orc_params3;

#define ID3(i, j, k) ((size_t)(i) + (size_t)Nx * ((size_t)(j) + (size_t)Ny * (size_t)(k)))

/* ghost fill of one volume along one axis (the 3-D continuation of fill_axis_x/y) */
static void fill_axis3(double *f, int Nx, int Ny, int Nz, int axis, int bc)
{
    const int N[3] = {Nx, Ny, Nz};
    const size_t st[3] = {1, (size_t)Nx, (size_t)Nx * Ny};
    const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
    const int n = N[axis] - 2 * NB;
#pragma omp parallel for schedule(static)
    for (int v = 0; v < N[a2]; ++v)
        for (int u = 0; u < N[a1]; ++u) {
            double *l = f + st[a1] * u + st[a2] * v;
            for (int g = 0; g < NB; ++g) {
                if (bc == 1) {
                    l[st[axis] * g] = l[st[axis] * (n + g)];
                    l[st[axis] * (N[axis] - NB + g)] = l[st[axis] * (NB + g)];
                } else {
                    l[st[axis] * g] = l[st[axis] * NB];
                    l[st[axis] * (N[axis] - NB + g)] = l[st[axis] * (N[axis] - NB - 1)];
                }
            }
        }
}

static void boundaries_q3(double *q, int ncomp, const orc_params3 *P)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB, Nz = P->nz + 2 * NB;
    const size_t vol = (size_t)Nx * Ny * Nz;
    for (int c = 0; c < ncomp; ++c) {
        fill_axis3(q + c * vol, Nx, Ny, Nz, 0, P->bc_x);
        fill_axis3(q + c * vol, Nx, Ny, Nz, 1, P->bc_y);
        fill_axis3(q + c * vol, Nx, Ny, Nz, 2, P->bc_z);
    }
}

/* the flux / face-array rules of boundaries! (Weno5_2D.jl:1704-1729) per axis: low side, high side, rim */
static void boundaries_flux3(double *f, const orc_params3 *P)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB, Nz = P->nz + 2 * NB;
    const int N[3] = {Nx, Ny, Nz}, bc[3] = {P->bc_x, P->bc_y, P->bc_z};
    const size_t st[3] = {1, (size_t)Nx, (size_t)Nx * Ny};
    for (int pass = 0; pass < 3; ++pass)          /* 0: low sides, 1: high sides, 2: zero rims */
        for (int axis = 0; axis < 3; ++axis) {
            const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
            if (bc[axis] == 1) { if (pass == 0) fill_axis3(f, Nx, Ny, Nz, axis, 1); continue; }
            for (int v = 0; v < N[a2]; ++v)
                for (int u = 0; u < N[a1]; ++u) {
                    double *l = f + st[a1] * u + st[a2] * v;
                    const size_t s = st[axis];
                    const int n = N[axis];
                    if (pass == 0) { l[s * 1] = l[s * 2]; l[0] = l[s * 2]; }
                    else if (pass == 1) { l[s * (n - 2)] = l[s * (n - 3)]; l[s * (n - 1)] = l[s * (n - 3)]; }
                    else { l[0] = 0.0; l[s * (n - 1)] = 0.0; }
                }
        }
}

/* One directional sweep of the E branch on the volume.  q: 8 volumes [rho,Mx,My,Mz,Bx,By,Bz,E]; dq: 5 volumes
 * [rho,Mx,My,Mz,E] (zeroed here; fhat differences, not yet divided by the cell size); bsA, bsB: the two transport
 * fluxes of the direction (table above).  Lines one ghost cell beyond the interior are swept too, like the 2-D
 * sweeps do for fsy/gsx. */
static void sweep3(const orc_params3 *P, const double *q, int dir, double *dq, double *bsA, double *bsB)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB, Nz = P->nz + 2 * NB;
    const size_t vol = (size_t)Nx * Ny * Nz;
    const int N[3] = {Nx, Ny, Nz}, bc[3] = {P->bc_x, P->bc_y, P->bc_z};
    const size_t st[3] = {1, (size_t)Nx, (size_t)Nx * Ny};
    /* line component c reads volume in[dir][c]: (x,y,z) -> (y,z,x) -> (z,x,y) */
    static const int in[3][8] = {{0, 1, 2, 3, 4, 5, 6, 7}, {0, 2, 3, 1, 5, 6, 4, 7}, {0, 3, 1, 2, 6, 4, 5, 7}};
    /* fhat comps [rho,M1,M2,M3,.,.,E] -> dq volume (0 rho, 1 Mx, 2 My, 3 Mz, 4 E) */
    static const int out[3][7] = {{0, 1, 2, 3, -1, -1, 4}, {0, 2, 3, 1, -1, -1, 4}, {0, 3, 1, 2, -1, -1, 4}};
    const int a1 = (dir + 1) % 3, a2 = (dir + 2) % 3, n = N[dir];
    memset(dq, 0, 5 * vol * sizeof(double));
    memset(bsA, 0, vol * sizeof(double));
    memset(bsB, 0, vol * sizeof(double));
#pragma omp parallel
    {
        double (*line)[8] = malloc(sizeof(double[8]) * n);
        double (*fh)[7] = malloc(sizeof(double[7]) * n);
        double *b2 = malloc(sizeof(double) * n), *b3 = malloc(sizeof(double) * n);
        double *work = malloc(sizeof(double) * 32 * n);
#pragma omp for schedule(static) collapse(2)
        for (int v = NB - 1; v <= N[a2] - NB; ++v)
            for (int u = NB - 1; u <= N[a1] - NB; ++u) {
                const size_t o = st[a1] * u + st[a2] * v;
                for (int s = 0; s < n; ++s)
                    for (int c = 0; c < 8; ++c) line[s][c] = q[in[dir][c] * vol + o + st[dir] * s];
                memset(fh, 0, sizeof(double[7]) * n);
                sweep_line(n, line, P->gam, P->eps, bc[dir], 0, 0, fh, b2, b3, work);
                for (int s = NB - 1; s <= n - NB; ++s) {
                    for (int c = 0; c < 7; ++c)
                        if (out[dir][c] >= 0) dq[out[dir][c] * vol + o + st[dir] * s] = fh[s][c] - fh[s - 1][c];
                    bsA[o + st[dir] * s] = b2[s];
                    bsB[o + st[dir] * s] = b3[s];
                }
            }
        free(line); free(fh); free(b2); free(b3); free(work);
    }
}

typedef struct { double *dF[3], *fl[6], *O[3], *qE; } stage_work3;

static void ct_update3(const orc_params3 *P, stage_work3 *W, const double *const bb[3], double c, double dt,
                       double *const bn[3], double *Bc /* 3 volumes */)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB, Nz = P->nz + 2 * NB;
    const size_t vol = (size_t)Nx * Ny * Nz;
    const double *fy = W->fl[0], *fz = W->fl[1], *gz = W->fl[2], *gx = W->fl[3], *hx = W->fl[4], *hy = W->fl[5];
    double *Ox = W->O[0], *Oy = W->O[1], *Oz = W->O[2];
    for (int m = 0; m < 3; ++m) memset(W->O[m], 0, vol * sizeof(double));
#pragma omp parallel for schedule(static)
    for (int k = 1; k < Nz - 1; ++k)
        for (int j = 1; j < Ny - 1; ++j)
            for (int i = 1; i < Nx - 1; ++i) {
                Ox[ID3(i, j, k)] = 0.5 * (hy[ID3(i, j, k)] + hy[ID3(i, j + 1, k)] - gz[ID3(i, j, k)] - gz[ID3(i, j, k + 1)]);
                Oy[ID3(i, j, k)] = 0.5 * (fz[ID3(i, j, k)] + fz[ID3(i, j, k + 1)] - hx[ID3(i, j, k)] - hx[ID3(i + 1, j, k)]);
                Oz[ID3(i, j, k)] = 0.5 * (gx[ID3(i, j, k)] + gx[ID3(i + 1, j, k)] - fy[ID3(i, j, k)] - fy[ID3(i, j + 1, k)]);
            }
    const double cx = c * dt / P->dx, cy = c * dt / P->dy, cz = c * dt / P->dz;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < Nz; ++k)
        for (int j = 0; j < Ny; ++j)
            for (int i = 0; i < Nx; ++i) {
                const int inner = i >= 1 && i < Nx - 1 && j >= 1 && j < Ny - 1 && k >= 1 && k < Nz - 1;
                const size_t a = ID3(i, j, k);
                double ox = Ox[a], oy = Oy[a], oz = Oz[a];
                double ox_j = inner ? Ox[ID3(i, j - 1, k)] : 0.0, ox_k = inner ? Ox[ID3(i, j, k - 1)] : 0.0;
                double oy_i = inner ? Oy[ID3(i - 1, j, k)] : 0.0, oy_k = inner ? Oy[ID3(i, j, k - 1)] : 0.0;
                double oz_i = inner ? Oz[ID3(i - 1, j, k)] : 0.0, oz_j = inner ? Oz[ID3(i, j - 1, k)] : 0.0;
                bn[0][a] = bb[0][a] - cy * (oz - oz_j) + cz * (oy - oy_k);
                bn[1][a] = bb[1][a] - cz * (ox - ox_k) + cx * (oz - oz_i);
                bn[2][a] = bb[2][a] - cx * (oy - oy_i) + cy * (ox - ox_j);
            }
    memset(Bc, 0, 3 * vol * sizeof(double));
#pragma omp parallel for schedule(static)
    for (int k = NB; k <= Nz - 3; ++k)
        for (int j = NB; j <= Ny - 3; ++j)
            for (int i = NB; i <= Nx - 3; ++i) {
                const double *x = bn[0], *y = bn[1], *z = bn[2];
                Bc[ID3(i, j, k)] = (-x[ID3(i - 2, j, k)] + 9.0 * x[ID3(i - 1, j, k)] + 9.0 * x[ID3(i, j, k)] - x[ID3(i + 1, j, k)]) / 16.0;
                Bc[vol + ID3(i, j, k)] = (-y[ID3(i, j - 2, k)] + 9.0 * y[ID3(i, j - 1, k)] + 9.0 * y[ID3(i, j, k)] - y[ID3(i, j + 1, k)]) / 16.0;
                Bc[2 * vol + ID3(i, j, k)] = (-z[ID3(i, j, k - 2)] + 9.0 * z[ID3(i, j, k - 1)] + 9.0 * z[ID3(i, j, k)] - z[ID3(i, j, k + 1)]) / 16.0;
            }
}

/* one stage: qk 9 volumes, base 8 volumes [..,E], out 9 volumes; roundtrip: E <- S2E(E2S(E)) as the 2-D reference */
static void stage_3d(const orc_params3 *P, const double *qk, const double *base, const double *const bb[3], double c,
                     double dt, int roundtrip, double *out, double *const bn[3], stage_work3 *W)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB, Nz = P->nz + 2 * NB;
    const size_t vol = (size_t)Nx * Ny * Nz;
    memcpy(W->qE, qk, 7 * vol * sizeof(double));
    memcpy(W->qE + 7 * vol, qk + 8 * vol, vol * sizeof(double));
    sweep3(P, W->qE, 0, W->dF[0], W->fl[0], W->fl[1]);
    sweep3(P, W->qE, 1, W->dF[1], W->fl[2], W->fl[3]);
    sweep3(P, W->qE, 2, W->dF[2], W->fl[4], W->fl[5]);
    double *qn = W->qE;
    static const int dst[5] = {0, 1, 2, 3, 7};
    const double ix = 1.0 / P->dx, iy = 1.0 / P->dy, iz = 1.0 / P->dz, cdt = c * dt;
    for (int m = 0; m < 5; ++m) {
        const double *b = base + dst[m] * vol, *fx = W->dF[0] + m * vol, *fy = W->dF[1] + m * vol, *fz = W->dF[2] + m * vol;
        double *o = qn + dst[m] * vol;
#pragma omp parallel for schedule(static)
        for (size_t a = 0; a < vol; ++a) o[a] = b[a] - cdt * ((fx[a] * ix + fy[a] * iy) + fz[a] * iz);
    }
    boundaries_q3(qn, 8, P);
    for (int m = 0; m < 6; ++m) boundaries_flux3(W->fl[m], P);
    ct_update3(P, W, bb, c, dt, bn, qn + 4 * vol);
    boundaries_q3(qn, 8, P);
    memcpy(out, qn, 7 * vol * sizeof(double));
    double *S = out + 7 * vol, *E = out + 8 * vol;
#pragma omp parallel for schedule(static)
    for (size_t a = 0; a < vol; ++a) {
        S[a] = e2s(qn[a], qn[vol + a], qn[2 * vol + a], qn[3 * vol + a], qn[4 * vol + a], qn[5 * vol + a],
                   qn[6 * vol + a], qn[7 * vol + a], P->gam);
        E[a] = roundtrip ? s2e(qn[a], qn[vol + a], qn[2 * vol + a], qn[3 * vol + a], qn[4 * vol + a], qn[5 * vol + a],
                               qn[6 * vol + a], S[a], P->gam)
                         : qn[7 * vol + a];
    }
    boundaries_q3(out, 9, P);
}

int orc_rk4_step_3d(const orc_params3 *P, const double *q0, const double *bxb0, const double *byb0, const double *bzb0,
                    double dt, int roundtrip, double *q4, double *bxb4, double *byb4, double *bzb4)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB, Nz = P->nz + 2 * NB;
    const size_t vol = (size_t)Nx * Ny * Nz;
    stage_work3 W;
    for (int m = 0; m < 3; ++m) { W.dF[m] = malloc(5 * vol * sizeof(double)); W.O[m] = malloc(vol * sizeof(double)); }
    for (int m = 0; m < 6; ++m) W.fl[m] = malloc(vol * sizeof(double));
    W.qE = malloc(8 * vol * sizeof(double));
    double *q0E = malloc(8 * vol * sizeof(double)), *base = malloc(8 * vol * sizeof(double));
    double *qs[4] = {NULL, malloc(9 * vol * sizeof(double)), malloc(9 * vol * sizeof(double)), malloc(9 * vol * sizeof(double))};
    double *bs[4][3], *bbase[3];
    for (int s = 1; s < 4; ++s) for (int m = 0; m < 3; ++m) bs[s][m] = malloc(vol * sizeof(double));
    for (int m = 0; m < 3; ++m) bbase[m] = malloc(vol * sizeof(double));
    if (!W.qE || !q0E || !base || !qs[1] || !qs[2] || !qs[3] || !bbase[2]) return -1;
    memcpy(q0E, q0, 7 * vol * sizeof(double));
    memcpy(q0E + 7 * vol, q0 + 8 * vol, vol * sizeof(double));
    const double *const b0[3] = {bxb0, byb0, bzb0};

    stage_3d(P, q0, q0E, b0, 0.5, dt, roundtrip, qs[1], bs[1], &W);
    stage_3d(P, qs[1], q0E, b0, 0.5, dt, roundtrip, qs[2], bs[2], &W);
    stage_3d(P, qs[2], q0E, b0, 1.0, dt, roundtrip, qs[3], bs[3], &W);
    for (int c = 0; c < 8; ++c) {
        const int src = (c == 7) ? 8 : c;
        for (size_t a = 0; a < vol; ++a)
            base[c * vol + a] = 1.0 / 3 * (-q0E[c * vol + a] + qs[1][src * vol + a] + 2 * qs[2][src * vol + a] + qs[3][src * vol + a]);
    }
    for (int m = 0; m < 3; ++m)
        for (size_t a = 0; a < vol; ++a)
            bbase[m][a] = 1.0 / 3 * (-b0[m][a] + bs[1][m][a] + 2 * bs[2][m][a] + bs[3][m][a]);
    double *const b4[3] = {bxb4, byb4, bzb4};
    const double *const bbc[3] = {bbase[0], bbase[1], bbase[2]};
    stage_3d(P, qs[3], base, bbc, 1.0 / 6, dt, roundtrip, q4, b4, &W);
    for (int m = 0; m < 3; ++m) boundaries_flux3(b4[m], P);

    for (int m = 0; m < 3; ++m) { free(W.dF[m]); free(W.O[m]); free(bbase[m]); }
    for (int m = 0; m < 6; ++m) free(W.fl[m]);
    free(W.qE); free(q0E); free(base);
    for (int s = 1; s < 4; ++s) { free(qs[s]); for (int m = 0; m < 3; ++m) free(bs[s][m]); }
    return 0;
}

/* timestep in 3-D: the 2-D rule (x-face averages, S-based pressure, Weno5_2D.jl:1942-1984) with a third direction */
int orc_timestep_3d(const orc_params3 *P, const double *q, double *dt, double vmax[3])
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB, Nz = P->nz + 2 * NB;
    const size_t vol = (size_t)Nx * Ny * Nz;
    double mx = -INFINITY, my = -INFINITY, mz = -INFINITY;
#pragma omp parallel for reduction(max : mx, my, mz) schedule(static)
    for (int k = NB; k <= Nz - NB - 1; ++k)
        for (int j = NB; j <= Ny - NB - 1; ++j)
            for (int i = NB; i <= Nx - NB - 1; ++i) {
                size_t a = ID3(i, j, k), b = ID3(i + 1, j, k);
                double rho = (q[a] + q[b]) / 2;
                double vx = (q[vol + a] + q[vol + b]) / 2 / rho;
                double vy = (q[2 * vol + a] + q[2 * vol + b]) / 2 / rho;
                double vz = (q[3 * vol + a] + q[3 * vol + b]) / 2 / rho;
                double Bx = (q[4 * vol + a] + q[4 * vol + b]) / 2;
                double By = (q[5 * vol + a] + q[5 * vol + b]) / 2;
                double Bz = (q[6 * vol + a] + q[6 * vol + b]) / 2;
                double S = (q[7 * vol + a] + q[7 * vol + b]) / 2;
                double pres = S * pow(rho, P->gam - 1);
                double cs2 = P->gam * fabs(pres / rho);
                double B2 = Bx * Bx + By * By + Bz * Bz;
                double bbn2 = B2 / rho, s = bbn2 + cs2;
                double rx = sqrt(fmax(0.0, s * s - 4 * (Bx * Bx / rho) * cs2));
                double ry = sqrt(fmax(0.0, s * s - 4 * (By * By / rho) * cs2));
                double rz = sqrt(fmax(0.0, s * s - 4 * (Bz * Bz / rho) * cs2));
                mx = fmax(mx, fabs(vx) + sqrt(0.5 * (s + rx)));
                my = fmax(my, fabs(vy) + sqrt(0.5 * (s + ry)));
                mz = fmax(mz, fabs(vz) + sqrt(0.5 * (s + rz)));
            }
    vmax[0] = mx; vmax[1] = my; vmax[2] = mz;
    *dt = P->cfl / (mx / P->dx + my / P->dy + mz / P->dz);
    return 0;
}

int orc_divb_3d(const orc_params3 *P, const double *bxb, const double *byb, const double *bzb, double *divb)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB, Nz = P->nz + 2 * NB;
    memset(divb, 0, sizeof(double) * (size_t)Nx * Ny * Nz);
    for (int k = NB - 1; k <= Nz - NB - 1; ++k)
        for (int j = NB - 1; j <= Ny - NB - 1; ++j)
            for (int i = NB - 1; i <= Nx - NB - 1; ++i)
                divb[ID3(i, j, k)] = (bxb[ID3(i, j, k)] - bxb[ID3(i - 1, j, k)]) / P->dx
                                   + (byb[ID3(i, j, k)] - byb[ID3(i, j - 1, k)]) / P->dy
                                   + (bzb[ID3(i, j, k)] - bzb[ID3(i, j, k - 1)]) / P->dz;
    return 0;
}

/* interpolateB_center2face (Weno5_2D.jl:1669-1682) with the z face field */
int orc_center2face_3d(const orc_params3 *P, const double *q, double *bxb, double *byb, double *bzb)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB, Nz = P->nz + 2 * NB;
    const size_t vol = (size_t)Nx * Ny * Nz;
    memset(bxb, 0, vol * sizeof(double)); memset(byb, 0, vol * sizeof(double)); memset(bzb, 0, vol * sizeof(double));
    const double *Bx = q + 4 * vol, *By = q + 5 * vol, *Bz = q + 6 * vol;
    for (int k = 1; k <= Nz - 3; ++k)
        for (int j = 1; j <= Ny - 3; ++j)
            for (int i = 1; i <= Nx - 3; ++i) {
                bxb[ID3(i, j, k)] = (-Bx[ID3(i - 1, j, k)] + 9 * Bx[ID3(i, j, k)] + 9 * Bx[ID3(i + 1, j, k)] - Bx[ID3(i + 2, j, k)]) / 16;
                byb[ID3(i, j, k)] = (-By[ID3(i, j - 1, k)] + 9 * By[ID3(i, j, k)] + 9 * By[ID3(i, j + 1, k)] - By[ID3(i, j + 2, k)]) / 16;
                bzb[ID3(i, j, k)] = (-Bz[ID3(i, j, k - 1)] + 9 * Bz[ID3(i, j, k)] + 9 * Bz[ID3(i, j, k + 1)] - Bz[ID3(i, j, k + 2)]) / 16;
            }
    return 0;
}

int orc_boundaries_3d(const orc_params3 *P, double *q9, double *a, double *b, double *c)
{
    if (q9) boundaries_q3(q9, 9, P);
    if (a) boundaries_flux3(a, P);
    if (b) boundaries_flux3(b, P);
    if (c) boundaries_flux3(c, P);
    return 0;
}

int orc_advance_3d(const orc_params3 *P, double *q, double *bxb, double *byb, double *bzb, int nsteps, int roundtrip,
                   double *t)
{
    const int Nx = P->nx + 2 * NB, Ny = P->ny + 2 * NB, Nz = P->nz + 2 * NB;
    const size_t vol = (size_t)Nx * Ny * Nz;
    double *q4 = malloc(9 * vol * sizeof(double)), *b4[3];
    for (int m = 0; m < 3; ++m) b4[m] = malloc(vol * sizeof(double));
    for (int s = 0; s < nsteps; ++s) {
        double dt, vm[3];
        orc_timestep_3d(P, q, &dt, vm);
        if (orc_rk4_step_3d(P, q, bxb, byb, bzb, dt, roundtrip, q4, b4[0], b4[1], b4[2])) return -1;
        memcpy(q, q4, 9 * vol * sizeof(double));
        memcpy(bxb, b4[0], vol * sizeof(double)); memcpy(byb, b4[1], vol * sizeof(double)); memcpy(bzb, b4[2], vol * sizeof(double));
        *t += dt;
    }
    free(q4); for (int m = 0; m < 3; ++m) free(b4[m]);
    return 0;
}
