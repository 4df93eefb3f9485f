// Per-cell and per-face arithmetic of the WENO5 MHD sweep, written for the B200 FP64 pipe.
//
// What the reference does per face (Weno5_2D.jl:1250-1531 eigenvectors, :1795-1895 WENO):
// build dense 7x7 matrices L and R (98 doubles), project the 6-cell stencil of F and q
// with 588 multiply-adds, run WENO5 on the 7 characteristic fields, multiply by R.
// This file computes the same quantities without ever forming L or R: the rows of L and
// the columns of R are linear combinations of a few shared inner products (see
// DESIGN.md "structured eigen-system"), which brings the projection from 49 to 32 FMAs per
// stencil vector and the back-projection from 49 to ~35.  The WENO weights use one
// reciprocal per side instead of five divisions (same rational function of the
// smoothness indicators).  All of it is algebraically identical to the reference; the
// difference is round-off (checked against the CPU oracle, tolerance 1e-10 on smooth
// problems per BASELINE.json).
//
// The functions are __host__ __device__ so that tests/ can compile them for the host and
// compare against the oracle without a GPU; the product only ever runs them on the device.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define W5_HD __host__ __device__ __forceinline__
#else
#define W5_HD inline
#endif

namespace w5 {

// quantities stored per cell in shared memory by the cell phase
enum CellQ {
    CQ_F0 = 0,            // F[0..6]  physical flux, Weno5_2D.jl:1533-1555
    CQ_Q0 = 7,            // q7[0..6] = rho, M1, M2, M3, B2, B3, E
    CQ_V1 = 14, CQ_V2 = 15, CQ_V3 = 16,
    CQ_B1 = 17,
    CQ_LF = 18, CQ_LA = 19, CQ_LS = 20,
    CQ_COUNT = 21
};

struct CellOut {
    double F[7];
    double v1, v2, v3;
    double lf, la, ls;
};

// 1/x for normal x > 0, branch-free: hardware seed (measured relative error 9.4e-7 on B200,
// scripts/microbench/seed_accuracy.cu) and one cubic step r(1 + e + e^2), e = 1 - x r, which leaves
// e^3 ~ 1e-18: 1 ulp (measured 2.2e-16) for MUFU + 3 FMA.  IEEE division costs ~10 dependent
// operations and a slow-path branch that stops ptxas from interleaving independent chains.
W5_HD double fast_rcp(double x)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    return fma(r, fma(e, e, e), r);
#else
    return 1.0 / x;
#endif
}

// max(0, x) on the integer pipe (sign bit -> all-zero mask): the FP64 form costs a DSETP on the FP64 pipe, which is
// the binding resource of the flux kernel, plus two selects.  NaN stays NaN (Julia's max propagates it too), -0 -> +0.
W5_HD double relu64(double x)
{
#if defined(__CUDA_ARCH__)
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int m = ~(hi >> 31);
    return __hiloint2double(hi & m, lo & m);
#else
    return x > 0.0 ? x : (x != x ? x : 0.0);
#endif
}

// 1/sqrt(x) for normal x > 0, branch-free: hardware seed + two Newton steps (e -> 1.5 e^2;
// measured 2.7e-16).
W5_HD double fast_rsqrt(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-(hx * y), y, 0.5);
    y = fma(y, e, y);
    e = fma(-(hx * y), y, 0.5);
    y = fma(y, e, y);
    return y;
#else
    return 1.0 / sqrt(x);
#endif
}

// sqrt(x) for x >= 0 (0 gives 0), branch-free, from the rsqrt seed.
W5_HD double fast_sqrt(double x)
{
#if defined(__CUDA_ARCH__)
    const double xs = x + 1e-300;                 // = x for every normal x of interest, keeps rsqrt(0) finite
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(xs));
    // coupled iteration on s ~ sqrt(x) and h ~ 1/(2 sqrt(x)) (one step from the seed: 1.5 e^2 ~ 1e-12) and one
    // correction s + (x - s^2) h, which squares that: five dependent operations after the seed.  The products take x
    // itself, not the guarded value: x = 0 gives 0 without a final select (and a NaN stays a NaN).
    const double s0 = x * y, h0 = 0.5 * y;
    const double r = fma(-s0, h0, 0.5);
    const double s1 = fma(s0, r, s0), h1 = fma(h0, r, h0);
    const double d = fma(-s1, s1, x);
    return fma(d, h1, s1);
#else
    return sqrt(x);
#endif
}

// rho^(gam-1), the one transcendental of the path (E2S/S2E, Weno5_2D.jl:1770-1793; compute_global_vmax, :1962;
// compute_primitive_variables_S, :964).  For the reference's gam = 5/3 the exponent is 2/3 to within one ulp
// (5.0/3.0 - 1.0 = 2/3 + 1.1e-16), so the device takes cbrt(rho)^2: within 3 ulp of pow and about a third of its
// instructions -- the kernels that call it were instruction-bound on pow.  Any other gam goes through pow.
W5_HD double pow_gm1(double rho, double gam)
{
#if defined(__CUDA_ARCH__)
    if (gam == 5.0 / 3.0) {
        const double c = cbrt(rho);
        return c * c;
    }
#endif
    return pow(rho, gam - 1.0);
}

// cons -> prim, cell-centre wave speeds and flux in the sweep frame
// (Weno5_2D.jl:1573-1606, :1608-1667, :1533-1555).  q = [rho,M1,M2,M3,B1,B2,B3,E].
W5_HD void cell_quantities(const double q[8], double gam, CellOut &o)
{
    const double rho = q[0];
    const double ri = fast_rcp(rho);
    const double v1 = q[1] * ri, v2 = q[2] * ri, v3 = q[3] * ri;
    const double B1 = q[4], B2 = q[5], B3 = q[6], E = q[7];
    const double vv = v1 * v1 + v2 * v2 + v3 * v3;
    const double BB = B1 * B1 + B2 * B2 + B3 * B3;
    const double P = (gam - 1.0) * (E - 0.5 * (rho * vv + BB));
    const double cs2 = gam * fabs(P * ri);                       // max(0, .) of Weno5_2D.jl:1619 is a no-op on |.|
    const double bnx2 = relu64(B1 * B1 * ri);
    const double bbn2 = BB * ri;
    const double s = bbn2 + cs2;
    const double root = fast_sqrt(relu64(s * s - 4.0 * bnx2 * cs2));
    o.lf = fast_sqrt(0.5 * (s + root));                          // s, root >= 0
    o.ls = fast_sqrt(relu64(0.5 * (s - root)));
    o.la = fast_sqrt(bnx2);
    o.v1 = v1; o.v2 = v2; o.v3 = v3;
    const double pt = P + 0.5 * BB;
    const double vB = v1 * B1 + v2 * B2 + v3 * B3;
    o.F[0] = rho * v1;
    o.F[1] = q[1] * v1 + pt - B1 * B1;
    o.F[2] = q[2] * v1 - B1 * B2;
    o.F[3] = q[3] * v1 - B1 * B3;
    o.F[4] = B2 * v1 - B1 * v2;
    o.F[5] = B3 * v1 - B1 * v3;
    o.F[6] = (E + pt) * v1 - B1 * vB;
}

// Entropy-variable twin (Weno5_2D.jl:953-969 primitives, :927-950 fluxes): q = [rho,M1,M2,M3,B1,B2,B3,S].
// P = S rho^(gam-1); the face needs P and rho^(1-gam) of both cells (:657-667), so both are returned.
struct CellOutS { double P, rg0; };

W5_HD void cell_quantities_S(const double q[8], double gam, CellOut &o, CellOutS &os)
{
    const double rho = q[0];
    const double ri = fast_rcp(rho);
    const double v1 = q[1] * ri, v2 = q[2] * ri, v3 = q[3] * ri;
    const double B1 = q[4], B2 = q[5], B3 = q[6], S = q[7];
    const double BB = B1 * B1 + B2 * B2 + B3 * B3;
    const double rg1 = pow_gm1(rho, gam);
    const double P = S * rg1;
    os.P = P;
    os.rg0 = fast_rcp(rg1);                                        // rho^(1-gam)
    const double cs2 = gam * fabs(P * ri);                       // max(0, .) of Weno5_2D.jl:1619 is a no-op on |.|
    const double bnx2 = relu64(B1 * B1 * ri);
    const double bbn2 = BB * ri;
    const double s = bbn2 + cs2;
    const double root = fast_sqrt(relu64(s * s - 4.0 * bnx2 * cs2));
    o.lf = fast_sqrt(0.5 * (s + root));                          // s, root >= 0
    o.ls = fast_sqrt(relu64(0.5 * (s - root)));
    o.la = fast_sqrt(bnx2);
    o.v1 = v1; o.v2 = v2; o.v3 = v3;
    const double pt = P + 0.5 * BB;
    o.F[0] = rho * v1;
    o.F[1] = q[1] * v1 + pt - B1 * B1;
    o.F[2] = q[2] * v1 - B1 * B2;
    o.F[3] = q[3] * v1 - B1 * B3;
    o.F[4] = B2 * v1 - B1 * v2;
    o.F[5] = B3 * v1 - B1 * v3;
    o.F[6] = S * v1;
}

// Face-averaged state and everything the structured projection needs.
// Field order m: fast-, Alfven-, slow-, entropy, slow+, Alfven+, fast+.
struct FaceCoef {
    // shared inner products
    double hv2, vx, vy, vz, By, Bz;       // n = -hv2 x1 + vx x2 + vy x3 + vz x4 + By x5 + Bz x6 - x7
    double bty, btz;                      // t = bty x5 + btz x6, u = bty x3 + btz x4
    // fast pair (rows 1,7), slow pair (rows 3,5): w = k_n n + k_t t +- (k_1 x1 + k_2 x2 + k_u u)
    double kf_n, kf_t, kf_1, kf_2, kf_u;
    double ks_n, ks_t, ks_1, ks_2, ks_u;
    // Alfven pair (rows 2,6): w = (ka_1 x1 + ka_3 x3 + ka_4 x4) +- (kb_5 x5 + kb_6 x6)
    double ka_1, ka_3, ka_4, kb_5, kb_6;
    double ke_n;                          // entropy row: w = ke_n n + x1
    // back-projection
    double af, as_, afcf, ascs, a_r, s_r, sgn, sig, hf, hs, axy;
    // S branch only: n = gS x1 + kap x7, entropy row w = x1/gam - kap x7, 7th back-projected component
    double gS, kap, ig, Sr, g0;
};

// Weno5_2D.jl:1273-1359 (face averages, wave speeds, degeneracy handling) and the
// coefficient groups of L (:1368-1426, scaling :1490-1502) and R (:1428-1486).
// The sgnBt flip (:1506-1528) multiplies row m of L and column m of R by the same +-1 and
// WENO is odd in its inputs, so it cancels exactly and is not applied.
// g2 = (gam-2)/(gam-1) is passed in: it is a constant of the run, and computed here it costs an IEEE division
// (with its slow-path call) once per march step in the kernel's loop.
W5_HD void face_coefficients(double rhoL, double rhoR, double v1L, double v1R, double v2L, double v2R,
                             double v3L, double v3R, double B1L, double B1R, double B2L, double B2R,
                             double B3L, double B3R, double EL, double ER, double gam, double g2, FaceCoef &c)
{
    const double g0 = 1.0 - gam;
    const double rho = 0.5 * (rhoL + rhoR);
    const double vx = 0.5 * (v1L + v1R), vy = 0.5 * (v2L + v2R), vz = 0.5 * (v3L + v3R);
    const double Bx = 0.5 * (B1L + B1R), By = 0.5 * (B2L + B2R), Bz = 0.5 * (B3L + B3R);
    const double E = 0.5 * (EL + ER);
    const double rinv = fast_rsqrt(rho);                          // 1/sqrt(rho)
    const double r = rho * rinv;
    const double ri = rinv * rinv;
    const double Bt2 = By * By + Bz * Bz;
    const double B2 = Bx * Bx + Bt2;
    const double v2 = vx * vx + vy * vy + vz * vz;
    const double pg = (gam - 1.0) * (E - 0.5 * (rho * v2 + B2));
    const double bbn2 = B2 * ri;
    const double cs2 = gam * fabs(pg * ri);                      // max(0, .) of Weno5_2D.jl:1310 is a no-op on |.|
    const double bnx2 = Bx * Bx * ri;
    const double a = fast_sqrt(cs2);
    const double s_ = bbn2 + cs2;
    const double root = fast_sqrt(relu64(s_ * s_ - 4.0 * bnx2 * cs2));
    const double cf = fast_sqrt(0.5 * (s_ + root));                // s_, root >= 0
    const double cs = fast_sqrt(relu64(0.5 * (s_ - root)));
    const double sg = (Bx < 0.0) ? -1.0 : 1.0;                    // sign(Bx), 0 -> 1
    // degenerate cases are resolved with selects, not branches (Weno5_2D.jl:1337-1359)
    const bool tang = Bt2 > 1e-30;
    const double ib = fast_rsqrt(tang ? Bt2 : 1.0);
    const double bty = tang ? By * ib : 0.70710678118654752440;
    const double btz = tang ? Bz * ib : 0.70710678118654752440;
    const double dl2 = cf * cf - cs * cs;
    const bool split = dl2 >= 1e-30;
    const double id = fast_rcp(split ? dl2 : 1.0);
    const double af = split ? fast_sqrt(relu64(cs2 - cs * cs) * id) : 1.0;
    const double as_ = split ? fast_sqrt(relu64(cf * cf - cs2) * id) : 1.0;
    const double sig = bty * vy + btz * vz;
    const double ih = 0.5 * fast_rcp(cs2);                         // 1/(2 a^2)
    const double afcf = af * cf, ascs = as_ * cs;

    c.hv2 = 0.5 * v2; c.vx = vx; c.vy = vy; c.vz = vz; c.By = By; c.Bz = Bz;
    c.bty = bty; c.btz = btz;
    // shared factors first: 28 multiplications for the 16 coefficients instead of 41
    const double Fh = afcf * ih, Sh = ascs * ih;                   // af cf / 2a^2, as cs / 2a^2
    const double sgsig = sg * sig, g0ih = g0 * ih, arh = a * r * ih, hsr = 0.5 * sg * r;
    c.kf_n = g0ih * af;
    c.kf_t = arh * as_;
    c.kf_1 = fma(Fh, vx, -Sh * sgsig);
    c.kf_2 = -Fh;
    c.kf_u = Sh * sg;
    c.ks_n = g0ih * as_;
    c.ks_t = -arh * af;
    c.ks_1 = fma(Sh, vx, Fh * sgsig);
    c.ks_2 = -Sh;
    c.ks_u = -Fh * sg;
    c.axy = bty * vz - btz * vy;
    c.ka_1 = -0.5 * c.axy;
    c.ka_3 = -0.5 * btz;
    c.ka_4 = 0.5 * bty;
    c.kb_5 = -hsr * btz;
    c.kb_6 = hsr * bty;
    c.ke_n = -2.0 * g0ih;
    c.af = af; c.as_ = as_; c.afcf = afcf; c.ascs = ascs;
    c.a_r = a * rinv; c.s_r = sg * rinv; c.sgn = sg; c.sig = sig;
    const double h = 0.5 * v2 - g2 * cs2;
    c.hf = af * (cf * cf + h);
    c.hs = as_ * (cs * cs + h);
}

// characteristic projection of one 7-vector x = [rho,M1,M2,M3,B2,B3,E]-shaped quantity
W5_HD void project(const FaceCoef &c, const double x[7], double w[7])
{
    const double n = fma(c.vx, x[1], fma(c.vy, x[2], fma(c.vz, x[3], fma(c.By, x[4], fma(c.Bz, x[5], fma(-c.hv2, x[0], -x[6]))))));
    const double t = fma(c.bty, x[4], c.btz * x[5]);
    const double u = fma(c.bty, x[2], c.btz * x[3]);
    const double pf = fma(c.kf_n, n, c.kf_t * t);
    const double mf = fma(c.kf_1, x[0], fma(c.kf_2, x[1], c.kf_u * u));
    const double ps = fma(c.ks_n, n, c.ks_t * t);
    const double ms = fma(c.ks_1, x[0], fma(c.ks_2, x[1], c.ks_u * u));
    const double A = fma(c.ka_1, x[0], fma(c.ka_3, x[2], c.ka_4 * x[3]));
    const double B = fma(c.kb_5, x[4], c.kb_6 * x[5]);
    w[0] = pf + mf; w[6] = pf - mf;
    w[2] = ps + ms; w[4] = ps - ms;
    w[1] = A + B;   w[5] = A - B;
    w[3] = fma(c.ke_n, n, x[0]);
}

// Two vectors through the same projection, statement by statement in lock-step.  On B200 a DFMA whose three
// sources are three different vector registers takes ~3 cycles of the FP64 pipe's 2-cycle slot -- the register
// file delivers two 64-bit operands per slot (scripts/microbench/fp64_operands.cu: 0.325 instead of 0.5 warp
// instructions per cycle) -- unless one source comes from the operand-reuse cache, i.e. the previous instruction read
// the same register in the same position.  Projecting dF and dq with the same coefficients back to back gives
// every second multiply-add that reuse.  Same operations and rounding as two project() calls.
W5_HD void project_pair(const FaceCoef &c, const double x[7], const double y[7], double w[7], double v[7])
{
    double nx = fma(-c.hv2, x[0], -x[6]), ny = fma(-c.hv2, y[0], -y[6]);
    nx = fma(c.Bz, x[5], nx); ny = fma(c.Bz, y[5], ny);
    nx = fma(c.By, x[4], nx); ny = fma(c.By, y[4], ny);
    nx = fma(c.vz, x[3], nx); ny = fma(c.vz, y[3], ny);
    nx = fma(c.vy, x[2], nx); ny = fma(c.vy, y[2], ny);
    nx = fma(c.vx, x[1], nx); ny = fma(c.vx, y[1], ny);
    double tx = c.btz * x[5], ty = c.btz * y[5];
    tx = fma(c.bty, x[4], tx); ty = fma(c.bty, y[4], ty);
    double ux = c.btz * x[3], uy = c.btz * y[3];
    ux = fma(c.bty, x[2], ux); uy = fma(c.bty, y[2], uy);
    double pfx = c.kf_t * tx, pfy = c.kf_t * ty;
    pfx = fma(c.kf_n, nx, pfx); pfy = fma(c.kf_n, ny, pfy);
    double mfx = c.kf_u * ux, mfy = c.kf_u * uy;
    mfx = fma(c.kf_2, x[1], mfx); mfy = fma(c.kf_2, y[1], mfy);
    mfx = fma(c.kf_1, x[0], mfx); mfy = fma(c.kf_1, y[0], mfy);
    double psx = c.ks_t * tx, psy = c.ks_t * ty;
    psx = fma(c.ks_n, nx, psx); psy = fma(c.ks_n, ny, psy);
    double msx = c.ks_u * ux, msy = c.ks_u * uy;
    msx = fma(c.ks_2, x[1], msx); msy = fma(c.ks_2, y[1], msy);
    msx = fma(c.ks_1, x[0], msx); msy = fma(c.ks_1, y[0], msy);
    double Ax = c.ka_4 * x[3], Ay = c.ka_4 * y[3];
    Ax = fma(c.ka_3, x[2], Ax); Ay = fma(c.ka_3, y[2], Ay);
    Ax = fma(c.ka_1, x[0], Ax); Ay = fma(c.ka_1, y[0], Ay);
    double Bx = c.kb_6 * x[5], By = c.kb_6 * y[5];
    Bx = fma(c.kb_5, x[4], Bx); By = fma(c.kb_5, y[4], By);
    w[0] = pfx + mfx; v[0] = pfy + mfy;
    w[6] = pfx - mfx; v[6] = pfy - mfy;
    w[2] = psx + msx; v[2] = psy + msy;
    w[4] = psx - msx; v[4] = psy - msy;
    w[1] = Ax + Bx;   v[1] = Ay + By;
    w[5] = Ax - Bx;   v[5] = Ay - By;
    w[3] = fma(c.ke_n, nx, x[0]); v[3] = fma(c.ke_n, ny, y[0]);
}

// Back-projection coefficients only (the part of FaceCoef that is cold during the stencil
// streaming); kept separate so that a kernel can park them in shared memory.
struct BackCoef {
    double vx, vy, vz, bty, btz, hv2;
    double af, as_, afcf, ascs, a_r, s_r, sgn, sig, hf, hs, axy;
};

W5_HD void split_back(const FaceCoef &c, BackCoef &b)
{
    b.vx = c.vx; b.vy = c.vy; b.vz = c.vz; b.bty = c.bty; b.btz = c.btz; b.hv2 = c.hv2;
    b.af = c.af; b.as_ = c.as_; b.afcf = c.afcf; b.ascs = c.ascs; b.a_r = c.a_r; b.s_r = c.s_r;
    b.sgn = c.sgn; b.sig = c.sig; b.hf = c.hf; b.hs = c.hs; b.axy = c.axy;
}

// fhat = R . Fs without forming R (Weno5_2D.jl:1879-1888 with :1428-1486)
W5_HD void back_project(const BackCoef &c, const double Fs[7], double fh[7])
{
    const double Sf = Fs[0] + Fs[6], Df = Fs[6] - Fs[0];
    const double Ss = Fs[2] + Fs[4], Ds = Fs[4] - Fs[2];
    const double Sa = Fs[1] + Fs[5], Da = Fs[5] - Fs[1];
    const double X = fma(c.afcf, Ds, -c.ascs * Df);
    const double Y = fma(c.afcf, Df, c.ascs * Ds);
    const double f1 = fma(c.af, Sf, fma(c.as_, Ss, Fs[3]));
    const double sX = c.sgn * X;
    const double Z = c.a_r * fma(c.as_, Sf, -c.af * Ss);
    const double sDa = c.s_r * Da;
    fh[0] = f1;
    fh[1] = fma(c.vx, f1, Y);
    fh[2] = fma(c.vy, f1, fma(c.bty, sX, -c.btz * Sa));
    fh[3] = fma(c.vz, f1, fma(c.btz, sX, c.bty * Sa));
    fh[4] = fma(c.bty, Z, c.btz * sDa);
    fh[5] = fma(c.btz, Z, -c.bty * sDa);
    fh[6] = fma(c.hf, Sf, fma(c.hs, Ss, fma(c.vx, Y, fma(c.sig, sX, fma(c.hv2, Fs[3], c.axy * Sa)))));
}

// ------------------------------------------------------------------ S branch (entropy variable)
// compute_eigenvectors_S_vec, Weno5_2D.jl:640-924.  The rows of L have the same structure as in the E
// branch with the inner product n replaced by n_S = gamS x1 + kappa x7, kappa = rho/(gam S) (after the
// 0.5/cs2 scaling the factor cs2 of columns 1 and 7 cancels), identical antisymmetric parts and
// Alfven rows, and the unscaled entropy row x1/gam - kappa x7 (:872).  R differs in its 7th row only.
W5_HD void face_coefficients_S(double rhoL, double rhoR, double rg0L, double rg0R, double v1L, double v1R,
                               double v2L, double v2R, double v3L, double v3R, double B1L, double B1R, double B2L,
                               double B2R, double B3L, double B3R, double PL, double PR, double gam, FaceCoef &c)
{
    const double g0 = 1.0 - gam;
    const double drho = fabs(rhoL - rhoR);
    // :657-667: arithmetic mean for (nearly) equal densities, else the mean of rho^(1-gam)
    double rho = 0.5 * (rhoL + rhoR);
    if (drho > 1e-12) rho = pow(fabs(g0 * drho / (rg0R - rg0L)), 1.0 / gam);
    const double vx = 0.5 * (v1L + v1R), vy = 0.5 * (v2L + v2R), vz = 0.5 * (v3L + v3R);
    const double Bx = 0.5 * (B1L + B1R), By = 0.5 * (B2L + B2R), Bz = 0.5 * (B3L + B3R);
    const double pg = 0.5 * (PL + PR);
    const double rinv = fast_rsqrt(rho);
    const double r = rho * rinv;
    const double ri = rinv * rinv;
    const double S = pg / pow_gm1(rho, gam);
    const double Bt2 = By * By + Bz * Bz;
    const double B2 = Bx * Bx + Bt2;
    const double bbn2 = B2 * ri;
    const double cs2 = gam * fabs(pg * ri);                      // max(0, .) of Weno5_2D.jl:1310 is a no-op on |.|
    const double bnx2 = Bx * Bx * ri;
    const double a = fast_sqrt(cs2);
    const double s_ = bbn2 + cs2;
    const double root = fast_sqrt(relu64(s_ * s_ - 4.0 * bnx2 * cs2));
    const double cf = fast_sqrt(0.5 * (s_ + root));                // s_, root >= 0
    const double cs = fast_sqrt(relu64(0.5 * (s_ - root)));
    const double sg = (Bx < 0.0) ? -1.0 : 1.0;
    const bool tang = Bt2 > 1e-30;
    const double ib = fast_rsqrt(tang ? Bt2 : 1.0);
    const double bty = tang ? By * ib : 0.70710678118654752440;
    const double btz = tang ? Bz * ib : 0.70710678118654752440;
    const double dl2 = cf * cf - cs * cs;
    const bool split = dl2 >= 1e-30;
    const double id = fast_rcp(split ? dl2 : 1.0);
    const double af = split ? fast_sqrt(relu64(cs2 - cs * cs) * id) : 1.0;
    const double as_ = split ? fast_sqrt(relu64(cf * cf - cs2) * id) : 1.0;
    const double sig = bty * vy + btz * vz;
    const double ih = 0.5 * fast_rcp(cs2);
    const double afcf = af * cf, ascs = as_ * cs;

    c.hv2 = 0.0; c.vx = vx; c.vy = vy; c.vz = vz; c.By = By; c.Bz = Bz;
    c.bty = bty; c.btz = btz;
    c.kf_n = 0.5 * af;
    c.kf_t = a * as_ * r * ih;
    c.kf_1 = (afcf * vx - ascs * sig * sg) * ih;
    c.kf_2 = -afcf * ih;
    c.kf_u = ascs * sg * ih;
    c.ks_n = 0.5 * as_;
    c.ks_t = -a * af * r * ih;
    c.ks_1 = (ascs * vx + afcf * sig * sg) * ih;
    c.ks_2 = -ascs * ih;
    c.ks_u = -afcf * sg * ih;
    c.ka_1 = 0.5 * (btz * vy - bty * vz);
    c.ka_3 = -0.5 * btz;
    c.ka_4 = 0.5 * bty;
    c.kb_5 = -0.5 * sg * r * btz;
    c.kb_6 = 0.5 * sg * r * bty;
    c.ke_n = 0.0;
    c.af = af; c.as_ = as_; c.afcf = afcf; c.ascs = ascs;
    c.a_r = a * rinv; c.s_r = sg * rinv; c.sgn = sg; c.sig = sig;
    c.hf = 0.0; c.hs = 0.0; c.axy = 0.0;
    c.gS = (gam - 1.0) / gam;
    c.kap = rho / (gam * S);
    c.ig = 1.0 / gam;
    c.Sr = S * ri;
    c.g0 = g0;
}

// characteristic projection, S branch: x = [rho,M1,M2,M3,B2,B3,S]-shaped quantity
W5_HD void project_S(const FaceCoef &c, const double x[7], double w[7])
{
    const double n = fma(c.gS, x[0], c.kap * x[6]);
    const double t = fma(c.bty, x[4], c.btz * x[5]);
    const double u = fma(c.bty, x[2], c.btz * x[3]);
    const double pf = fma(c.kf_n, n, c.kf_t * t);
    const double mf = fma(c.kf_1, x[0], fma(c.kf_2, x[1], c.kf_u * u));
    const double ps = fma(c.ks_n, n, c.ks_t * t);
    const double ms = fma(c.ks_1, x[0], fma(c.ks_2, x[1], c.ks_u * u));
    const double A = fma(c.ka_1, x[0], fma(c.ka_3, x[2], c.ka_4 * x[3]));
    const double B = fma(c.kb_5, x[4], c.kb_6 * x[5]);
    w[0] = pf + mf; w[6] = pf - mf;
    w[2] = ps + ms; w[4] = ps - ms;
    w[1] = A + B;   w[5] = A - B;
    w[3] = fma(c.ig, x[0], -c.kap * x[6]);
}

// fhat = R_S . Fs (Weno5_2D.jl:1879-1888 with :812-866): rows 1-6 as in the E branch, row 7 = S/rho (...)
struct BackCoefS {
    double vx, vy, vz, bty, btz;
    double af, as_, afcf, ascs, a_r, s_r, sgn, Sr, g0;
};

W5_HD void split_back_S(const FaceCoef &c, BackCoefS &b)
{
    b.vx = c.vx; b.vy = c.vy; b.vz = c.vz; b.bty = c.bty; b.btz = c.btz;
    b.af = c.af; b.as_ = c.as_; b.afcf = c.afcf; b.ascs = c.ascs; b.a_r = c.a_r; b.s_r = c.s_r;
    b.sgn = c.sgn; b.Sr = c.Sr; b.g0 = c.g0;
}

W5_HD void back_project_S(const BackCoefS &c, const double Fs[7], double fh[7])
{
    const double Sf = Fs[0] + Fs[6], Df = Fs[6] - Fs[0];
    const double Ss = Fs[2] + Fs[4], Ds = Fs[4] - Fs[2];
    const double Sa = Fs[1] + Fs[5], Da = Fs[5] - Fs[1];
    const double X = fma(c.afcf, Ds, -c.ascs * Df);
    const double Y = fma(c.afcf, Df, c.ascs * Ds);
    const double fs = fma(c.af, Sf, c.as_ * Ss);
    const double f1 = fs + Fs[3];
    const double sX = c.sgn * X;
    const double Z = c.a_r * fma(c.as_, Sf, -c.af * Ss);
    const double sDa = c.s_r * Da;
    fh[0] = f1;
    fh[1] = fma(c.vx, f1, Y);
    fh[2] = fma(c.vy, f1, fma(c.bty, sX, -c.btz * Sa));
    fh[3] = fma(c.vz, f1, fma(c.btz, sX, c.bty * Sa));
    fh[4] = fma(c.bty, Z, c.btz * sDa);
    fh[5] = fma(c.btz, Z, -c.bty * sDa);
    fh[6] = c.Sr * fma(c.g0, Fs[3], fs);
}

// One side of the Jiang & Wu correction (Weno5_2D.jl:1836-1853) for UNHALVED inputs
// A = 2a, B = 2b, ...; returns 6 phi (the factor is undone where the two sides are combined).
//  * smoothness indicators as quadratic forms, scaled by 1/16:
//      IS0 = 13(a-b)^2 + 3(a-3b)^2 = 16 a^2 - 44 ab + 40 b^2   ->  A(A - 2.75 B) + 2.5 B^2
//      IS1 = 13(b-c)^2 + 3(b+c)^2  = 16 b^2 - 20 bc + 16 c^2   ->  B(B - 1.25 C) + C^2
//      IS2 = 13(c-d)^2 + 3(3c-d)^2 = 40 c^2 - 44 cd + 16 d^2   ->  D(D - 2.75 C) + 2.5 C^2
//    (10 operations instead of 18; with eps folded in, 11 for the three eps + IS_k instead of 13).  With unhalved inputs these are IS_k/4, so eps is passed as
//    eps/4 and the weight ratios are unchanged.
//  * omega0 = al0/(al0+al1+al2), al = (1,6,3)/d_k, d_k = (eps+IS_k)^2, written over the common
//    denominator so that a single reciprocal replaces five divisions.
// Algebraically identical to the reference; differs in round-off only.
W5_HD double weno_phi6(double A, double B, double C, double D, double epsq)
{
    const double BB = B * B, CC = C * C;
    // eps + IS_k with eps folded into the constant term of each quadratic form (one operation less per indicator)
    double d0 = fma(A, fma(-2.75, B, A), fma(2.5, BB, epsq));
    double d1 = fma(B, fma(-1.25, C, B), CC + epsq);
    double d2 = fma(D, fma(-2.75, C, D), fma(2.5, CC, epsq));
    d0 *= d0; d1 *= d1; d2 *= d2;
    const double p12 = d1 * d2, p02 = d0 * d2, p01 = d0 * d1;
    const double inv = fast_rcp(fma(6.0, p02, fma(3.0, p01, p12)));
    const double D1 = fma(-2.0, B, A) + C;
    const double D2 = fma(-2.0, C, B) + D;
    // 6 phi = omega0 D1 + (omega2 - 1/2) D2 / 2 with omega0 = p12 inv, omega2 = 3 p01 inv, over the common reciprocal
    return fma(fma(1.5, p01 * D2, p12 * D1), inv, -0.25 * D2);
}

// No-op stash: everything stays in registers (host tests, and the reference point for the
// kernel's shared-memory stash).
struct RegStash {
    double q[4][7], first[7];
    W5_HD void put_q(int k, int m, double v) { q[k][m] = v; }
    W5_HD double get_q(int k, int m) const { return q[k][m]; }
    W5_HD void put_first(int m, double v) { first[m] = v; }
    W5_HD double get_first(int m) const { return first[m]; }
};

// Numerical flux of one face.  load(k, F, q) fills the flux and conserved 7-vectors of
// stencil cell k = 0..5 (cells f-2 .. f+3); amax[m] = max(|a_m(f)|, |a_m(f+1)|)
// (Weno5_2D.jl:1834).  Streams over the stencil so that only the Lax-Friedrichs split
// differences are live, not the 84 stencil operands.  The minus-side differences Q and the
// central term `first` are only needed after the streaming loop; they go through `stash`
// (registers, or thread-private shared memory in the kernel) to keep register pressure down.
template <int BR> W5_HD void project_br(const FaceCoef &c, const double x[7], double w[7])
{
    if (BR == 0) project(c, x, w); else project_S(c, x, w);
}

template <int BR, class Loader, class Stash>
W5_HD void face_flux_br(const FaceCoef &c, const double amax[7], double eps, Loader load, Stash &stash, double Fs[7]);

template <int BR, class Loader, class Stash>
W5_HD void face_flux_diff(const FaceCoef &c, const double amax[7], double eps, Loader load, Stash &stash, double Fs[7]);

template <class Loader, class Stash>
W5_HD void face_flux(const FaceCoef &c, const double amax[7], double eps, Loader load, Stash &stash, double Fs[7])
{
    // difference-first projection: measured 1.5 % faster per step on B200 (fewer FP64 operations, no value
    // carried between stencil cells); -DW5_FACE_PLAIN selects the 12-projection form
#ifndef W5_FACE_PLAIN
    face_flux_diff<0>(c, amax, eps, load, stash, Fs);
#else
    face_flux_br<0>(c, amax, eps, load, stash, Fs);
#endif
}

template <int BR, class Loader, class Stash>
W5_HD void face_flux_br(const FaceCoef &c, const double amax[7], double eps, Loader load, Stash &stash, double Fs[7])
{
    double P[4][7];                // P[k-1] = dw_k + amax dv_k (k=1..4); stash.q[k-2] = dw_k - amax dv_k (k=2..5)
    double wp[7], vp[7];
    {
        double F[7], q[7];
        load(0, F, q);
        project_br<BR>(c, F, wp);
        project_br<BR>(c, q, vp);
    }
#pragma unroll
    for (int k = 1; k < 6; ++k) {
        double F[7], q[7], w[7], v[7];
        load(k, F, q);
        project_br<BR>(c, F, w);
        project_br<BR>(c, q, v);
#pragma unroll
        for (int m = 0; m < 7; ++m) {
            const double dw = w[m] - wp[m], dv = v[m] - vp[m];
            if (k <= 4) P[k - 1][m] = fma(amax[m], dv, dw);
            if (k >= 2) stash.put_q(k - 2, m, fma(-amax[m], dv, dw));
            // first = (-w1 + 7 w2 + 7 w3 - w4)/12, Weno5_2D.jl:1827, accumulated as
            // 7 (w2 + w3) - (w1 + w4)
            if (k == 2) stash.put_first(m, fma(7.0, w[m], -wp[m]));                 // 7 w2 - w1
            if (k == 4) stash.put_first(m, fma(7.0, wp[m], stash.get_first(m)) - w[m]);   // + 7 w3 - w4
            wp[m] = w[m]; vp[m] = v[m];
        }
    }
    const double epsq = 0.25 * eps;
#pragma unroll
    for (int m = 0; m < 7; ++m) {
        const double second6 = weno_phi6(P[0][m], P[1][m], P[2][m], P[3][m], epsq);
        const double third6 = weno_phi6(stash.get_q(3, m), stash.get_q(2, m), stash.get_q(1, m), stash.get_q(0, m), epsq);
        Fs[m] = fma(stash.get_first(m), 1.0 / 12.0, (1.0 / 6.0) * (third6 - second6));
    }
}

// Variant of face_flux_br that projects DIFFERENCES of neighbouring stencil cells (the projection is linear):
// 10 projections of dF_k, dq_k plus one of the central combination 7(F2+F3) - (F1+F4) instead of 12 projections
// of F_k, q_k.  No value is carried from one stencil cell to the next except the raw F, q (which the loader
// re-reads from shared memory), so the five k-steps are independent instruction streams.
// Algebraically identical; differs in round-off only.  The default of the E branch (see face_flux).
template <int BR, class Loader, class Stash>
W5_HD void face_flux_diff(const FaceCoef &c, const double amax[7], double eps, Loader load, Stash &stash, double Fs[7])
{
    double P[4][7];
    double Fa[7], qa[7], comb[7];
    const double epsq = 0.25 * eps;
    load(0, Fa, qa);
    auto kstep = [&](const int k) {
        double Fb[7], qb[7], dF[7], dq[7], dw[7], dv[7];
        load(k, Fb, qb);
#pragma unroll
        for (int m = 0; m < 7; ++m) {
            dF[m] = Fb[m] - Fa[m];
            dq[m] = qb[m] - qa[m];
            // central term: 7 (F2 + F3) - (F1 + F4), accumulated as the cells stream by (k = 1..4 are cells 1..4)
            if (k == 1) comb[m] = -Fb[m];
            if (k == 2) comb[m] = fma(7.0, Fb[m], comb[m]);
            if (k == 3) comb[m] = fma(7.0, Fb[m], comb[m]);
            if (k == 4) comb[m] = comb[m] - Fb[m];
            Fa[m] = Fb[m]; qa[m] = qb[m];
        }
        if constexpr (BR == 0) {
            project_pair(c, dF, dq, dw, dv);
        } else {
            project_br<BR>(c, dF, dw);
            project_br<BR>(c, dq, dv);
        }
#pragma unroll
        for (int m = 0; m < 7; ++m) {
            if (k <= 4) P[k - 1][m] = fma(amax[m], dv[m], dw[m]);
            if (k >= 2) stash.put_q(k - 2, m, fma(-amax[m], dv[m], dw[m]));
        }
    };
    kstep(1); kstep(2); kstep(3); kstep(4);
    kstep(5);
    double second6[7];
#pragma unroll
    for (int m = 0; m < 7; ++m) second6[m] = weno_phi6(P[0][m], P[1][m], P[2][m], P[3][m], epsq);
    double first[7];
    project_br<BR>(c, comb, first);
#pragma unroll
    for (int m = 0; m < 7; ++m) {
        const double third6 = weno_phi6(stash.get_q(3, m), stash.get_q(2, m), stash.get_q(1, m), stash.get_q(0, m), epsq);
        Fs[m] = fma(first[m], 1.0 / 12.0, (1.0 / 6.0) * (third6 - second6[m]));
    }
}

// max(|a|, |b|): non-negative doubles order like their bit patterns, so the maximum is an integer compare off the
// FP64 pipe (a NaN has the largest pattern and survives, like Julia's max)
W5_HD double absmax(double a, double b)
{
#if defined(__CUDA_ARCH__)
    const long long x = __double_as_longlong(fabs(a)), y = __double_as_longlong(fabs(b));
    return __longlong_as_double(x > y ? x : y);
#else
    return fmax(fabs(a), fabs(b));
#endif
}

W5_HD void face_amax(double v1L, double lfL, double laL, double lsL,
                     double v1R, double lfR, double laR, double lsR, double amax[7])
{
    amax[0] = absmax(v1L - lfL, v1R - lfR);
    amax[1] = absmax(v1L - laL, v1R - laR);
    amax[2] = absmax(v1L - lsL, v1R - lsR);
    amax[3] = absmax(v1L, v1R);
    amax[4] = absmax(v1L + lsL, v1R + lsR);
    amax[5] = absmax(v1L + laL, v1R + laR);
    amax[6] = absmax(v1L + lfL, v1R + lfR);
}

// Weno5_2D.jl:1770-1793
W5_HD double e2s(double rho, double m1, double m2, double m3, double b1, double b2, double b3, double E, double gam)
{
    const double mom2 = m1 * m1 + m2 * m2 + m3 * m3;
    const double BB = b1 * b1 + b2 * b2 + b3 * b3;
    return (E - 0.5 * mom2 / rho - 0.5 * BB) * (gam - 1.0) / pow_gm1(rho, gam);
}

W5_HD double s2e(double rho, double m1, double m2, double m3, double b1, double b2, double b3, double S, double gam)
{
    const double mom2 = m1 * m1 + m2 * m2 + m3 * m3;
    const double BB = b1 * b1 + b2 * b2 + b3 * b3;
    return pow_gm1(rho, gam) * S / (gam - 1.0) + 0.5 * BB + 0.5 * mom2 / rho;
}

// E2S followed by S2E on the same cell (the residue of ES_Switch, Weno5_2D.jl:1738 and :1765): the two
// functions above evaluate the same pow(rho, gam-1) and mom2/rho, so one evaluation serves both;
// the results are bit-identical to calling e2s() and then s2e().
W5_HD void e2s_s2e(double rho, double m1, double m2, double m3, double b1, double b2, double b3, double E, double gam,
                   double &S, double &E2)
{
    const double mom2 = m1 * m1 + m2 * m2 + m3 * m3;
    const double BB = b1 * b1 + b2 * b2 + b3 * b3;
    const double pr = pow_gm1(rho, gam);
    const double kin = 0.5 * mom2 / rho;
    S = (E - kin - 0.5 * BB) * (gam - 1.0) / pr;
    E2 = pr * S / (gam - 1.0) + 0.5 * BB + kin;
}

}  // namespace w5
