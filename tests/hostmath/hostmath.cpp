// TEST INFRASTRUCTURE: compiles the product's per-face arithmetic (julia_b200/csrc/weno5_math.cuh)
// for the HOST so the structured eigen-projection / WENO / back-projection can be compared with
// the CPU oracle in a container without a GPU.  Not part of the product library.
#include "../../julia_b200/csrc/weno5_math.cuh"
#include <vector>
#include <cstring>

// One directional sweep over a public-layout grid (Nx,Ny,8 planes: rho,Mx,My,Mz,Bx,By,Bz,E).
// dir 0: along x; dir 1: along y with the (x,y,z)->(y,z,x) permutation.  Faces 2..N-3 (0-based)
// of every line; the stencil cell past the end is clamped (bc 0) or wrapped (bc 1).
// Outputs: dq (8 planes, physical components) and bs (fsy for dir 0, gsx for dir 1).
// BR 0: E branch (8th plane = E), BR 1: S branch (8th plane = S; two more per-cell entries: P, rho^(1-gam)).
template <int BR>
static int sweep_t(int Nx, int Ny, const double *q, int dir, int bc, double gam, double eps,
                   double *dq, double *bs)
{
    constexpr int NCQ = w5::CQ_COUNT + 2, CQ_P = w5::CQ_COUNT, CQ_RG0 = w5::CQ_COUNT + 1;
    const size_t pl = (size_t)Nx * Ny;
    std::memset(dq, 0, 8 * pl * sizeof(double));
    std::memset(bs, 0, pl * sizeof(double));
    static const int inx[8] = {0, 1, 2, 3, 4, 5, 6, 7}, iny[8] = {0, 2, 3, 1, 5, 6, 4, 7};
    static const int outx[7] = {0, 1, 2, 3, 5, 6, 7}, outy[7] = {0, 2, 3, 1, 6, 4, 7};
    const int *in = dir ? iny : inx, *out = dir ? outy : outx;
    const int N = dir ? Ny : Nx, NP = dir ? Nx : Ny;
    const int ni = N - 6;
    std::vector<double> cq((size_t)N * NCQ), fh((size_t)N * 7);
    for (int p = 0; p < NP; ++p) {
        auto at = [&](int s, int c) { return dir ? q[(size_t)in[c] * pl + p + (size_t)Nx * s] : q[(size_t)in[c] * pl + s + (size_t)Nx * p]; };
        for (int s = 0; s < N; ++s) {
            double qq[8];
            for (int c = 0; c < 8; ++c) qq[c] = at(s, c);
            w5::CellOut o;
            w5::CellOutS os = {0.0, 0.0};
            if (BR == 0) w5::cell_quantities(qq, gam, o); else w5::cell_quantities_S(qq, gam, o, os);
            double *cc = &cq[(size_t)s * NCQ];
            cc[CQ_P] = os.P; cc[CQ_RG0] = os.rg0;
            for (int c = 0; c < 7; ++c) cc[w5::CQ_F0 + c] = o.F[c];
            cc[w5::CQ_Q0 + 0] = qq[0]; cc[w5::CQ_Q0 + 1] = qq[1]; cc[w5::CQ_Q0 + 2] = qq[2]; cc[w5::CQ_Q0 + 3] = qq[3];
            cc[w5::CQ_Q0 + 4] = qq[5]; cc[w5::CQ_Q0 + 5] = qq[6]; cc[w5::CQ_Q0 + 6] = qq[7];
            cc[w5::CQ_V1] = o.v1; cc[w5::CQ_V2] = o.v2; cc[w5::CQ_V3] = o.v3; cc[w5::CQ_B1] = qq[4];
            cc[w5::CQ_LF] = o.lf; cc[w5::CQ_LA] = o.la; cc[w5::CQ_LS] = o.ls;
        }
        std::fill(fh.begin(), fh.end(), 0.0);
        for (int f = 2; f <= N - 3; ++f) {
            const double *L = &cq[(size_t)f * NCQ], *R = &cq[(size_t)(f + 1) * NCQ];
            w5::FaceCoef c;
            if (BR == 0)
                w5::face_coefficients(L[w5::CQ_Q0], R[w5::CQ_Q0], L[w5::CQ_V1], R[w5::CQ_V1], L[w5::CQ_V2], R[w5::CQ_V2],
                                      L[w5::CQ_V3], R[w5::CQ_V3], L[w5::CQ_B1], R[w5::CQ_B1], L[w5::CQ_Q0 + 4], R[w5::CQ_Q0 + 4],
                                      L[w5::CQ_Q0 + 5], R[w5::CQ_Q0 + 5], L[w5::CQ_Q0 + 6], R[w5::CQ_Q0 + 6], gam, (gam - 2.0) / (gam - 1.0), c);
            else
                w5::face_coefficients_S(L[w5::CQ_Q0], R[w5::CQ_Q0], L[CQ_RG0], R[CQ_RG0], L[w5::CQ_V1], R[w5::CQ_V1],
                                        L[w5::CQ_V2], R[w5::CQ_V2], L[w5::CQ_V3], R[w5::CQ_V3], L[w5::CQ_B1], R[w5::CQ_B1],
                                        L[w5::CQ_Q0 + 4], R[w5::CQ_Q0 + 4], L[w5::CQ_Q0 + 5], R[w5::CQ_Q0 + 5], L[CQ_P], R[CQ_P],
                                        gam, c);
            double amax[7];
            w5::face_amax(L[w5::CQ_V1], L[w5::CQ_LF], L[w5::CQ_LA], L[w5::CQ_LS],
                          R[w5::CQ_V1], R[w5::CQ_LF], R[w5::CQ_LA], R[w5::CQ_LS], amax);
            auto load = [&](int k, double F[7], double x[7]) {
                int s = f - 2 + k;
                if (s >= N) s = bc ? s - ni : N - 1;
                const double *cc = &cq[(size_t)s * NCQ];
                for (int m = 0; m < 7; ++m) { F[m] = cc[w5::CQ_F0 + m]; x[m] = cc[w5::CQ_Q0 + m]; }
            };
            w5::RegStash stash;
            double Fs[7];
            // the same forms the kernel instantiates: E branch difference-first, S branch plain
            if (BR == 0) w5::face_flux(c, amax, eps, load, stash, Fs);
            else w5::face_flux_br<BR>(c, amax, eps, load, stash, Fs);
            if (BR == 0) {
                w5::BackCoef bc_;
                w5::split_back(c, bc_);
                w5::back_project(bc_, Fs, &fh[(size_t)f * 7]);
            } else {
                w5::BackCoefS bc_;
                w5::split_back_S(c, bc_);
                w5::back_project_S(bc_, Fs, &fh[(size_t)f * 7]);
            }
            const double b = dir ? fh[(size_t)f * 7 + 5] + 0.5 * (L[w5::CQ_B1] * L[w5::CQ_V3] + R[w5::CQ_B1] * R[w5::CQ_V3])
                                 : fh[(size_t)f * 7 + 4] + 0.5 * (L[w5::CQ_B1] * L[w5::CQ_V2] + R[w5::CQ_B1] * R[w5::CQ_V2]);
            const size_t g = dir ? (size_t)p + (size_t)Nx * f : (size_t)f + (size_t)Nx * p;
            bs[g] = b;
            for (int m = 0; m < 7; ++m) dq[(size_t)out[m] * pl + g] = fh[(size_t)f * 7 + m] - fh[(size_t)(f - 1) * 7 + m];
        }
    }
    return 0;
}

extern "C" {

int hm_sweep(int Nx, int Ny, const double *q, int dir, int bc, double gam, double eps, double *dq, double *bs)
{ return sweep_t<0>(Nx, Ny, q, dir, bc, gam, eps, dq, bs); }
// the entropy-variable sweep (weno5_flux_difference_S, Weno5_2D.jl:361): planes rho,Mx,My,Mz,Bx,By,Bz,S
int hm_sweep_S(int Nx, int Ny, const double *q, int dir, int bc, double gam, double eps, double *dq, double *bs)
{ return sweep_t<1>(Nx, Ny, q, dir, bc, gam, eps, dq, bs); }

double hm_e2s(double rho, double m1, double m2, double m3, double b1, double b2, double b3, double E, double gam)
{ return w5::e2s(rho, m1, m2, m3, b1, b2, b3, E, gam); }
double hm_s2e(double rho, double m1, double m2, double m3, double b1, double b2, double b3, double S, double gam)
{ return w5::s2e(rho, m1, m2, m3, b1, b2, b3, S, gam); }
void hm_e2s_s2e(double rho, double m1, double m2, double m3, double b1, double b2, double b3, double E, double gam, double *out)
{ w5::e2s_s2e(rho, m1, m2, m3, b1, b2, b3, E, gam, out[0], out[1]); }
}
