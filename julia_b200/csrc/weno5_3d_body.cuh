// 3-D extension of the WENO5 MHD step: per-thread bodies of every kernel and the stage sequence.
//
// The reference has no 3-D solver; its classical_RK4_step docstring (Weno5_2D.jl:113-123) states the design this file
// follows: the scheme is dimensionally unsplit and every direction is swept by the same functions on a rotated state.
// So the step is classical_RK4_step (:125-317) with a third sweep in the frame (z,x,y) -- the cyclic continuation of
// rotate_state (:1897-1927) -- flux differences of rho, M, E from three sweeps, and constrained transport of three face
// fields from edge-centred EMFs built of six transport fluxes (corner_fluxes :319-342 and the face updates :150-151
// continued to 3-D, Ryu et al. 1998):
//   x sweep: f_y = By vx, f_z = Bz vx;  y sweep: g_z = Bz vy, g_x = Bx vy;  z sweep: h_x = Bx vz, h_y = By vz
//   Ox = (h_y + h_y[j+1] - g_z - g_z[k+1])/2,  Oy = (f_z + f_z[k+1] - h_x - h_x[i+1])/2,  Oz = (g_x + g_x[i+1] - f_y - f_y[j+1])/2
//   bxb -= c dt ((Oz - Oz[j-1])/dy - (Oy - Oy[k-1])/dz), byb, bzb cyclically.
// A state that is uniform along one axis reproduces the 2-D step in every orientation (tests/test_gpu_3d.py).
//
// Everything here is __host__ __device__: weno5_3d.cu wraps the bodies in kernels (one thread = one `tid` / one cell),
// tests/hostmath/host3d.cpp runs the very same bodies and the same stage sequence in loops on the CPU, where they are
// compared with the oracle without a GPU.  The product only ever runs them on the device.
//
// Layout: the public one, continued to 3-D: volume index i + Nx (j + Ny k), NBnd = 3 ghost layers, no padding --
// weno5_upload_3d is a plain copy.  Ghost handling is by gather: a ghost location reads the location the boundary
// rule maps it to (per axis: periodic wrap, outflow clamp), which reproduces boundaries! (:1684-1732) applied axis by
// axis, including its rules for flux / face arrays (copy index 3 / N-2 outwards, zero rim).
#pragma once
#include <stddef.h>
#include "weno5_math.cuh"

namespace w5 {
namespace d3 {

constexpr int NB3 = 3;
// CTA shape of the sweep kernels (overridable at compile time for tuning runs): threads, faces along the sweep and
// lines per tile for the x sweep (0) and for the y / z sweeps (1).  Two 128-thread CTAs per SM overlap one CTA's cell
// phase and barriers with the other's face phase.
#ifndef W5_3D_NT
#define W5_3D_NT 128
#endif
#ifndef W5_3D_NF0
#define W5_3D_NF0 123
#endif
#ifndef W5_3D_LW0
#define W5_3D_LW0 1
#endif
#ifndef W5_3D_NF1
#define W5_3D_NF1 32
#endif
#ifndef W5_3D_LW1
#define W5_3D_LW1 8
#endif
#ifndef W5_3D_MINB
#define W5_3D_MINB 2
#endif
constexpr int NT3 = W5_3D_NT;            // threads per CTA of the sweep kernels

struct Dims {
    int Nx, Ny, Nz;
    int bc[3];                           // 0 outflow, 1 periodic
};

W5_HD size_t vidx(const Dims &d, int i, int j, int k)
{
    return (size_t)i + (size_t)d.Nx * ((size_t)j + (size_t)d.Ny * (size_t)k);
}
W5_HD int dimN(const Dims &d, int axis) { return axis == 0 ? d.Nx : (axis == 1 ? d.Ny : d.Nz); }

// boundaries!(q) (Weno5_2D.jl:1686-1702) as a map: ghost cell -> interior cell [3, N-4] (0-based)
W5_HD int map_cell(int idx, int N, int bc)
{
    const int n = N - 2 * NB3;
    if (idx < NB3) return bc == 1 ? idx + n : NB3;
    if (idx > N - NB3 - 1) return bc == 1 ? idx - n : N - NB3 - 1;
    return idx;
}
// the flux / face-array rules (:1704-1729): outflow copies index 2 / N-3 (0-based) outwards, periodic wraps
W5_HD int map_flux(int idx, int N, int bc)
{
    const int n = N - 2 * NB3;
    if (bc == 1) {
        if (idx < NB3) return idx + n;
        if (idx > N - NB3 - 1) return idx - n;
        return idx;
    }
    return idx < 2 ? 2 : (idx > N - 3 ? N - 3 : idx);
}
W5_HD bool on_rim(int idx, int N, int bc) { return bc != 1 && (idx == 0 || idx == N - 1); }

// value of a flux / face array after boundaries! has been applied to it, read from the un-filled array
W5_HD double flux_at(const double *f, const Dims &d, int i, int j, int k)
{
    if (on_rim(i, d.Nx, d.bc[0]) || on_rim(j, d.Ny, d.bc[1]) || on_rim(k, d.Nz, d.bc[2])) return 0.0;
    return f[vidx(d, map_flux(i, d.Nx, d.bc[0]), map_flux(j, d.Ny, d.bc[1]), map_flux(k, d.Nz, d.bc[2]))];
}

// Lines a sweep has to cover along a lateral axis: the transport fluxes are read one line beyond the interior only
// where flux_at clamps (outflow); on a periodic axis it wraps into the interior and the ghost lines are never read.
W5_HD int line_lo(int bc) { return bc == 1 ? NB3 : NB3 - 1; }
W5_HD int line_count(int N, int bc) { return N - 2 * line_lo(bc); }

// ------------------------------------------------------------------ directional sweep
// weno5_flux_difference_E (Weno5_2D.jl:973-1018) along axis DIR of the volume, in the frame (DIR, DIR+1, DIR+2).
struct SweepArgs {
    const double *q[8];      // physical volumes rho,Mx,My,Mz,Bx,By,Bz,E
    double *rhs[5];          // rho,Mx,My,Mz,E: sum over the sweeps of (fhat[c] - fhat[c-1]) / h
    double *bsA, *bsB;       // the two transport fluxes of this direction (table above)
    const double *dtp;       // device time step; 0 = finished, the kernel returns
    Dims d;
    double inv_h, gam, eps, g2;
};

// CTA tile: NF faces along the sweep x LW lines; x is always the fastest thread coordinate (coalesced loads)
template <int DIR> struct SweepGeo {
    static constexpr int NF = DIR == 0 ? W5_3D_NF0 : W5_3D_NF1;
    static constexpr int LW = DIR == 0 ? W5_3D_LW0 : W5_3D_LW1;
    static constexpr int SW = NF + 5;                  // cells along the sweep: faces f0 .. f0+NF-1 need f0-2 .. f0+NF+2
    static constexpr int NCELLS = SW * LW, NFACES = NF * LW;
    static constexpr int CBS = NCELLS, FBS = NFACES;
    static constexpr int NFB = 5;                      // fhat components kept for the flux difference: rho, M1, M2, M3, E
    static constexpr size_t SMEM = (size_t)(CQ_COUNT * CBS + NFB * FBS) * sizeof(double);
    W5_HD static int cell_slot(int s, int l) { return DIR == 0 ? l * SW + s : s * LW + l; }
    W5_HD static void cell_decode(int idx, int &s, int &l)
    {
        if (DIR == 0) { s = idx % SW; l = idx / SW; } else { l = idx % LW; s = idx / LW; }
    }
    W5_HD static int face_slot(int fs, int l) { return DIR == 0 ? l * NF + fs : fs * LW + l; }
    W5_HD static void face_decode(int idx, int &fs, int &l)
    {
        if (DIR == 0) { fs = idx % NF; l = idx / NF; } else { l = idx % LW; fs = idx / LW; }
    }
    // number of tiles along the sweep: faces 2 .. n-3, a tile advances by NF-1 cells
    W5_HD static int tiles_along(int n) { return (n - 5 + NF - 2) / (NF - 1); }
    // grid: DIR 0: (tiles along x, line groups in j, k);  DIR 1: (line groups in i, tiles along y, k);
    //       DIR 2: (line groups in i, tiles along z, j);  lines run over the interior, plus one ghost line each side on
    //       outflow axes (line_lo)
    W5_HD static void grid(const Dims &d, int g[3])
    {
        const int cx = line_count(d.Nx, d.bc[0]), cy = line_count(d.Ny, d.bc[1]), cz = line_count(d.Nz, d.bc[2]);
        if (DIR == 0) { g[0] = tiles_along(d.Nx); g[1] = (cy + LW - 1) / LW; g[2] = cz; }
        if (DIR == 1) { g[0] = (cx + LW - 1) / LW; g[1] = tiles_along(d.Ny); g[2] = cz; }
        if (DIR == 2) { g[0] = (cx + LW - 1) / LW; g[1] = tiles_along(d.Nz); g[2] = cy; }
    }
};

// Position of a CTA: first face f0 along the sweep and the lateral coordinates (p1, p2) of line l
template <int DIR> struct SweepPos {
    int f0, n;            // first face, cells along the sweep
    int l0, p2;           // first lateral coordinate of line 0, second lateral coordinate
    int N1, N2;           // extents of the lateral axes
    int hi1;              // last line along the first lateral axis
    W5_HD SweepPos(const Dims &d, int bx, int by, int bz)
    {
        using T = SweepGeo<DIR>;
        const int a1 = DIR == 0 ? 1 : 0, a2 = DIR == 2 ? 1 : 2;          // lateral axes: (y,z), (x,z), (x,y)
        const int lo1 = line_lo(d.bc[a1]), lo2 = line_lo(d.bc[a2]);
        n = dimN(d, DIR); N1 = dimN(d, a1); N2 = dimN(d, a2);
        hi1 = N1 - 1 - lo1;
        p2 = lo2 + bz;                                                   // <= N2-1-lo2 by the grid size
        if (DIR == 0) { f0 = 2 + bx * (T::NF - 1); l0 = lo1 + by * T::LW; }
        else          { l0 = lo1 + bx * T::LW; f0 = 2 + by * (T::NF - 1); }
    }
    W5_HD bool line_ok(int l) const { return l0 + l <= hi1; }
    W5_HD bool line_interior(int l) const { return l0 + l >= 3 && l0 + l <= N1 - 4 && p2 >= 3 && p2 <= N2 - 4; }
    W5_HD size_t gidx(const Dims &d, int s, int l) const
    {
        const int p1 = l0 + l;
        return DIR == 0 ? vidx(d, s, p1, p2) : (DIR == 1 ? vidx(d, p1, s, p2) : vidx(d, p1, p2, s));
    }
};

// sweep-frame component c reads physical volume sweep_in(DIR, c): (x,y,z) -> (y,z,x) -> (z,x,y)
W5_HD int sweep_in(int dir, int c)
{
    // rho | M1 M2 M3 | B1 B2 B3 | E
    if (c == 0 || c == 7) return c;
    const int base = c < 4 ? 1 : 4, m = c - base;
    return base + (m + dir) % 3;
}
// fhat component kept in slot m = 0..4 (rho, M1, M2, M3, E) -> rhs volume (rho, Mx, My, Mz, E)
W5_HD int sweep_out(int dir, int m)
{
    if (m == 0 || m == 4) return m;
    return 1 + (m - 1 + dir) % 3;
}

// phase 0: cons -> prim, wave speeds, flux of every cell of the tile -> shared memory
template <int DIR>
W5_HD void sweep_cells(const SweepArgs &a, int bx, int by, int bz, int tid, double *smem)
{
    using T = SweepGeo<DIR>;
    const SweepPos<DIR> P(a.d, bx, by, bz);
    double *cb = smem;
    const int ni = P.n - 2 * NB3;
    for (int idx = tid; idx < T::NCELLS; idx += NT3) {
        int s, l;
        T::cell_decode(idx, s, l);
        if (!P.line_ok(l)) continue;
        int cs = P.f0 - 2 + s;
        // the stencil cell past the end of the line continues the boundary condition (quirk Q3 of the 2-D path)
        if (cs >= P.n) cs = a.d.bc[DIR] == 1 ? cs - ni : P.n - 1;
        if (cs >= P.n) cs = P.n - 1;                       // tiles reaching further than any face they own
        const size_t g = P.gidx(a.d, cs, l);
        double q[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) q[c] = a.q[sweep_in(DIR, c)][g];
        CellOut o;
        cell_quantities(q, a.gam, o);
        const int i = T::cell_slot(s, l);
#pragma unroll
        for (int m = 0; m < 7; ++m) cb[(CQ_F0 + m) * T::CBS + i] = o.F[m];
        cb[(CQ_Q0 + 0) * T::CBS + i] = q[0];
        cb[(CQ_Q0 + 1) * T::CBS + i] = q[1];
        cb[(CQ_Q0 + 2) * T::CBS + i] = q[2];
        cb[(CQ_Q0 + 3) * T::CBS + i] = q[3];
        cb[(CQ_Q0 + 4) * T::CBS + i] = q[5];
        cb[(CQ_Q0 + 5) * T::CBS + i] = q[6];
        cb[(CQ_Q0 + 6) * T::CBS + i] = q[7];
        cb[CQ_V1 * T::CBS + i] = o.v1;
        cb[CQ_V2 * T::CBS + i] = o.v2;
        cb[CQ_V3 * T::CBS + i] = o.v3;
        cb[CQ_B1 * T::CBS + i] = q[4];
        cb[CQ_LF * T::CBS + i] = o.lf;
        cb[CQ_LA * T::CBS + i] = o.la;
        cb[CQ_LS * T::CBS + i] = o.ls;
    }
}

// phase 1: numerical flux of every face of the tile (compute_eigenvectors_E_vec + weno5_interpolation,
// Weno5_2D.jl:1250-1531, :1795-1895, through the structured forms of weno5_math.cuh) and the two transport fluxes
template <int DIR>
W5_HD void sweep_faces(const SweepArgs &a, int bx, int by, int bz, int tid, double *smem)
{
    using T = SweepGeo<DIR>;
    const SweepPos<DIR> P(a.d, bx, by, bz);
    const double *cb = smem;
    double *fb = smem + CQ_COUNT * T::CBS;
    for (int idx = tid; idx < T::NFACES; idx += NT3) {
        int fs, l;
        T::face_decode(idx, fs, l);
        const int f = P.f0 + fs;
        if (f > P.n - 3 || !P.line_ok(l)) continue;
        const int iL = T::cell_slot(fs + 2, l), iR = T::cell_slot(fs + 3, l);
        FaceCoef fc;
        face_coefficients(cb[CQ_Q0 * T::CBS + iL], cb[CQ_Q0 * T::CBS + iR], cb[CQ_V1 * T::CBS + iL], cb[CQ_V1 * T::CBS + iR],
                          cb[CQ_V2 * T::CBS + iL], cb[CQ_V2 * T::CBS + iR], cb[CQ_V3 * T::CBS + iL], cb[CQ_V3 * T::CBS + iR],
                          cb[CQ_B1 * T::CBS + iL], cb[CQ_B1 * T::CBS + iR], cb[(CQ_Q0 + 4) * T::CBS + iL],
                          cb[(CQ_Q0 + 4) * T::CBS + iR], cb[(CQ_Q0 + 5) * T::CBS + iL], cb[(CQ_Q0 + 5) * T::CBS + iR],
                          cb[(CQ_Q0 + 6) * T::CBS + iL], cb[(CQ_Q0 + 6) * T::CBS + iR], a.gam, a.g2, fc);
        double amax[7];
        face_amax(cb[CQ_V1 * T::CBS + iL], cb[CQ_LF * T::CBS + iL], cb[CQ_LA * T::CBS + iL], cb[CQ_LS * T::CBS + iL],
                  cb[CQ_V1 * T::CBS + iR], cb[CQ_LF * T::CBS + iR], cb[CQ_LA * T::CBS + iR], cb[CQ_LS * T::CBS + iR], amax);
        auto load = [&](int k, double F[7], double x[7]) {
            const int ii = T::cell_slot(fs + k, l);
#pragma unroll
            for (int m = 0; m < 7; ++m) {
                F[m] = cb[(CQ_F0 + m) * T::CBS + ii];
                x[m] = cb[(CQ_Q0 + m) * T::CBS + ii];
            }
        };
        double fh[7], Fs[7];
        RegStash st;
        BackCoef bc;
        split_back(fc, bc);
        face_flux(fc, amax, a.eps, load, st, Fs);
        back_project(bc, Fs, fh);
        const int io = T::face_slot(fs, l);
        fb[0 * T::FBS + io] = fh[0];
        fb[1 * T::FBS + io] = fh[1];
        fb[2 * T::FBS + io] = fh[2];
        fb[3 * T::FBS + io] = fh[3];
        fb[4 * T::FBS + io] = fh[6];
        // transport fluxes (Weno5_2D.jl:1005-1008): the advective part of the induction flux, upwinded
        const double bv2 = cb[CQ_B1 * T::CBS + iL] * cb[CQ_V2 * T::CBS + iL] + cb[CQ_B1 * T::CBS + iR] * cb[CQ_V2 * T::CBS + iR];
        const double bv3 = cb[CQ_B1 * T::CBS + iL] * cb[CQ_V3 * T::CBS + iL] + cb[CQ_B1 * T::CBS + iR] * cb[CQ_V3 * T::CBS + iR];
        const size_t g = P.gidx(a.d, f, l);
        a.bsA[g] = fma(0.5, bv2, fh[4]);
        a.bsB[g] = fma(0.5, bv3, fh[5]);
    }
}

// phase 2: flux difference of the cells whose two faces the tile holds (Weno5_2D.jl:993-1003)
template <int DIR>
W5_HD void sweep_update(const SweepArgs &a, int bx, int by, int bz, int tid, const double *smem)
{
    using T = SweepGeo<DIR>;
    const SweepPos<DIR> P(a.d, bx, by, bz);
    const double *fb = smem + CQ_COUNT * T::CBS;
    for (int idx = tid; idx < T::NFACES; idx += NT3) {
        int fs, l;
        T::face_decode(idx, fs, l);
        const int c = P.f0 + fs;                       // cell c: faces c-1 (slot fs-1) and c (slot fs)
        if (fs == 0 || c > P.n - NB3 - 1 || !P.line_ok(l) || !P.line_interior(l)) continue;
        const int i0 = T::face_slot(fs - 1, l), i1 = T::face_slot(fs, l);
        const size_t g = P.gidx(a.d, c, l);
#pragma unroll
        for (int m = 0; m < 5; ++m) {
            const double dfl = (fb[m * T::FBS + i1] - fb[m * T::FBS + i0]) * a.inv_h;
            double *r = a.rhs[sweep_out(DIR, m)];
            if (DIR == 0) r[g] = dfl; else r[g] += dfl;
        }
    }
}

// ------------------------------------------------------------------ y / z sweep: one thread marches one line
// In the y and z sweeps the stencil runs across the threads' lines, never along them: a thread that marches along the
// sweep axis needs nobody else's cells.  So every thread owns one line (x index = lane: coalesced), keeps the last six
// cells in a private ring in shared memory (thread-interleaved, conflict-free), computes per march step one cell, one
// face and one flux difference, and carries the previous face's fhat in registers.  No barrier in the kernel, no cell
// is computed twice and no face but the one two segments share, and the eight warps of an SM drift freely, so the cell
// and update work of one warp overlaps the face work of another.
// The raw state of a cell travels global -> shared memory by cp.async one march step ahead (straight into the ring
// slots it will occupy), so no global-load latency sits in the march.
// Per thread: 7 slots x q7[7] (cells f-2 .. f+4, the last one in flight) + 6 x F[7] + 5 x B1 + 4 x (1/rho, lf, la, ls)
// = 112 doubles; 256 threads = 224 KB.
struct MarchGeo {
    static constexpr int NT = 256;                     // 32 lines along x  x  8 values of the other lateral axis
    static constexpr int WARPS = NT / 32;
    static constexpr int NQ = 7, NF = 6, NB = 5, NW = 4;             // ring lengths (cells)
    static constexpr int OFF_Q = 0, OFF_F = OFF_Q + NQ * 7, OFF_B = OFF_F + NF * 7, OFF_W = OFF_B + NB;
    static constexpr int PER_THREAD = OFF_W + NW * 4;
    static constexpr size_t SMEM = (size_t)NT * PER_THREAD * sizeof(double);
    static_assert(PER_THREAD == 112, "ring budget");
    // lines of a warp run along the lane axis, the warps of a CTA along the row axis: (x, z) for the y sweep, (x, y) for
    // the z sweep -- lanes along x, every access coalesced -- and (y, z) for the x sweep, where a warp's 32 lines are
    // Nx apart: its loads and stores touch one 32-byte sector per lane, which the next three march steps use up (L2)
    W5_HD static int lane_axis(int dir) { return dir == 0 ? 1 : 0; }
    W5_HD static int row_axis(int dir) { return dir == 2 ? 1 : 2; }
    // grid: (groups of 32 lines along the lane axis, groups of 8 along the row axis, segments along the sweep)
    W5_HD static void grid(const Dims &d, int dir, int nseg, int g[3])
    {
        const int al = lane_axis(dir), ar = row_axis(dir);
        g[0] = (line_count(dimN(d, al), d.bc[al]) + 31) / 32;
        g[1] = (line_count(dimN(d, ar), d.bc[ar]) + WARPS - 1) / WARPS;
        g[2] = nseg;
    }
    // segments per line: the count whose last wave of CTAs is fullest, a segment's prologue (5 cells + the shared
    // face, ~1.5 march steps) charged against its length
    static int segments(const Dims &d, int dir, int ctas_per_wave)   // host only
    {
        int g[3];
        grid(d, dir, 1, g);
        const int base = g[0] * g[1], ni = dimN(d, dir) - 2 * NB3;
        int best = 1;
        double best_cost = 1e300;
        for (int ns = 1; ns <= ni / 8 && ns <= 64; ++ns) {
            const int per = (ni + ns - 1) / ns;
            const int used = (ni + per - 1) / per;                       // segments that own cells
            const long waves = ((long)base * used + ctas_per_wave - 1) / ctas_per_wave;
            const double cost = (double)waves * (per + 1.5);
            if (cost < best_cost * (1.0 - 1e-9)) { best_cost = cost; best = ns; }
        }
        return best;
    }
};

W5_HD void prefetch_l2(const double *p)
{
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// 8-byte asynchronous copy global -> shared (LDGSTS); on the host the copy happens at once
W5_HD void cp_async_f64(double *dst_smem, const double *src)
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
#else
    *dst_smem = *src;
#endif
}
W5_HD void cp_async_commit_wait_all()
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#endif
}
W5_HD void cp_async_commit()
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
W5_HD void cp_async_wait_all()
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

W5_HD void red_add(double *p, double v)
{
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);                                   // result unused: a fire-and-forget RED, no load latency in the march
#else
    *p += v;
#endif
}

template <int DIR>
W5_HD void sweep_march(const SweepArgs &a, int nseg, int bx, int by, int bz, int tid, double *smem)
{
    using M = MarchGeo;
    const Dims &d = a.d;
    const int al = M::lane_axis(DIR), ar = M::row_axis(DIR);
    const int n = dimN(d, DIR), N1 = dimN(d, al), N2 = dimN(d, ar);
    const int p1 = line_lo(d.bc[al]) + bx * 32 + (tid & 31);
    const int p2 = line_lo(d.bc[ar]) + by * M::WARPS + (tid >> 5);
    if (p1 > N1 - 1 - line_lo(d.bc[al]) || p2 > N2 - 1 - line_lo(d.bc[ar])) return;
    const bool interior = p1 >= NB3 && p1 <= N1 - NB3 - 1 && p2 >= NB3 && p2 <= N2 - NB3 - 1;
    // segment: cells [cA, cB) of the interior 3 .. n-4; faces cA-1 .. cB-1, and face n-3 where flux_at clamps onto it
    const int ni = n - 2 * NB3;
    const int per = (ni + nseg - 1) / nseg;
    const int cA = NB3 + bz * per, cB = cA + per < n - NB3 ? cA + per : n - NB3;
    if (cA >= cB) return;
    const int fA = cA - 1;
    const int fLast = (cB == n - NB3 && d.bc[DIR] != 1) ? n - NB3 : cB - 1;
    const size_t sstride = DIR == 0 ? 1 : (DIR == 1 ? (size_t)d.Nx : (size_t)d.Nx * d.Ny);
    const size_t g0 = DIR == 0 ? vidx(d, 0, p1, p2) : (DIR == 1 ? vidx(d, p1, 0, p2) : vidx(d, p1, p2, 0));   // cell 0 of the line
    double *sm = smem + tid;                           // entry e of this thread: sm[e * NT]
    auto Q = [&](int slot, int m) -> double & { return sm[(size_t)(M::OFF_Q + slot * 7 + m) * M::NT]; };
    auto F = [&](int slot, int m) -> double & { return sm[(size_t)(M::OFF_F + slot * 7 + m) * M::NT]; };
    auto B = [&](int slot) -> double & { return sm[(size_t)(M::OFF_B + slot) * M::NT]; };
    auto W = [&](int slot, int m) -> double & { return sm[(size_t)(M::OFF_W + slot * 4 + m) * M::NT]; };

    auto line_cell = [&](int c) {                      // stencil cell past the end of the line: quirk Q3
        if (c >= n) c = d.bc[DIR] == 1 ? c - ni : n - 1;
        return c >= n ? n - 1 : c;
    };
    // raw state of cell c -> its ring slots (q7 and B1), asynchronously
    auto fetch = [&](int c, int sq, int sb) {
        const size_t g = g0 + sstride * (size_t)line_cell(c);
        cp_async_f64(&Q(sq, 0), a.q[0] + g);
        cp_async_f64(&Q(sq, 1), a.q[sweep_in(DIR, 1)] + g);
        cp_async_f64(&Q(sq, 2), a.q[sweep_in(DIR, 2)] + g);
        cp_async_f64(&Q(sq, 3), a.q[sweep_in(DIR, 3)] + g);
        cp_async_f64(&B(sb), a.q[sweep_in(DIR, 4)] + g);
        cp_async_f64(&Q(sq, 4), a.q[sweep_in(DIR, 5)] + g);
        cp_async_f64(&Q(sq, 5), a.q[sweep_in(DIR, 6)] + g);
        cp_async_f64(&Q(sq, 6), a.q[7] + g);
    };
    // cell quantities of a cell whose raw state sits in the ring: flux -> F ring, 1/rho and wave speeds -> W ring
    auto cell_in = [&](int sq, int sf, int sb, int sw) {
        double q[8];
        q[0] = Q(sq, 0); q[1] = Q(sq, 1); q[2] = Q(sq, 2); q[3] = Q(sq, 3);
        q[4] = B(sb); q[5] = Q(sq, 4); q[6] = Q(sq, 5); q[7] = Q(sq, 6);
        CellOut o;
        cell_quantities(q, a.gam, o);
#pragma unroll
        for (int m = 0; m < 7; ++m) F(sf, m) = o.F[m];
        W(sw, 0) = fast_rcp(q[0]);                     // v = M / rho is re-formed at the face: M * (1/rho), as cell_quantities does
        W(sw, 1) = o.lf; W(sw, 2) = o.la; W(sw, 3) = o.ls;
    };

    // cell c has the running index r = c - (fA-2) and lives in slot r mod NQ / NF / NB / NW of the four rings
    for (int r = 0; r < 5; ++r) fetch(fA - 2 + r, r, r);               // prologue: cells fA-2 .. fA+2
    cp_async_commit_wait_all();
    for (int r = 0; r < 5; ++r) cell_in(r, r, r, r & 3);
    fetch(fA + 3, 5, 0);                               // the first step's cell (B1 slot of cell fA-2, no longer needed)
    cp_async_commit();
    int rq = 0, rf = 0, rb = 2, rw = 2;                // slots of cell f-2 (Q, F rings) and of cell f (B1, W rings)
    double fprev[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int f = fA; f <= fLast; ++f) {
        cp_async_wait_all();                           // cell f+3 has landed (issued a step ago)
        const int q3 = rq + 5 >= M::NQ ? rq + 5 - M::NQ : rq + 5, q4 = rq + 6 >= M::NQ ? rq + 6 - M::NQ : rq + 6;
        const int b3 = rb + 3 >= M::NB ? rb + 3 - M::NB : rb + 3, b4 = rb + 4 >= M::NB ? rb + 4 - M::NB : rb + 4;
        cell_in(q3, rf + 5 >= M::NF ? rf + 5 - M::NF : rf + 5, b3, (rw + 3) & 3);
        if (f < fLast) { fetch(f + 4, q4, b4); cp_async_commit(); }    // next step's cell, into slots nothing reads now
        const int qL = rq + 2 >= M::NQ ? rq + 2 - M::NQ : rq + 2, qR = rq + 3 >= M::NQ ? rq + 3 - M::NQ : rq + 3;
        const int bR = rb + 1 >= M::NB ? rb + 1 - M::NB : rb + 1, wR = (rw + 1) & 3;
        const double riL = W(rw, 0), riR = W(wR, 0);
        const double v1L = Q(qL, 1) * riL, v2L = Q(qL, 2) * riL, v3L = Q(qL, 3) * riL;
        const double v1R = Q(qR, 1) * riR, v2R = Q(qR, 2) * riR, v3R = Q(qR, 3) * riR;
        const double B1L = B(rb), B1R = B(bR);
        FaceCoef fc;
        face_coefficients(Q(qL, 0), Q(qR, 0), v1L, v1R, v2L, v2R, v3L, v3R, B1L, B1R, Q(qL, 4), Q(qR, 4), Q(qL, 5), Q(qR, 5),
                          Q(qL, 6), Q(qR, 6), a.gam, a.g2, fc);
        double amax[7];
        face_amax(v1L, W(rw, 1), W(rw, 2), W(rw, 3), v1R, W(wR, 1), W(wR, 2), W(wR, 3), amax);
        auto load = [&](int k, double Fk[7], double x[7]) {
            const int sq = rq + k >= M::NQ ? rq + k - M::NQ : rq + k, sf = rf + k >= M::NF ? rf + k - M::NF : rf + k;
#pragma unroll
            for (int m = 0; m < 7; ++m) {
                Fk[m] = F(sf, m);
                x[m] = Q(sq, m);
            }
        };
        double fh[7], Fs[7];
        RegStash st;
        BackCoef bc;
        split_back(fc, bc);
        face_flux(fc, amax, a.eps, load, st, Fs);
        back_project(bc, Fs, fh);
        const double bv2 = B1L * v2L + B1R * v2R;
        const double bv3 = B1L * v3L + B1R * v3R;
        const size_t g = g0 + sstride * (size_t)f;
        a.bsA[g] = fma(0.5, bv2, fh[4]);
        a.bsB[g] = fma(0.5, bv3, fh[5]);
        const double fcur[5] = {fh[0], fh[1], fh[2], fh[3], fh[6]};
        if (f > fA && f < cB && interior) {            // cell f: faces f-1 (fprev) and f
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                const double dfl = (fcur[m] - fprev[m]) * a.inv_h;
                if (DIR == 0) a.rhs[m][g] = dfl; else red_add(a.rhs[sweep_out(DIR, m)] + g, dfl);
            }
        }
#pragma unroll
        for (int m = 0; m < 5; ++m) fprev[m] = fcur[m];
        rq = rq + 1 >= M::NQ ? 0 : rq + 1;
        rf = rf + 1 >= M::NF ? 0 : rf + 1;
        rb = rb + 1 >= M::NB ? 0 : rb + 1;
        rw = (rw + 1) & 3;
    }
}

// ------------------------------------------------------------------ x sweep: a warp marches its own lines
// Along x the stencil runs along the lanes, so here groups of LPL lanes own a line (32 / LPL lines per warp): per march
// step a group computes LPL new cells (one contiguous load of LPL doubles per component), LPL faces and LPL flux
// differences; the cells live in a warp-private ring in shared memory (a cell is read by the six faces around it), the
// fhat of the faces in a double-buffered strip so that a group's first lane finds its left face from the previous
// step.  Warps never meet a CTA barrier -- the three phases of a step are separated by __syncwarp only -- and drift
// freely like the threads of sweep_march.  LPL = 8, 16 or 32 is chosen per grid so that the last step of a line is
// as full as possible (a line of nx cells has nx+1 faces: with 32 lanes per line 256 cells would take 9 steps for
// 8.03 steps of work).
template <int LPL> struct XWarpGeo {
    static constexpr int NT = 256, WARPS = NT / 32;
    static constexpr int LPW = 32 / LPL;               // lines per warp
    static constexpr int RL = 2 * LPL;                 // ring slots per line (power of two >= LPL + 5)
    // distance of the lines' rings in shared memory: the 16 lanes of a half-warp must fall into 16 different 8-byte
    // banks.  With 8 lanes per line the two lines of a half-warp sit at the same ring phase, so their rings are placed
    // 8 banks apart (24 = 16 + 8); with 16 or 32 lanes per line a half-warp reads 16 consecutive slots of one ring.
    static constexpr int RLS = LPL == 8 ? 24 : RL;
    static constexpr int RING = LPW * RLS;             // cell slots per warp
    static constexpr int NFB = 5;
    static constexpr int FBL = LPL + 1;                // fhat strip of a line: entry l+1 = face of lane l
    static constexpr int FBW = LPW * FBL;
    static constexpr int OFF_STAGE = CQ_COUNT * RING + 2 * NFB * FBW;   // raw state of the next step's cells: 2 x 8 x 32
    static constexpr int PER_WARP = OFF_STAGE + 2 * 8 * 32;
    static constexpr size_t SMEM = (size_t)WARPS * PER_WARP * sizeof(double);
    static_assert(LPL >= 8 && RL >= LPL + 5, "the prologue needs five lanes and the ring LPL + 5 slots");
    // grid: (groups of 8 * LPW lines along y, lines along z, segments along x)
    W5_HD static void grid(const Dims &d, int nseg, int g[3])
    {
        g[0] = (line_count(d.Ny, d.bc[1]) + WARPS * LPW - 1) / (WARPS * LPW);
        g[1] = line_count(d.Nz, d.bc[2]);
        g[2] = nseg;
    }
};

// host only: segments per line (lines are cut only when there are too few of them to fill the GPU) and lanes per line
inline void xwarp_plan(const Dims &d, int ctas_per_wave, int &nseg, int &lpl)
{
    const int ni = d.Nx - 2 * NB3;
    const int lines = line_count(d.Ny, d.bc[1]) * line_count(d.Nz, d.bc[2]);
    nseg = 1;
    while ((lines + 31) / 32 * nseg < 2 * ctas_per_wave && ni / (nseg + 1) >= 32) ++nseg;
    const int per = (ni + nseg - 1) / nseg;
    const int nf = per + 1 + (d.bc[0] != 1 ? 1 : 0);   // faces of the longest segment
    lpl = 32;
    double best = 1e300;
    for (int l = 32; l >= 8; l /= 2) {                 // steps x lanes per useful face; ties go to the wider group
        const double cost = (double)((nf + l - 1) / l) * l / nf;
        if (cost < best - 1e-12) { best = cost; lpl = l; }
    }
}

// geometry of a lane's line and segment, shared by the three phases
template <int LPL> struct XWarpPos {
    using X = XWarpGeo<LPL>;
    int sub, ll;              // line of the warp, lane within the line's group
    int j, k, n, ni, cA, cB, fA, fLast, nchunks;
    bool ok, interior;
    size_t g0;
    W5_HD XWarpPos(const Dims &d, int nseg, int bx, int by, int bz, int warp, int lane)
    {
        sub = lane / LPL; ll = lane % LPL;
        j = line_lo(d.bc[1]) + (bx * X::WARPS + warp) * X::LPW + sub;
        k = line_lo(d.bc[2]) + by;
        n = d.Nx; ni = n - 2 * NB3;
        const int per = (ni + nseg - 1) / nseg;
        cA = NB3 + bz * per;
        cB = cA + per < n - NB3 ? cA + per : n - NB3;
        fA = cA - 1;
        fLast = (cB == n - NB3 && d.bc[0] != 1) ? n - NB3 : cB - 1;
        nchunks = cA < cB ? (fLast - fA + LPL) / LPL : 0;              // the same for every lane of the CTA
        ok = j <= d.Ny - 1 - line_lo(d.bc[1]) && cA < cB;
        interior = j >= NB3 && j <= d.Ny - NB3 - 1 && k >= NB3 && k <= d.Nz - NB3 - 1;
        g0 = ok ? vidx(d, 0, j, k) : 0;
    }
    W5_HD int slot(int c) const { return sub * X::RLS + ((c - (fA - 2)) & (X::RL - 1)); }
    W5_HD int fslot(int l) const { return sub * X::FBL + l; }
    W5_HD int line_cell(int c, int bc) const
    {
        if (c >= n) c = bc == 1 ? c - ni : n - 1;
        return c >= n ? n - 1 : c;
    }
};

// cell quantities of cell c from its raw state q (sweep frame = physical frame for x) -> the line's ring
template <int LPL>
W5_HD void xwarp_cell_from(const SweepArgs &a, const XWarpPos<LPL> &P, int c, const double q[8], double *cb)
{
    constexpr int CBS = XWarpGeo<LPL>::RING;
    CellOut o;
    cell_quantities(q, a.gam, o);
    const int i = P.slot(c);
#pragma unroll
    for (int m = 0; m < 7; ++m) cb[(CQ_F0 + m) * CBS + i] = o.F[m];
    cb[(CQ_Q0 + 0) * CBS + i] = q[0];
    cb[(CQ_Q0 + 1) * CBS + i] = q[1];
    cb[(CQ_Q0 + 2) * CBS + i] = q[2];
    cb[(CQ_Q0 + 3) * CBS + i] = q[3];
    cb[(CQ_Q0 + 4) * CBS + i] = q[5];
    cb[(CQ_Q0 + 5) * CBS + i] = q[6];
    cb[(CQ_Q0 + 6) * CBS + i] = q[7];
    cb[CQ_V1 * CBS + i] = o.v1;
    cb[CQ_V2 * CBS + i] = o.v2;
    cb[CQ_V3 * CBS + i] = o.v3;
    cb[CQ_B1 * CBS + i] = q[4];
    cb[CQ_LF * CBS + i] = o.lf;
    cb[CQ_LA * CBS + i] = o.la;
    cb[CQ_LS * CBS + i] = o.ls;
}

template <int LPL>
W5_HD void xwarp_cell(const SweepArgs &a, const XWarpPos<LPL> &P, int c, double *cb)
{
    const size_t g = P.g0 + (size_t)P.line_cell(c, a.d.bc[0]);
    double q[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) q[m] = a.q[m][g];
    xwarp_cell_from<LPL>(a, P, c, q, cb);
}

// phase 0 of march step `chunk`: the new cells (and, before the first step, the five cells in front of the first face).
// From the second step on a lane finds the raw state of its cell in its private staging slots, where cp.async put it
// while the previous step's faces were computed, and sends for the cell of the step after.
template <int LPL>
W5_HD void xwarp_cells(const SweepArgs &a, int nseg, int bx, int by, int bz, int warp, int lane, int chunk, double *wsm)
{
    using X = XWarpGeo<LPL>;
    const XWarpPos<LPL> P(a.d, nseg, bx, by, bz, warp, lane);
    if (!P.ok || chunk >= P.nchunks) return;
    double *stage = wsm + X::OFF_STAGE + lane;          // entry m of buffer b: stage[(b * 8 + m) * 32]
    const int c = P.fA + 3 + LPL * chunk + P.ll;
    if (chunk == 0) {
        if (P.ll < 5) xwarp_cell(a, P, P.fA - 2 + P.ll, wsm);
        if (c <= P.fLast + 3) xwarp_cell(a, P, c, wsm);
    } else if (c <= P.fLast + 3) {
        cp_async_wait_all();
        double q[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) q[m] = stage[((chunk & 1) * 8 + m) * 32];
        xwarp_cell_from<LPL>(a, P, c, q, wsm);
    }
    if (c + LPL <= P.fLast + 3 && chunk + 1 < P.nchunks) {
        const size_t gp = P.g0 + (size_t)P.line_cell(c + LPL, a.d.bc[0]);
#pragma unroll
        for (int m = 0; m < 8; ++m) cp_async_f64(&stage[(((chunk + 1) & 1) * 8 + m) * 32], a.q[m] + gp);
        cp_async_commit();
    }
}

// phase 1: the faces
template <int LPL>
W5_HD void xwarp_faces(const SweepArgs &a, int nseg, int bx, int by, int bz, int warp, int lane, int chunk, double *wsm)
{
    using X = XWarpGeo<LPL>;
    constexpr int CBS = X::RING;
    const XWarpPos<LPL> P(a.d, nseg, bx, by, bz, warp, lane);
    if (!P.ok || chunk >= P.nchunks) return;
    const int f = P.fA + LPL * chunk + P.ll;
    if (f > P.fLast) return;
    const double *cb = wsm;
    double *fb = wsm + CQ_COUNT * X::RING + (chunk & 1) * X::NFB * X::FBW;
    const int iL = P.slot(f), iR = P.slot(f + 1);
    FaceCoef fc;
    face_coefficients(cb[CQ_Q0 * CBS + iL], cb[CQ_Q0 * CBS + iR], cb[CQ_V1 * CBS + iL], cb[CQ_V1 * CBS + iR],
                      cb[CQ_V2 * CBS + iL], cb[CQ_V2 * CBS + iR], cb[CQ_V3 * CBS + iL], cb[CQ_V3 * CBS + iR],
                      cb[CQ_B1 * CBS + iL], cb[CQ_B1 * CBS + iR], cb[(CQ_Q0 + 4) * CBS + iL], cb[(CQ_Q0 + 4) * CBS + iR],
                      cb[(CQ_Q0 + 5) * CBS + iL], cb[(CQ_Q0 + 5) * CBS + iR], cb[(CQ_Q0 + 6) * CBS + iL],
                      cb[(CQ_Q0 + 6) * CBS + iR], a.gam, a.g2, fc);
    double amax[7];
    face_amax(cb[CQ_V1 * CBS + iL], cb[CQ_LF * CBS + iL], cb[CQ_LA * CBS + iL], cb[CQ_LS * CBS + iL],
              cb[CQ_V1 * CBS + iR], cb[CQ_LF * CBS + iR], cb[CQ_LA * CBS + iR], cb[CQ_LS * CBS + iR], amax);
    auto load = [&](int kk, double F[7], double x[7]) {
        const int ii = P.slot(f - 2 + kk);
#pragma unroll
        for (int m = 0; m < 7; ++m) {
            F[m] = cb[(CQ_F0 + m) * CBS + ii];
            x[m] = cb[(CQ_Q0 + m) * CBS + ii];
        }
    };
    double fh[7], Fs[7];
    RegStash st;
    BackCoef bc;
    split_back(fc, bc);
    face_flux(fc, amax, a.eps, load, st, Fs);
    back_project(bc, Fs, fh);
    const int io = P.fslot(P.ll + 1);
    fb[0 * X::FBW + io] = fh[0];
    fb[1 * X::FBW + io] = fh[1];
    fb[2 * X::FBW + io] = fh[2];
    fb[3 * X::FBW + io] = fh[3];
    fb[4 * X::FBW + io] = fh[6];
    const double bv2 = cb[CQ_B1 * CBS + iL] * cb[CQ_V2 * CBS + iL] + cb[CQ_B1 * CBS + iR] * cb[CQ_V2 * CBS + iR];
    const double bv3 = cb[CQ_B1 * CBS + iL] * cb[CQ_V3 * CBS + iL] + cb[CQ_B1 * CBS + iR] * cb[CQ_V3 * CBS + iR];
    a.bsA[P.g0 + f] = fma(0.5, bv2, fh[4]);
    a.bsB[P.g0 + f] = fma(0.5, bv3, fh[5]);
}

// phase 2: the flux differences (the x sweep runs first in a stage: plain stores)
template <int LPL>
W5_HD void xwarp_update(const SweepArgs &a, int nseg, int bx, int by, int bz, int warp, int lane, int chunk, const double *wsm)
{
    using X = XWarpGeo<LPL>;
    const XWarpPos<LPL> P(a.d, nseg, bx, by, bz, warp, lane);
    if (!P.ok || chunk >= P.nchunks || !P.interior) return;
    const int c = P.fA + LPL * chunk + P.ll;           // cell c: faces c-1 and c
    if (c <= P.fA || c >= P.cB) return;
    const double *cur = wsm + CQ_COUNT * X::RING + (chunk & 1) * X::NFB * X::FBW;
    const double *prv = wsm + CQ_COUNT * X::RING + ((chunk & 1) ^ 1) * X::NFB * X::FBW;
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        const double left = P.ll > 0 ? cur[m * X::FBW + P.fslot(P.ll)] : prv[m * X::FBW + P.fslot(LPL)];
        a.rhs[m][P.g0 + c] = (cur[m * X::FBW + P.fslot(P.ll + 1)] - left) * a.inv_h;
    }
}

// ------------------------------------------------------------------ constrained transport
struct CtArgs3 {
    const double *fl[6];       // f_y, f_z, g_z, g_x, h_x, h_y (as written by the sweeps: no ghost fill)
    double *O[3];              // Ox, Oy, Oz
    const double *b0[3], *b1[3], *b2[3], *b3[3];   // face fields of the step input and of stages 1-3
    double *bn[3];             // face fields written by this stage
    const double *dtp;
    Dims d;
    double c;                  // stage coefficient 1/2, 1/2, 1, 1/6
    double idx_, idy_, idz_;   // 1/dx, 1/dy, 1/dz
    int stage;
};

// corner_fluxes (Weno5_2D.jl:319-342) in 3-D: the three EMFs on the edges of cell (i,j,k); zero outside 1 .. N-2
W5_HD void emf_cell(const CtArgs3 &a, int i, int j, int k)
{
    const Dims &d = a.d;
    const size_t g = vidx(d, i, j, k);
    const bool inner = i >= 1 && i < d.Nx - 1 && j >= 1 && j < d.Ny - 1 && k >= 1 && k < d.Nz - 1;
    double ox = 0.0, oy = 0.0, oz = 0.0;
    const double *fy = a.fl[0], *fz = a.fl[1], *gz = a.fl[2], *gx = a.fl[3], *hx = a.fl[4], *hy = a.fl[5];
    // away from the ghost layers flux_at is the identity on all twelve reads (bit-identical, a third of the instructions)
    const bool deep = i >= NB3 && i < d.Nx - NB3 - 1 && j >= NB3 && j < d.Ny - NB3 - 1 && k >= NB3 && k < d.Nz - NB3 - 1;
    if (deep) {
        const size_t sy = (size_t)d.Nx, sz = (size_t)d.Nx * d.Ny;
        ox = 0.5 * (hy[g] + hy[g + sy] - gz[g] - gz[g + sz]);
        oy = 0.5 * (fz[g] + fz[g + sz] - hx[g] - hx[g + 1]);
        oz = 0.5 * (gx[g] + gx[g + 1] - fy[g] - fy[g + sy]);
    } else if (inner) {
        ox = 0.5 * (flux_at(hy, d, i, j, k) + flux_at(hy, d, i, j + 1, k) - flux_at(gz, d, i, j, k) - flux_at(gz, d, i, j, k + 1));
        oy = 0.5 * (flux_at(fz, d, i, j, k) + flux_at(fz, d, i, j, k + 1) - flux_at(hx, d, i, j, k) - flux_at(hx, d, i + 1, j, k));
        oz = 0.5 * (flux_at(gx, d, i, j, k) + flux_at(gx, d, i + 1, j, k) - flux_at(fy, d, i, j, k) - flux_at(fy, d, i, j + 1, k));
    }
    a.O[0][g] = ox; a.O[1][g] = oy; a.O[2][g] = oz;
}

// face-field RK update (Weno5_2D.jl:150-151, :195-196, :240-241, :284-285 with the stage-4 combination :279-281)
W5_HD void face_update_cell(const CtArgs3 &a, double dt, int i, int j, int k)
{
    const Dims &d = a.d;
    const size_t g = vidx(d, i, j, k);
    const bool inner = i >= 1 && i < d.Nx - 1 && j >= 1 && j < d.Ny - 1 && k >= 1 && k < d.Nz - 1;
    const double ox = a.O[0][g], oy = a.O[1][g], oz = a.O[2][g];
    const double ox_j = inner ? a.O[0][vidx(d, i, j - 1, k)] : 0.0, ox_k = inner ? a.O[0][vidx(d, i, j, k - 1)] : 0.0;
    const double oy_i = inner ? a.O[1][vidx(d, i - 1, j, k)] : 0.0, oy_k = inner ? a.O[1][vidx(d, i, j, k - 1)] : 0.0;
    const double oz_i = inner ? a.O[2][vidx(d, i - 1, j, k)] : 0.0, oz_j = inner ? a.O[2][vidx(d, i, j - 1, k)] : 0.0;
    double bb[3];
#pragma unroll
    for (int m = 0; m < 3; ++m)
        bb[m] = a.stage < 4 ? a.b0[m][g] : 1.0 / 3 * (-a.b0[m][g] + a.b1[m][g] + 2 * a.b2[m][g] + a.b3[m][g]);
    const double cx = a.c * dt * a.idx_, cy = a.c * dt * a.idy_, cz = a.c * dt * a.idz_;
    a.bn[0][g] = bb[0] - cy * (oz - oz_j) + cz * (oy - oy_k);
    a.bn[1][g] = bb[1] - cz * (ox - ox_k) + cx * (oz - oz_i);
    a.bn[2][g] = bb[2] - cx * (oy - oy_i) + cy * (ox - ox_j);
}

// EMFs and face update in one pass (opt-in, WENO5_3D_CT=fused; measured slower than the two passes: the nine EMF
// evaluations per cell cost more than the round trip of the three EMF volumes saves): a face needs the EMFs of its four edges -- for the three faces of a cell nine
// evaluations, whose 36 reads of the transport fluxes are shared between neighbouring threads (L1) -- so the three EMF
// volumes are never written or read.  The same expressions, evaluated in the same order, as emf_cell +
// face_update_cell: bit-identical.
W5_HD double emf_at(const CtArgs3 &a, int m, int i, int j, int k)
{
    const Dims &d = a.d;
    if (!(i >= 1 && i < d.Nx - 1 && j >= 1 && j < d.Ny - 1 && k >= 1 && k < d.Nz - 1)) return 0.0;
    const double *fy = a.fl[0], *fz = a.fl[1], *gz = a.fl[2], *gx = a.fl[3], *hx = a.fl[4], *hy = a.fl[5];
    if (m == 0) return 0.5 * (flux_at(hy, d, i, j, k) + flux_at(hy, d, i, j + 1, k) - flux_at(gz, d, i, j, k) - flux_at(gz, d, i, j, k + 1));
    if (m == 1) return 0.5 * (flux_at(fz, d, i, j, k) + flux_at(fz, d, i, j, k + 1) - flux_at(hx, d, i, j, k) - flux_at(hx, d, i + 1, j, k));
    return 0.5 * (flux_at(gx, d, i, j, k) + flux_at(gx, d, i + 1, j, k) - flux_at(fy, d, i, j, k) - flux_at(fy, d, i, j + 1, k));
}

W5_HD void ct_cell(const CtArgs3 &a, double dt, int i, int j, int k)
{
    const Dims &d = a.d;
    const size_t g = vidx(d, i, j, k);
    const bool inner = i >= 1 && i < d.Nx - 1 && j >= 1 && j < d.Ny - 1 && k >= 1 && k < d.Nz - 1;
    double ox, oy, oz, ox_j, ox_k, oy_i, oy_k, oz_i, oz_j;
    // one more layer away from the ghost cells than emf_cell's fast path: the shifted EMFs read index - 1 too
    const bool deep = i > NB3 && i < d.Nx - NB3 - 1 && j > NB3 && j < d.Ny - NB3 - 1 && k > NB3 && k < d.Nz - NB3 - 1;
    if (deep) {
        const double *fy = a.fl[0], *fz = a.fl[1], *gz = a.fl[2], *gx = a.fl[3], *hx = a.fl[4], *hy = a.fl[5];
        const size_t sy = (size_t)d.Nx, sz = (size_t)d.Nx * d.Ny;
        ox = 0.5 * (hy[g] + hy[g + sy] - gz[g] - gz[g + sz]);
        ox_j = 0.5 * (hy[g - sy] + hy[g] - gz[g - sy] - gz[g - sy + sz]);
        ox_k = 0.5 * (hy[g - sz] + hy[g - sz + sy] - gz[g - sz] - gz[g]);
        oy = 0.5 * (fz[g] + fz[g + sz] - hx[g] - hx[g + 1]);
        oy_i = 0.5 * (fz[g - 1] + fz[g - 1 + sz] - hx[g - 1] - hx[g]);
        oy_k = 0.5 * (fz[g - sz] + fz[g] - hx[g - sz] - hx[g - sz + 1]);
        oz = 0.5 * (gx[g] + gx[g + 1] - fy[g] - fy[g + sy]);
        oz_i = 0.5 * (gx[g - 1] + gx[g] - fy[g - 1] - fy[g - 1 + sy]);
        oz_j = 0.5 * (gx[g - sy] + gx[g - sy + 1] - fy[g - sy] - fy[g]);
    } else {
        ox = emf_at(a, 0, i, j, k); oy = emf_at(a, 1, i, j, k); oz = emf_at(a, 2, i, j, k);
        ox_j = inner ? emf_at(a, 0, i, j - 1, k) : 0.0; ox_k = inner ? emf_at(a, 0, i, j, k - 1) : 0.0;
        oy_i = inner ? emf_at(a, 1, i - 1, j, k) : 0.0; oy_k = inner ? emf_at(a, 1, i, j, k - 1) : 0.0;
        oz_i = inner ? emf_at(a, 2, i - 1, j, k) : 0.0; oz_j = inner ? emf_at(a, 2, i, j - 1, k) : 0.0;
    }
    double bb[3];
#pragma unroll
    for (int m = 0; m < 3; ++m)
        bb[m] = a.stage < 4 ? a.b0[m][g] : 1.0 / 3 * (-a.b0[m][g] + a.b1[m][g] + 2 * a.b2[m][g] + a.b3[m][g]);
    const double cx = a.c * dt * a.idx_, cy = a.c * dt * a.idy_, cz = a.c * dt * a.idz_;
    a.bn[0][g] = bb[0] - cy * (oz - oz_j) + cz * (oy - oy_k);
    a.bn[1][g] = bb[1] - cz * (ox - ox_k) + cx * (oz - oz_i);
    a.bn[2][g] = bb[2] - cx * (oy - oy_i) + cy * (ox - ox_j);
}

// boundaries! on a face array, in place: every location the rules write reads the location they copy (or 0)
W5_HD bool face_is_ghost(const Dims &d, int i, int j, int k)
{
    const int c[3] = {i, j, k};
    bool gh = false;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        const int N = dimN(d, ax);
        gh = gh || (d.bc[ax] == 1 ? (c[ax] < NB3 || c[ax] > N - NB3 - 1) : (c[ax] < 2 || c[ax] > N - 3));
    }
    return gh;
}
W5_HD void face_bc_cell(double *f, const Dims &d, int i, int j, int k)
{
    if (face_is_ghost(d, i, j, k)) f[vidx(d, i, j, k)] = flux_at(f, d, i, j, k);
}

// ------------------------------------------------------------------ cell update of one stage
struct UpdArgs3 {
    const double *q0[9], *q1[9], *q2[9], *q3[9];   // step input and stages 1-3 (public component order, 7 = S, 8 = E)
    const double *rhs[5];
    const double *bn[3];       // this stage's face fields
    double *qn[9];
    const double *dtp;
    Dims d;
    double c, gam;
    int stage, roundtrip;
};

// inverse of map_cell: the positions (own index first) whose value is that of interior cell idx after boundaries!
W5_HD int cell_images(int idx, int N, int bc, int out[4])
{
    const int n = N - 2 * NB3;
    int cnt = 0;
    out[cnt++] = idx;
    if (bc == 1) {
        if (idx < 2 * NB3) out[cnt++] = idx + n;          // the first three interior cells appear behind the last
        if (idx > N - 2 * NB3 - 1) out[cnt++] = idx - n;  // and the last three in front of the first
    } else {
        if (idx == NB3) for (int g = 0; g < NB3; ++g) out[cnt++] = g;
        if (idx == N - NB3 - 1) for (int g = 0; g < NB3; ++g) out[cnt++] = N - NB3 + g;
    }
    return cnt;
}

// RK combination of rho, M, E (Weno5_2D.jl:139-146 ... :279), interpolateB_face2center (:344-356) for all three field
// components, and what survives of ES_Switch (:1734-1768): S = E2S(q), E = S2E(q) when the round trip is on
W5_HD void update_cell(const UpdArgs3 &a, double dt, int i, int j, int k)
{
    const Dims &d = a.d;
    const size_t g = vidx(d, i, j, k);
    const int comp[5] = {0, 1, 2, 3, 8};
    double v[5];
    const double cdt = a.c * dt;
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        const int c = comp[m];
        const double b = a.stage < 4 ? a.q0[c][g] : 1.0 / 3 * (-a.q0[c][g] + a.q1[c][g] + 2 * a.q2[c][g] + a.q3[c][g]);
        v[m] = b - cdt * a.rhs[m][g];
    }
    const double *x = a.bn[0], *y = a.bn[1], *z = a.bn[2];
    const double Bx = (-x[vidx(d, i - 2, j, k)] + 9.0 * x[vidx(d, i - 1, j, k)] + 9.0 * x[g] - x[vidx(d, i + 1, j, k)]) / 16.0;
    const double By = (-y[vidx(d, i, j - 2, k)] + 9.0 * y[vidx(d, i, j - 1, k)] + 9.0 * y[g] - y[vidx(d, i, j + 1, k)]) / 16.0;
    const double Bz = (-z[vidx(d, i, j, k - 2)] + 9.0 * z[vidx(d, i, j, k - 1)] + 9.0 * z[g] - z[vidx(d, i, j, k + 1)]) / 16.0;
    double S, E2;
    e2s_s2e(v[0], v[1], v[2], v[3], Bx, By, Bz, v[4], a.gam, S, E2);
    const double out[9] = {v[0], v[1], v[2], v[3], Bx, By, Bz, S, a.roundtrip ? E2 : v[4]};
#pragma unroll
    for (int c = 0; c < 9; ++c) a.qn[c][g] = out[c];
    // boundaries!(q) by scatter: the cells next to a boundary also store their ghost images (the inverse of map_cell),
    // so no separate ghost-fill pass runs over the volume after the update
    int xi[4], yi[4], zi[4];
    const int nxi = cell_images(i, d.Nx, d.bc[0], xi), nyi = cell_images(j, d.Ny, d.bc[1], yi),
              nzi = cell_images(k, d.Nz, d.bc[2], zi);
    if (nxi + nyi + nzi > 3) {
        for (int cz = 0; cz < nzi; ++cz)
            for (int cy = 0; cy < nyi; ++cy)
                for (int cx = 0; cx < nxi; ++cx) {
                    if (cx + cy + cz == 0) continue;
                    const size_t gi = vidx(d, xi[cx], yi[cy], zi[cz]);
#pragma unroll
                    for (int c = 0; c < 9; ++c) a.qn[c][gi] = out[c];
                }
    }
}

// boundaries!(q) in place: a ghost cell reads the interior cell the rules map it to (interior cells are not written)
W5_HD void cell_bc(double *f, const Dims &d, int i, int j, int k)
{
    const int si = map_cell(i, d.Nx, d.bc[0]), sj = map_cell(j, d.Ny, d.bc[1]), sk = map_cell(k, d.Nz, d.bc[2]);
    if (si != i || sj != j || sk != k) f[vidx(d, i, j, k)] = f[vidx(d, si, sj, sk)];
}

// ------------------------------------------------------------------ CFL, diagnostics, start-up
// compute_global_vmax (Weno5_2D.jl:1942-1984: x-face averages, pressure from S) with a third direction
W5_HD void vmax_cell(const double *const q[9], const Dims &d, double gam, int i, int j, int k, double v[3])
{
    const size_t a = vidx(d, i, j, k), b = a + 1;
    const double rho = (q[0][a] + q[0][b]) / 2;
    const double ri = fast_rcp(rho);
    const double vx = (q[1][a] + q[1][b]) / 2 * ri, vy = (q[2][a] + q[2][b]) / 2 * ri, vz = (q[3][a] + q[3][b]) / 2 * ri;
    const double Bx = (q[4][a] + q[4][b]) / 2, By = (q[5][a] + q[5][b]) / 2, Bz = (q[6][a] + q[6][b]) / 2;
    const double S = (q[7][a] + q[7][b]) / 2;
    const double pres = S * pow_gm1(rho, gam);
    const double cs2 = gam * fabs(pres * ri);
    const double bbn2 = (Bx * Bx + By * By + Bz * Bz) * ri;
    const double sm = bbn2 + cs2;
    const double rx = fast_sqrt(fmax(0.0, sm * sm - 4 * (Bx * Bx * ri) * cs2));
    const double ry = fast_sqrt(fmax(0.0, sm * sm - 4 * (By * By * ri) * cs2));
    const double rz = fast_sqrt(fmax(0.0, sm * sm - 4 * (Bz * Bz * ri) * cs2));
    v[0] = fabs(vx) + fast_sqrt(0.5 * (sm + rx));
    v[1] = fabs(vy) + fast_sqrt(0.5 * (sm + ry));
    v[2] = fabs(vz) + fast_sqrt(0.5 * (sm + rz));
    if (sm != sm) { v[0] = sm; v[1] = sm; v[2] = sm; }
}

// ctl[0]=dt ctl[1]=t ctl[2..4]=vmax ctl[5]=tmax ctl[6]=steps done ctl[7]=bad flag
W5_HD void dt_from_vmax(const double vm[3], double *ctl, double dx, double dy, double dz, double cfl)
{
    const double t = ctl[1], tmax = ctl[5];
    if (tmax > 0.0 && t >= tmax) { ctl[0] = 0.0; return; }
    ctl[2] = vm[0]; ctl[3] = vm[1]; ctl[4] = vm[2];
    double dt = cfl / (vm[0] / dx + vm[1] / dy + vm[2] / dz);       // timestep(), Weno5_2D.jl:1930-1940, three terms
    if (!(dt > 0.0) || !(dt < 1e300)) { ctl[7] = 1.0; dt = 0.0; }
    if (tmax > 0.0 && t + dt > tmax) dt = tmax - t;
    ctl[0] = dt;
}

// compute_divB (Weno5_2D.jl:1986-1999) with the z term
W5_HD void divb_cell(const double *bx, const double *by, const double *bz, double *out, const Dims &d, double idx_,
                     double idy_, double idz_, int i, int j, int k)
{
    const size_t g = vidx(d, i, j, k);
    const bool in = i >= 2 && i <= d.Nx - 4 && j >= 2 && j <= d.Ny - 4 && k >= 2 && k <= d.Nz - 4;
    out[g] = in ? (bx[g] - bx[vidx(d, i - 1, j, k)]) * idx_ + (by[g] - by[vidx(d, i, j - 1, k)]) * idy_ +
                      (bz[g] - bz[vidx(d, i, j, k - 1)]) * idz_
                : 0.0;
}

// interpolateB_center2face (Weno5_2D.jl:1669-1682) with the z face field; 0 where the reference leaves it
W5_HD void center2face_cell(const double *Bx, const double *By, const double *Bz, double *bx, double *by, double *bz,
                            const Dims &d, int i, int j, int k)
{
    const size_t g = vidx(d, i, j, k);
    const bool in = i >= 1 && i <= d.Nx - 3 && j >= 1 && j <= d.Ny - 3 && k >= 1 && k <= d.Nz - 3;
    if (!in) { bx[g] = 0.0; by[g] = 0.0; bz[g] = 0.0; return; }
    bx[g] = (-Bx[vidx(d, i - 1, j, k)] + 9 * Bx[g] + 9 * Bx[vidx(d, i + 1, j, k)] - Bx[vidx(d, i + 2, j, k)]) / 16;
    by[g] = (-By[vidx(d, i, j - 1, k)] + 9 * By[g] + 9 * By[vidx(d, i, j + 1, k)] - By[vidx(d, i, j + 2, k)]) / 16;
    bz[g] = (-Bz[vidx(d, i, j, k - 1)] + 9 * Bz[g] + 9 * Bz[vidx(d, i, j, k + 1)] - Bz[vidx(d, i, j, k + 2)]) / 16;
}

// ------------------------------------------------------------------ the step, written once for device and host
// Buffers of one solver (device memory in the product, host memory in tests/hostmath/host3d.cpp)
struct Bufs3 {
    Dims d;
    double dx, dy, dz, gam, eps, cfl;
    int roundtrip;
    double *Q[4][9];           // Q[0]: step input and result; Q[1..3]: stages 1-3
    double *B[4][3];           // face fields, same roles
    double *rhs[5], *fl[6], *O[3];
    double *ctl;               // 8 doubles, see dt_from_vmax
};

// Exec provides: sweep<DIR>(SweepArgs), ct(CtArgs3), update(UpdArgs3), cell_bc(double *const v[], n, Dims),
// face_bc(double *const v[3], Dims); each returns when the work is enqueued (device) or done (host)
template <class Exec>
void enqueue_stage3(const Bufs3 &b, Exec &ex, int stage)
{
    static const double coef[5] = {0.0, 0.5, 0.5, 1.0, 1.0 / 6};
    double *const *src = b.Q[stage - 1];
    double *const *dst = b.Q[stage & 3];                 // stage 4 writes the result over the step input
    SweepArgs s;
    for (int c = 0; c < 7; ++c) s.q[c] = src[c];
    s.q[7] = src[8];
    for (int m = 0; m < 5; ++m) s.rhs[m] = b.rhs[m];
    s.dtp = b.ctl; s.d = b.d; s.gam = b.gam; s.eps = b.eps; s.g2 = (b.gam - 2.0) / (b.gam - 1.0);
    s.bsA = b.fl[0]; s.bsB = b.fl[1]; s.inv_h = 1.0 / b.dx; ex.template sweep<0>(s);
    s.bsA = b.fl[2]; s.bsB = b.fl[3]; s.inv_h = 1.0 / b.dy; ex.template sweep<1>(s);
    s.bsA = b.fl[4]; s.bsB = b.fl[5]; s.inv_h = 1.0 / b.dz; ex.template sweep<2>(s);

    CtArgs3 ct;
    for (int m = 0; m < 6; ++m) ct.fl[m] = b.fl[m];
    for (int m = 0; m < 3; ++m) {
        ct.O[m] = b.O[m];
        ct.b0[m] = b.B[0][m]; ct.b1[m] = b.B[1][m]; ct.b2[m] = b.B[2][m]; ct.b3[m] = b.B[3][m];
        ct.bn[m] = b.B[stage & 3][m];
    }
    ct.dtp = b.ctl; ct.d = b.d; ct.c = coef[stage]; ct.stage = stage;
    ct.idx_ = 1.0 / b.dx; ct.idy_ = 1.0 / b.dy; ct.idz_ = 1.0 / b.dz;
    ex.ct(ct);                                            // EMFs + face update (fused; or emf + face_update)

    UpdArgs3 u;
    for (int c = 0; c < 9; ++c) { u.q0[c] = b.Q[0][c]; u.q1[c] = b.Q[1][c]; u.q2[c] = b.Q[2][c]; u.q3[c] = b.Q[3][c]; u.qn[c] = dst[c]; }
    for (int m = 0; m < 5; ++m) u.rhs[m] = b.rhs[m];
    for (int m = 0; m < 3; ++m) u.bn[m] = b.B[stage & 3][m];
    u.dtp = b.ctl; u.d = b.d; u.c = coef[stage]; u.gam = b.gam; u.stage = stage; u.roundtrip = b.roundtrip;
    ex.update(u);                                         // writes the ghost images too: no cell_bc pass
    if (stage == 4) ex.face_bc(b.B[0], b.d);              // Weno5_2D.jl:313
}

template <class Exec>
void enqueue_step3(const Bufs3 &b, Exec &ex)
{
    for (int stage = 1; stage <= 4; ++stage) enqueue_stage3(b, ex, stage);
}

}  // namespace d3
}  // namespace w5
