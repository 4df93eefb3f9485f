// TEST INFRASTRUCTURE: runs the PRODUCT's 3-D kernel bodies and stage sequence (julia_b200/csrc/weno5_3d_body.cuh) on
// the host, thread by thread and CTA by CTA, so that tile decomposition, index maps, ghost handling and the RK
// bookkeeping of the 3-D extension are compared with the CPU oracle in a container without a GPU.  Shared memory and
// every work buffer start as NaN: a read of anything the kernels did not write shows up in the result.
// Not part of the product library.
#include "../../julia_b200/csrc/weno5_3d_body.cuh"
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

using namespace w5;
using namespace w5::d3;

namespace {

struct HostExec {
    std::vector<double> smem;
    int nseg = 0;                   // > 0: sweeps by the marching kernel body with this many segments per line (< 0: y, z only)
    int xwarp = 0;                  // > 0: x sweep by the warp-per-line body with this many segments
    int lpl = 32;                   // and this many lanes per line
    template <int DIR> void sweep(const SweepArgs &a)
    {
        if (a.dtp[0] == 0.0) return;
        int g[3];
        if (DIR == 0 && xwarp > 0) {
            if (lpl == 8) run_xwarp<8>(a); else if (lpl == 16) run_xwarp<16>(a); else run_xwarp<32>(a);
            return;
        }
        {
            const int ns = nseg > 0 ? nseg : (DIR != 0 ? -nseg : 0);
            if (ns > 0) {
                MarchGeo::grid(a.d, DIR, ns, g);
                smem.resize(MarchGeo::SMEM / sizeof(double));
                for (int bz = 0; bz < g[2]; ++bz)
                    for (int by = 0; by < g[1]; ++by)
                        for (int bx = 0; bx < g[0]; ++bx) {
                            std::fill(smem.begin(), smem.end(), std::numeric_limits<double>::quiet_NaN());
                            for (int t = 0; t < MarchGeo::NT; ++t) sweep_march<DIR>(a, ns, bx, by, bz, t, smem.data());
                        }
                return;
            }
        }
        SweepGeo<DIR>::grid(a.d, g);
        smem.resize(SweepGeo<DIR>::SMEM / sizeof(double));
        for (int bz = 0; bz < g[2]; ++bz)
            for (int by = 0; by < g[1]; ++by)
                for (int bx = 0; bx < g[0]; ++bx) {
                    std::fill(smem.begin(), smem.end(), std::numeric_limits<double>::quiet_NaN());
                    for (int t = 0; t < NT3; ++t) sweep_cells<DIR>(a, bx, by, bz, t, smem.data());
                    for (int t = 0; t < NT3; ++t) sweep_faces<DIR>(a, bx, by, bz, t, smem.data());
                    for (int t = 0; t < NT3; ++t) sweep_update<DIR>(a, bx, by, bz, t, smem.data());
                }
    }
    template <int LPL> void run_xwarp(const SweepArgs &a)
    {
        using X = XWarpGeo<LPL>;
        int g[3];
        X::grid(a.d, xwarp, g);
        smem.resize(X::SMEM / sizeof(double));
        for (int bz = 0; bz < g[2]; ++bz)
            for (int by = 0; by < g[1]; ++by)
                for (int bx = 0; bx < g[0]; ++bx) {
                    std::fill(smem.begin(), smem.end(), std::numeric_limits<double>::quiet_NaN());
                    for (int w = 0; w < X::WARPS; ++w) {
                        double *wsm = smem.data() + (size_t)w * X::PER_WARP;
                        const XWarpPos<LPL> P(a.d, xwarp, bx, by, bz, w, 0);
                        for (int ch = 0; ch < P.nchunks; ++ch) {
                            for (int l = 0; l < 32; ++l) xwarp_cells<LPL>(a, xwarp, bx, by, bz, w, l, ch, wsm);
                            for (int l = 0; l < 32; ++l) xwarp_faces<LPL>(a, xwarp, bx, by, bz, w, l, ch, wsm);
                            for (int l = 0; l < 32; ++l) xwarp_update<LPL>(a, xwarp, bx, by, bz, w, l, ch, wsm);
                        }
                    }
                }
    }
    template <class F> static void cells(const Dims &d, int lo, F f)
    {
        for (int k = lo; k < d.Nz - lo; ++k)
            for (int j = lo; j < d.Ny - lo; ++j)
                for (int i = lo; i < d.Nx - lo; ++i) f(i, j, k);
    }
    bool fused_ct = true;
    void ct(const CtArgs3 &a)
    {
        const double dt = a.dtp[0];
        if (dt == 0.0) return;
        if (fused_ct) {
            cells(a.d, 0, [&](int i, int j, int k) { ct_cell(a, dt, i, j, k); });
        } else {
            cells(a.d, 0, [&](int i, int j, int k) { emf_cell(a, i, j, k); });
            cells(a.d, 0, [&](int i, int j, int k) { face_update_cell(a, dt, i, j, k); });
        }
    }
    void update(const UpdArgs3 &a)
    {
        const double dt = a.dtp[0];
        if (dt != 0.0) cells(a.d, NB3, [&](int i, int j, int k) { update_cell(a, dt, i, j, k); });
    }
    void cell_bc(double *const v[], int n, const Dims &d)
    {
        for (int c = 0; c < n; ++c) cells(d, 0, [&](int i, int j, int k) { d3::cell_bc(v[c], d, i, j, k); });
    }
    void face_bc(double *const v[3], const Dims &d)
    {
        for (int c = 0; c < 3; ++c) cells(d, 0, [&](int i, int j, int k) { face_bc_cell(v[c], d, i, j, k); });
    }
};

struct Host3 {
    Bufs3 b;
    std::vector<double> mem;
    Host3(int nx, int ny, int nz, const double h[3], double gam, double eps, double cfl, const int bc[3], int roundtrip)
    {
        b.d.Nx = nx + 6; b.d.Ny = ny + 6; b.d.Nz = nz + 6;
        for (int m = 0; m < 3; ++m) b.d.bc[m] = bc[m];
        b.dx = h[0]; b.dy = h[1]; b.dz = h[2]; b.gam = gam; b.eps = eps; b.cfl = cfl; b.roundtrip = roundtrip;
        const size_t n = (size_t)b.d.Nx * b.d.Ny * b.d.Nz;
        mem.assign((4 * 12 + 5 + 6 + 3) * n + 8, std::numeric_limits<double>::quiet_NaN());
        double *cur = mem.data();
        auto take = [&]() { double *p = cur; cur += n; return p; };
        for (int s = 0; s < 4; ++s) { for (int k = 0; k < 9; ++k) b.Q[s][k] = take(); for (int k = 0; k < 3; ++k) b.B[s][k] = take(); }
        for (int k = 0; k < 5; ++k) b.rhs[k] = take();
        for (int k = 0; k < 6; ++k) b.fl[k] = take();
        for (int k = 0; k < 3; ++k) b.O[k] = take();
        b.ctl = cur;
        for (int k = 0; k < 8; ++k) b.ctl[k] = 0.0;
    }
    size_t n() const { return (size_t)b.d.Nx * b.d.Ny * b.d.Nz; }
};

}  // namespace

// one classical_RK4_step of the 3-D extension through the product's kernel bodies; arrays in the public layout
extern "C" int h3_rk4_step(int nx, int ny, int nz, const double h[3], double gam, double eps, const int bc[3], int roundtrip,
                           int march_nseg, const double *q0, const double *bx0, const double *by0, const double *bz0, double dt,
                           double *q4, double *bx4, double *by4, double *bz4)
{
    Host3 H(nx, ny, nz, h, gam, eps, 0.8, bc, roundtrip);
    const size_t n = H.n();
    for (int c = 0; c < 9; ++c) std::memcpy(H.b.Q[0][c], q0 + c * n, n * sizeof(double));
    const double *f[3] = {bx0, by0, bz0};
    for (int m = 0; m < 3; ++m) std::memcpy(H.b.B[0][m], f[m], n * sizeof(double));
    HostExec ex;
    // march_nseg < 0: the split EMF / face-update kernels instead of the fused one (then as for -march_nseg)
    if (march_nseg < 0) { ex.fused_ct = false; march_nseg = -march_nseg; }
    // march_nseg = 10000 * (lanes per line) + 100 * (x-warp segments) + marching segments
    ex.lpl = march_nseg / 10000 ? march_nseg / 10000 : 32;
    march_nseg %= 10000;
    ex.nseg = march_nseg % 100;
    ex.xwarp = march_nseg / 100;
    if (ex.xwarp > 0 && ex.nseg > 0) ex.nseg = -ex.nseg;
    // what weno5_upload_3d does: boundaries! on the state and the face arrays
    ex.cell_bc(H.b.Q[0], 9, H.b.d);
    ex.face_bc(H.b.B[0], H.b.d);
    H.b.ctl[0] = dt;
    enqueue_step3(H.b, ex);
    for (int c = 0; c < 9; ++c) std::memcpy(q4 + c * n, H.b.Q[0][c], n * sizeof(double));
    double *o[3] = {bx4, by4, bz4};
    for (int m = 0; m < 3; ++m) std::memcpy(o[m], H.b.B[0][m], n * sizeof(double));
    return 0;
}

extern "C" int h3_timestep(int nx, int ny, int nz, const double h[3], double gam, double cfl, const double *q, double *dt,
                           double vmax[3])
{
    Dims d;
    d.Nx = nx + 6; d.Ny = ny + 6; d.Nz = nz + 6;
    d.bc[0] = d.bc[1] = d.bc[2] = 1;
    const size_t n = (size_t)d.Nx * d.Ny * d.Nz;
    const double *qv[9];
    for (int c = 0; c < 9; ++c) qv[c] = q + c * n;
    double vm[3] = {0.0, 0.0, 0.0};
    for (int k = NB3; k < d.Nz - NB3; ++k)
        for (int j = NB3; j < d.Ny - NB3; ++j)
            for (int i = NB3; i < d.Nx - NB3; ++i) {
                double v[3];
                vmax_cell(qv, d, gam, i, j, k, v);
                for (int m = 0; m < 3; ++m) vm[m] = std::fmax(vm[m], v[m]);
            }
    double ctl[8] = {0};
    dt_from_vmax(vm, ctl, h[0], h[1], h[2], cfl);
    *dt = ctl[0];
    for (int m = 0; m < 3; ++m) vmax[m] = ctl[2 + m];
    return 0;
}

extern "C" int h3_divb_center2face(int nx, int ny, int nz, const double h[3], const double *q, double *bx, double *by,
                                   double *bz, double *divb)
{
    Dims d;
    d.Nx = nx + 6; d.Ny = ny + 6; d.Nz = nz + 6;
    d.bc[0] = d.bc[1] = d.bc[2] = 1;
    const size_t n = (size_t)d.Nx * d.Ny * d.Nz;
    HostExec::cells(d, 0, [&](int i, int j, int k) { center2face_cell(q + 4 * n, q + 5 * n, q + 6 * n, bx, by, bz, d, i, j, k); });
    HostExec::cells(d, 0, [&](int i, int j, int k) { divb_cell(bx, by, bz, divb, d, 1.0 / h[0], 1.0 / h[1], 1.0 / h[2], i, j, k); });
    return 0;
}

// the segment count the product would choose (MarchGeo::segments) for a grid on a GPU with `sms` SMs
extern "C" int h3_march_segments(int nx, int ny, int nz, int dir, int sms)
{
    Dims d;
    d.Nx = nx + 6; d.Ny = ny + 6; d.Nz = nz + 6;
    d.bc[0] = d.bc[1] = d.bc[2] = 1;
    return MarchGeo::segments(d, dir, sms);
}

extern "C" int h3_xwarp_plan(int nx, int ny, int nz, int bcx, int sms, int *nseg, int *lpl)
{
    Dims d;
    d.Nx = nx + 6; d.Ny = ny + 6; d.Nz = nz + 6;
    d.bc[0] = bcx; d.bc[1] = d.bc[2] = 1;
    xwarp_plan(d, sms, *nseg, *lpl);
    return 0;
}
