// The 1-D solver (Weno5_1D.jl:33-236, energy branch) as ONE resident kernel on a thread-block cluster.
//
// Config 1 of BASELINE.json is a 1024-cell shock tube: 1030 cells x 9 doubles = 74 KB of state.  Driven through the
// 2-D machinery it costs ~15 launches of 4-8 us per RK4 step for a few microseconds of arithmetic each.  Here a
// cluster of 8 CTAs keeps the whole line in (distributed) shared memory for the entire run:
//   * CTA r owns a contiguous chunk of the line plus 3 halo cells per side; the RK base state, the running RK
//     combination and both ping-pong copies of the current state never leave shared memory;
//   * per stage: cell phase (cons->prim, wave speeds, flux; Weno5_1D.jl:1021-1036, :407-437, :1005-1019), face phase
//     (the same face_coefficients / face_flux / back_project as the 2-D sweeps; :784-1003, :287-405), update phase
//     (flux difference + RK combination, :152-205), then ONE cluster barrier, after which every CTA pulls its 3+3
//     halo cells of the new state straight out of its neighbours' shared memory (DSMEM) -- or applies the boundary
//     condition at the ends of the line (boundaries!, :1076-1090);
//   * per step: S = E2S(q) (:261-272), the CFL reduction compute_global_vmax (:1038-1074) over the cluster (every
//     CTA stores its block maximum into every CTA's slot array before the stage-4 barrier, so the reduction costs no
//     barrier of its own) and dt = courFac dx / vmax (:55), the clock and the tmax logic of weno5_advance;
//   * the step loop runs inside the kernel: one launch per weno5_advance / weno5_rk4_step call.
// Same arithmetic (same device functions, same update formulas) as the multi-kernel path, which remains for lines
// that do not fit (more than 8 x 255 cells), for the dual-energy modes and as the reference in the tests.
#include "weno5_kernels.cuh"
#include "weno5_math.cuh"
#include <cooperative_groups.h>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace w5 {

constexpr int L1_CLUSTER = 8;          // CTAs per cluster (portable maximum)
constexpr int L1_HALO = 3;             // stencil cells f-2 .. f+3 of the faces c0-1 .. c1-1
constexpr int L1_MAXCHUNK = 255;       // one face per thread, chunk + 1 faces, <= 256 threads

struct Line1dArgs {
    double *q[NCELL];        // planes of the resident state (Q0), internal indexing, G ghost cells per side
    double *ctl;             // dt, t, vxmax, vymax, tmax, steps, bad
    int N;                   // cells incl. ghosts
    int chunk;               // cells per CTA
    int side_lo, side_hi;    // SIDE_OUTFLOW / SIDE_PERIODIC
    int nsteps;
    double fixed_dt;         // > 0: classical_RK4_step(q, dt) with this dt, once; else the driver loop with CFL steps
    double dx, gam, g2, eps, cfl;
};

// shared-memory layout (doubles), W = chunk + 2 HALO cells, index 0 <-> cell c0 - HALO
//   qs  [2][8][W]   current state, ping-pong: rho,Mx,My,Mz,Bx,By,Bz,E
//   cb  [21][W]     cell quantities (CellQ order of weno5_math.cuh)
//   fh  [7][chunk+1] numerical fluxes of the faces c0-1 .. c1-1
//   q0  [7][chunk]  RK base: rho,Mx,My,Mz,Bz,E,By      acc [7][chunk]  -q0 + q1 + 2 q2
//   Sv  [W]         entropy of the current state (CFL reduction)
//   red [CLUSTER]   block maxima of the cluster (bit patterns)
__host__ __device__ inline size_t line1d_smem_doubles(int chunk)
{
    const size_t W = chunk + 2 * L1_HALO;
    return 2 * 8 * W + 21 * W + 7 * (size_t)(chunk + 1) + 14 * (size_t)chunk + W + L1_CLUSTER + 8;
}

__global__ void __launch_bounds__(256, 1) line1d_kernel(const Line1dArgs a)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int t = threadIdx.x, NT = blockDim.x;
    const int chunk = a.chunk, W = chunk + 2 * L1_HALO;
    extern __shared__ __align__(16) double sm[];
    double *qs = sm;                                  // [2][8][W]
    double *cb = qs + 2 * 8 * W;                      // [21][W]
    double *fh = cb + 21 * W;                         // [7][chunk+1]
    double *q0 = fh + 7 * (chunk + 1);                // [7][chunk]
    double *acc = q0 + 7 * chunk;                     // [7][chunk]
    double *Sv = acc + 7 * chunk;                     // [W]
    unsigned long long *red = reinterpret_cast<unsigned long long *>(Sv + W);   // [CLUSTER]
    unsigned long long *wmax = red + L1_CLUSTER;                                // [8] per-warp maxima

    const int c0 = G + rank * chunk;                  // first owned cell
    const int c1 = min(c0 + chunk, a.N - G);          // one past the last owned cell (the last CTA may own fewer)
    const int nown = max(c1 - c0, 0);
    const int n_int = a.N - 2 * G;                    // interior cells of the line
    // the physical state planes in the order of qs: rho,Mx,My,Mz,Bx,By,Bz,E
    const int plane8[8] = {C_RHO, C_MX, C_MY, C_MZ, C_BX, C_BY, C_BZ, C_E};
    // the seven flux-advanced components in the order of q0/acc/rhs: rho,Mx,My,Mz,Bz,E,By -> index into qs / fhat
    const int adv_q[7] = {0, 1, 2, 3, 6, 7, 5};
    const int adv_f[7] = {0, 1, 2, 3, 5, 6, 4};

    // global cell index behind local index w of CTA `rank`, through the boundary condition (boundaries!, :1076)
    auto src_cell = [&](int cell) {
        if (cell < G) return a.side_lo == SIDE_PERIODIC ? cell + n_int : G;
        if (cell >= a.N - G) return a.side_hi == SIDE_PERIODIC ? cell - n_int : a.N - G - 1;
        return cell;
    };

    // ---- load: own cells and halos of the start state (ghosts through the boundary condition, like bc_fill)
    for (int e = t; e < 8 * W; e += NT) {
        const int k = e / W, w = e % W;
        const int cell = src_cell(min(max(c0 - L1_HALO + w, 0), a.N - 1));
        qs[k * W + w] = a.q[plane8[k]][cell];
    }
    for (int w = t; w < W; w += NT) Sv[w] = a.q[C_S][src_cell(min(max(c0 - L1_HALO + w, 0), a.N - 1))];
    if (t < L1_CLUSTER) red[t] = 0ull;
    __syncthreads();
    cluster.sync();

    double tnow = a.ctl[1];
    const double tmax = a.ctl[4];
    double dt = a.fixed_dt, vmax = 0.0;
    int done = 0, bad = 0;
    int cur = 0;                                      // ping-pong index of the current state

    // block maximum of the CFL speeds of the owned cells -> every CTA's slot array (Weno5_1D.jl:1038-1074)
    auto publish_vmax = [&]() {
        double v = 0.0;
        if (t < nown) {
            const double *q = qs + cur * 8 * W;
            const int w = L1_HALO + t;
            const double rho = (q[0 * W + w] + q[0 * W + w + 1]) / 2;
            const double ri = fast_rcp(rho);
            const double vx = (q[1 * W + w] + q[1 * W + w + 1]) / 2 * ri;
            const double Bx = (q[4 * W + w] + q[4 * W + w + 1]) / 2;
            const double By = (q[5 * W + w] + q[5 * W + w + 1]) / 2;
            const double Bz = (q[6 * W + w] + q[6 * W + w + 1]) / 2;
            const double S = (Sv[w] + Sv[w + 1]) / 2;
            const double pres = S * pow_gm1(rho, a.gam);
            const double cs2 = a.gam * fabs(pres * ri);
            const double B2 = Bx * Bx + By * By + Bz * Bz;
            const double bbn2 = B2 * ri, bnx2 = Bx * Bx * ri;
            const double smm = bbn2 + cs2;
            const double rootx = fast_sqrt(fmax(0.0, smm * smm - 4 * bnx2 * cs2));
            v = fabs(vx) + fast_sqrt(0.5 * (smm + rootx));
            if (smm != smm) v = smm;
        }
        unsigned long long b = (unsigned long long)__double_as_longlong(v);
        if (v != v || v < 0.0) b = ~0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
        if ((t & 31) == 0) wmax[t >> 5] = b;
        __syncthreads();
        if (t < L1_CLUSTER) {                          // thread r stores this block's maximum into CTA r's slot
            unsigned long long m = 0ull;
            for (int w2 = 0; w2 < (NT + 31) / 32; ++w2) m = max(m, wmax[w2]);
            unsigned long long *remote = cluster.map_shared_rank(red, t);
            remote[rank] = m;
        }
    };
    // pull the halo cells of the current state (and of S) out of the neighbours' shared memory / the boundary
    auto pull_halos = [&]() {
        for (int e = t; e < 9 * 2 * L1_HALO; e += NT) {
            const int k = e / (2 * L1_HALO), h = e % (2 * L1_HALO);
            const int w = h < L1_HALO ? h : nown + h;                      // local halo index (right halo follows the owned cells)
            const int cell = src_cell(c0 - L1_HALO + w);                   // interior cell that holds the value
            const int owner = min((cell - G) / chunk, L1_CLUSTER - 1);
            const int wl = L1_HALO + cell - (G + owner * chunk);
            const double *base = k < 8 ? qs + cur * 8 * W + k * W : Sv;
            const double *remote = cluster.map_shared_rank(base, owner);
            const double v = remote[wl];
            if (k < 8) qs[cur * 8 * W + k * W + w] = v; else Sv[w] = v;
        }
    };

    if (!(a.fixed_dt > 0.0)) publish_vmax();
    cluster.sync();

    for (int step = 0; step < a.nsteps; ++step) {
        if (!(a.fixed_dt > 0.0)) {
            unsigned long long m = 0ull;
            for (int r = 0; r < L1_CLUSTER; ++r) m = max(m, red[r]);
            vmax = __longlong_as_double((long long)m);
            if (tmax > 0.0 && tnow >= tmax) break;
            dt = a.cfl * a.dx / vmax;
            if (!(dt > 0.0) || !(dt < 1e300)) { bad = 1; break; }
            if (tmax > 0.0 && tnow + dt > tmax) dt = tmax - tnow;
        }
        // RK base = state at the start of the step
        for (int e = t; e < 7 * nown; e += NT) {
            const int m = e / nown, i = e % nown;
            q0[m * chunk + i] = qs[cur * 8 * W + adv_q[m] * W + L1_HALO + i];
        }
        for (int stage = 1; stage <= 4; ++stage) {
            const double ck = stage <= 2 ? 0.5 : (stage == 3 ? 1.0 : 1.0 / 6.0);
            const double cx = ck / a.dx * dt;
            const double *q = qs + cur * 8 * W;
            double *qn = qs + (cur ^ 1) * 8 * W;
            // ---- cell phase
            for (int w = t; w < nown + 2 * L1_HALO; w += NT) {
                double qq[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) qq[k] = q[k * W + w];
                CellOut o;
                cell_quantities(qq, a.gam, o);
#pragma unroll
                for (int m = 0; m < 7; ++m) cb[(CQ_F0 + m) * W + w] = o.F[m];
                cb[(CQ_Q0 + 0) * W + w] = qq[0]; cb[(CQ_Q0 + 1) * W + w] = qq[1]; cb[(CQ_Q0 + 2) * W + w] = qq[2];
                cb[(CQ_Q0 + 3) * W + w] = qq[3]; cb[(CQ_Q0 + 4) * W + w] = qq[5]; cb[(CQ_Q0 + 5) * W + w] = qq[6];
                cb[(CQ_Q0 + 6) * W + w] = qq[7];
                cb[CQ_V1 * W + w] = o.v1; cb[CQ_V2 * W + w] = o.v2; cb[CQ_V3 * W + w] = o.v3;
                cb[CQ_B1 * W + w] = qq[4];
                cb[CQ_LF * W + w] = o.lf; cb[CQ_LA * W + w] = o.la; cb[CQ_LS * W + w] = o.ls;
            }
            __syncthreads();
            // ---- face phase: face c0 - 1 + t between the local cells t + 2 and t + 3
            if (t < nown + 1) {
                const int iL = t + 2, iR = t + 3;
                FaceCoef fc;
                face_coefficients(cb[CQ_Q0 * W + iL], cb[CQ_Q0 * W + iR], cb[CQ_V1 * W + iL], cb[CQ_V1 * W + iR],
                                  cb[CQ_V2 * W + iL], cb[CQ_V2 * W + iR], cb[CQ_V3 * W + iL], cb[CQ_V3 * W + iR],
                                  cb[CQ_B1 * W + iL], cb[CQ_B1 * W + iR], cb[(CQ_Q0 + 4) * W + iL], cb[(CQ_Q0 + 4) * W + iR],
                                  cb[(CQ_Q0 + 5) * W + iL], cb[(CQ_Q0 + 5) * W + iR], cb[(CQ_Q0 + 6) * W + iL],
                                  cb[(CQ_Q0 + 6) * W + iR], a.gam, a.g2, fc);
                double amax[7];
                face_amax(cb[CQ_V1 * W + iL], cb[CQ_LF * W + iL], cb[CQ_LA * W + iL], cb[CQ_LS * W + iL],
                          cb[CQ_V1 * W + iR], cb[CQ_LF * W + iR], cb[CQ_LA * W + iR], cb[CQ_LS * W + iR], amax);
                auto load = [&](int k, double F[7], double x[7]) {
                    const int ii = t + k;
#pragma unroll
                    for (int m = 0; m < 7; ++m) {
                        F[m] = cb[(CQ_F0 + m) * W + ii];
                        x[m] = cb[(CQ_Q0 + m) * W + ii];
                    }
                };
                double Fs[7], fhat[7];
                RegStash st;
                BackCoef bc;
                split_back(fc, bc);
                face_flux(fc, amax, a.eps, load, st, Fs);
                back_project(bc, Fs, fhat);
#pragma unroll
                for (int m = 0; m < 7; ++m) fh[m * (chunk + 1) + t] = fhat[m];
            }
            __syncthreads();
            // ---- update phase (Weno5_1D.jl:152-205; same formulas as update_1d_kernel)
            if (t < nown) {
#pragma unroll
                for (int m = 0; m < 7; ++m) {
                    const double rhs = fh[adv_f[m] * (chunk + 1) + t + 1] - fh[adv_f[m] * (chunk + 1) + t];
                    const double qcen = q[adv_q[m] * W + L1_HALO + t];
                    double b;
                    if (stage <= 3) {
                        b = q0[m * chunk + t];
                        if (stage == 2) acc[m * chunk + t] = qcen - b;
                        if (stage == 3) acc[m * chunk + t] = fma(2.0, qcen, acc[m * chunk + t]);
                    } else {
                        b = (acc[m * chunk + t] + qcen) * (1.0 / 3.0);
                    }
                    qn[adv_q[m] * W + L1_HALO + t] = fma(-cx, rhs, b);
                }
                qn[4 * W + L1_HALO + t] = q[4 * W + L1_HALO + t];            // Bx is constant (:771)
            }
            cur ^= 1;
            if (stage == 4) {
                // S = E2S(q) of the new state (:261), then this block's CFL maximum into every CTA's slots
                __syncthreads();
                if (t < nown) {
                    const double *qc = qs + cur * 8 * W;
                    const int w = L1_HALO + t;
                    double S, E2;
                    e2s_s2e(qc[0 * W + w], qc[1 * W + w], qc[2 * W + w], qc[3 * W + w], qc[4 * W + w], qc[5 * W + w],
                            qc[6 * W + w], qc[7 * W + w], a.gam, S, E2);
                    Sv[w] = S;
                }
            }
            __syncthreads();
            cluster.sync();                            // every CTA's new state (and S) is complete
            pull_halos();
            __syncthreads();
            if (stage == 4 && !(a.fixed_dt > 0.0)) {
                publish_vmax();                        // needs the right-hand halo cell of the new state
                cluster.sync();
            }
        }
        tnow += dt;
        ++done;
        if (a.fixed_dt > 0.0) break;
    }

    // ---- write back: owned cells, and the ghost cells next to the ends of the line (boundaries!, :1076)
    cluster.sync();
    {
        const double *q = qs + cur * 8 * W;
        for (int e = t; e < 9 * nown; e += NT) {
            const int k = e / nown, i = e % nown;
            if (k < 8) a.q[plane8[k]][c0 + i] = q[k * W + L1_HALO + i];
            else a.q[C_S][c0 + i] = Sv[L1_HALO + i];
        }
        // ghost cells: CTA that owns the source cell writes them (the source is an owned cell of exactly one CTA)
        for (int e = t; e < 9 * 2 * G; e += NT) {
            const int k = e / (2 * G), gidx = e % (2 * G);
            const int cell = gidx < G ? gidx : a.N - 2 * G + gidx;
            const int src = src_cell(cell);
            if (src >= c0 && src < c1) {
                const double v = k < 8 ? q[k * W + L1_HALO + src - c0] : Sv[L1_HALO + src - c0];
                a.q[k < 8 ? plane8[k] : C_S][cell] = v;
            }
        }
    }
    if (rank == 0 && t == 0 && !(a.fixed_dt > 0.0)) {   // the driver loop's clock (classical_RK4_step alone has none)
        a.ctl[0] = dt;
        a.ctl[1] = tnow;
        a.ctl[2] = vmax;
        a.ctl[5] += (double)done;
        if (bad) a.ctl[6] = 1.0;
    }
    cluster.sync();                                    // no CTA exits while its shared memory may still be read
}

// cells per CTA for a line of N cells (incl. ghosts), or 0 if the line does not fit the cluster kernel
int line1d_chunk(int N)
{
    if (getenv("WENO5_NO_LINE1D")) return 0;          // development knob: the multi-kernel path (read per call: tests toggle it)
    const int n_int = N - 2 * G;
    const int chunk = (n_int + L1_CLUSTER - 1) / L1_CLUSTER;
    if (chunk < 2 * L1_HALO || chunk > L1_MAXCHUNK) return 0;           // halos must come from the direct neighbours
    if ((L1_CLUSTER - 1) * chunk >= n_int) return 0;                    // every CTA owns at least one cell
    return chunk;
}

cudaError_t launch_line1d(double *const q9[NCELL], double *ctl, const Geom &g, const int side[4], int nsteps, double fixed_dt,
                          double dx, double gam, double eps, double cfl, cudaStream_t s)
{
    Line1dArgs a;
    for (int k = 0; k < NCELL; ++k) a.q[k] = q9[k];
    a.ctl = ctl;
    a.N = g.NXI;
    a.chunk = line1d_chunk(g.NXI);
    a.side_lo = side[0]; a.side_hi = side[1];
    a.nsteps = nsteps; a.fixed_dt = fixed_dt;
    a.dx = dx; a.gam = gam; a.g2 = (gam - 2.0) / (gam - 1.0); a.eps = eps; a.cfl = cfl;
    const size_t smem = line1d_smem_doubles(a.chunk) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(line1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(L1_CLUSTER, 1, 1);
    cfg.blockDim = dim3((a.chunk + 1 + 31) / 32 * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = L1_CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, line1d_kernel, a);
}

}  // namespace w5
