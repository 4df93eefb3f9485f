// Hand-written sm_100a kernels of libweno5mhd.
//
// flux_kernel<DIR>  the fused per-face kernel: cons->prim, wave speeds, physical flux (cell
//                   phase, once per cell, kept in shared memory), then per face the structured
//                   characteristic projection, WENO5 and back-projection (weno5_math.cuh),
//                   then the flux difference.  DIR 0 sweeps x and writes d fhat + fsy;
//                   DIR 1 sweeps y IN PLACE (threads stay along x, stencil along j -- no
//                   transposed copy like the reference's rotate_state, Weno5_2D.jl:1897-1927)
//                   and is fused with the RK stage update of the six flux-advanced fields.
//                   A CTA marches along the sweep direction; cell data of the 5 overlapping
//                   stencil cells is carried in shared memory, so nothing is recomputed.
// ct_faces/ct_centers  corner EMF, face-field RK update, 4th-order face->centre interpolation
//                   (Weno5_2D.jl:319-356 and the bxb/byb lines of :125-317).
// bc_fill/face_bc   boundaries! (Weno5_2D.jl:1684-1732) + periodic addition.
// entropy/vmax/dt   E2S/S2E (:1770-1793), compute_global_vmax/timestep (:1930-1984).
#include "weno5_kernels.cuh"
#include "weno5_math.cuh"
#include <cstdlib>

#ifndef W5_DEFAULT_VARIANT
#define W5_DEFAULT_VARIANT 5
#endif

namespace w5 {

// A CTA of NT threads handles a chunk of NT faces per march step: DIR 0 (x sweep) NT/64 rows of
// 64 faces, DIR 1 (y sweep) NT/YW face rows of YW columns -- lanes always run along x so that
// global accesses are coalesced in both sweeps.  Cell data lives in a ring of RS = NS + 5 slots
// along the sweep (the 5 trailing stencil cells of a chunk stay where they are for the next one).
template <int DIR, int NT, int YW, int NCQ = CQ_COUNT> struct Tile {
    static constexpr int NS = DIR == 0 ? 64 : NT / YW;     // faces along the sweep per chunk
    static constexpr int NPB = DIR == 0 ? NT / 64 : YW;    // lines across the sweep per CTA
    static constexpr int RS = (NS + 5 + 15) / 16 * 16;     // ring slots along the sweep (>= NS + 5; a multiple of
                                                           // 16 keeps the wrapped x-sweep accesses bank-conflict free)
    static constexpr int CBS = RS * NPB;                   // cell-buffer entries per quantity
    static constexpr int FBS = (NS + 1) * NPB;             // face-flux-buffer entries per component
    // staging: [field][NT], thread t's element at t.  A TMA box (x: NS cells x NPB rows, y: NPB
    // columns x NS rows) lands in exactly this layout because t = row-major index inside the box.
    // The start address of a TMA box must be 16-byte aligned, i.e. an even cell index: the x-sweep
    // box is 2 cells wider and starts at c0 - (c0 & 1); threads read at offset +(c0 & 1).
    static constexpr int BOX_X = DIR == 0 ? NS + 2 : NPB, BOX_Y = DIR == 0 ? NPB : NS;
    static constexpr int FST = BOX_X * BOX_Y;              // staging stride per field (one multi-plane TMA box is contiguous)
    static constexpr int OFFQ = (NCQ * CBS + 7 * FBS + 15) / 16 * 16;        // 128-byte aligned (TMA destination)
    static constexpr int STQ = 8 * FST;                    // staged raw state of the next chunk
    static constexpr int STU = DIR == 0 ? 0 : 18 * FST;    // staged RK operands (rhs, base, acc)
    static constexpr int SMEM = (OFFQ + (STQ + 15) / 16 * 16 + STU) * (int)sizeof(double) + 16;
    // staging index of the thread at (sweep position s, line p); e = even-start shift (x sweep, TMA)
    __device__ static __forceinline__ int sti(int s, int p, int e) { return DIR == 0 ? p * BOX_X + s + e : s * NPB + p; }
    __device__ static __forceinline__ int cbi(int slot, int p) { return DIR == 0 ? p * RS + slot : slot * NPB + p; }
    __device__ static __forceinline__ int fbi(int slot, int p) { return DIR == 0 ? p * (NS + 1) + slot : slot * NPB + p; }
};

// Ampere-style asynchronous copy global -> shared (LDGSTS): no register staging, so the loads of
// the NEXT march step and of the RK operands fly while the FP64 pipe works on the current faces.
__device__ __forceinline__ void cp_async8(double *smem, const double *gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// TMA bulk copies (cp.async.bulk, SASS UBLKCP) with mbarrier transaction counting: one warp issues
// whole row segments global -> shared; the consumers wait on the barrier's phase parity.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// one box of the 3-D tensor map (x, y, plane) -> shared memory; out-of-bounds elements are zero-filled
__device__ __forceinline__ void tma_load_box(double *dst, const CUtensorMap *tm, int x, int y, int z, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst)),
                 "l"(tm), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    }
}

// global index of sweep cell c on line `line` (clamped: cells past the array are never used by an
// owned face, they only have to be finite)
template <int DIR>
__device__ __forceinline__ size_t cell_index(const FluxArgs &a, int c, int line)
{
    const int cc = min(c, a.N - 1), ll = min(line, a.NP - 1);
    return DIR == 0 ? (size_t)cc + (size_t)a.pitch * ll : (size_t)ll + (size_t)a.pitch * cc;
}

// cons->prim, eigenvalues, flux of one cell -> shared memory entry i.  q is in the sweep frame.
// BR 1 (entropy branch): q[7] = S; two more ring entries, P(S) and rho^(1-gam), feed the face average
constexpr int CQ_P = CQ_COUNT, CQ_RG0 = CQ_COUNT + 1, CQ_COUNT_S = CQ_COUNT + 2;

template <int CBS, int BR = 0>
__device__ __forceinline__ void cell_store(const FluxArgs &a, double *cb, const double q[8], int i)
{
    CellOut o;
    if constexpr (BR == 0) {
        cell_quantities(q, a.gam, o);
    } else {
        CellOutS os;
        cell_quantities_S(q, a.gam, o, os);
        cb[CQ_P * CBS + i] = os.P;
        cb[CQ_RG0 * CBS + i] = os.rg0;
    }
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

// sweep-frame component k reads physical plane perm<DIR>(k): (x,y,z) -> (y,z,x) for the y sweep
template <int DIR> __device__ __forceinline__ constexpr int perm(int k)
{
    constexpr int px[8] = {0, 1, 2, 3, 4, 5, 6, 7}, py[8] = {0, 2, 3, 1, 5, 6, 4, 7};
    return DIR ? py[k] : px[k];
}

// The tensor maps of one sweep direction: the same 3-D view (x, y, plane) of the plane pool with
// boxes of 8, 6, 4 and 2 planes (the box shape is part of a tensor map).
struct FluxMaps { CUtensorMap m8, m6, m4, m2; };

// TMA path, executed by one thread: stage the raw state of the chunk whose first new cell is c0.
// The 8 state planes are consecutive in the pool, so ONE box carries all of them.
template <int DIR, int NT, int YW>
__device__ __forceinline__ void tma_issue_state(const FluxArgs &a, const FluxMaps *tm, double *stq, unsigned long long *bar, int c0,
                                                int strip = (int)blockIdx.x)
{
    using T = Tile<DIR, NT, YW>;
    mbar_expect_tx(bar, 8 * T::FST * (unsigned)sizeof(double));
    const int bx = DIR == 0 ? c0 - (c0 & 1) : strip * T::NPB;
    const int by = DIR == 0 ? strip * T::NPB : c0;
    tma_load_box(stq, &tm->m8, bx, by, a.qc_plane[0], bar);
}

// TMA path: stage the RK operands of the cells s0 .. s0+NS-1 of the y sweep: rhs (6 consecutive
// planes), base (rho,Mx,My,Mz | Bz,E: planes 0-3 and 6-7 of the base state), acc (6 consecutive)
template <int NT, int YW>
__device__ __forceinline__ void tma_issue_operands(const FluxArgs &a, const FluxMaps *tm, double *stu, unsigned long long *bar, int s0)
{
    using T = Tile<1, NT, YW>;
    const int nf = a.stage == 3 ? 18 : 12;
    mbar_expect_tx(bar, nf * T::FST * (unsigned)sizeof(double));
    const int bx = (int)blockIdx.x * T::NPB;
    tma_load_box(stu, &tm->m6, bx, s0, a.rhs_plane[0], bar);
    if (a.stage <= 3) {
        tma_load_box(stu + 6 * T::FST, &tm->m4, bx, s0, a.base_plane[0], bar);
        tma_load_box(stu + 10 * T::FST, &tm->m2, bx, s0, a.base_plane[4], bar);
    }
    if (a.stage >= 3) tma_load_box(stu + 12 * T::FST, &tm->m6, bx, s0, a.acc_plane[0], bar);
}

template <int DIR, int NT, int YW, int MINB, bool TMA, int BR = 0>
__global__ void __launch_bounds__(NT, MINB) flux_kernel(const FluxArgs a, const __grid_constant__ FluxMaps tmap)
{
    constexpr int NCQ = BR == 0 ? CQ_COUNT : CQ_COUNT_S;
    using T = Tile<DIR, NT, YW, NCQ>;
    constexpr int NS = T::NS, NPB = T::NPB, RS = T::RS;
    constexpr int CBS = T::CBS, FBS = T::FBS;
    extern __shared__ __align__(128) double sm[];
    double *cb = sm;
    double *fb = sm + NCQ * CBS;
    double *stq = sm + T::OFFQ;                  // [8][NT]   raw state of the next chunk
    double *stu = stq + (T::STQ + 15) / 16 * 16; // [18][NT]  rhs, base, acc of this chunk (y sweep)
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(stu + T::STU);   // [0] state, [1] operands

    const double dt = *a.dtp;
    if (dt == 0.0) return;                       // advance() past tmax: leave the state untouched

    const int t = threadIdx.x;
    const int s = DIR == 0 ? (t % NS) : (t / NPB);
    const int p = DIR == 0 ? (t / NS) : (t % NPB);
    const int line = blockIdx.x * NPB + p;
    const int seg = blockIdx.y;

    // this CTA owns faces [fA, fB) of its lines; it starts one face early (seg > 0) so that the
    // first owned cell has both of its faces
    const int fA = 2 + seg * a.seg_len;
    const int fB = min(fA + a.seg_len, a.N - 3);           // last face N-4: its stencil ends at cell N-1
    const int fStart = seg == 0 ? fA : fA - 1;

    const bool line_ok = a.is1d ? (line == 0) : (line >= G && line < a.NP - G);

    unsigned phq = 0, phu = 0;                   // mbarrier phase parities
    if constexpr (TMA) {
        if (t == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
        __syncthreads();
    }

    // stage the first chunk asynchronously, meanwhile compute the 5 leading stencil cells directly.
    // Ring slot of sweep cell c is (c - (fStart - 2)) mod RS.
    if constexpr (TMA) {
        if (t == 0) tma_issue_state<DIR, NT, YW>(a, &tmap, stq, &bars[0], fStart + 3);
    } else {
        const size_t g0 = cell_index<DIR>(a, fStart + 3 + s, line);
#pragma unroll
        for (int k = 0; k < 8; ++k) cp_async8(stq + k * T::FST + T::sti(s, p, 0), a.qc[k] + g0);
        cp_async_commit();
    }
    for (int e = t; e < 5 * NPB; e += NT) {
        const int sl = DIR == 0 ? e % 5 : e / NPB;
        const int pp = DIR == 0 ? e / 5 : e % NPB;
        const size_t g = cell_index<DIR>(a, fStart - 2 + sl, blockIdx.x * NPB + pp);
        double q[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) q[k] = __ldg(a.qc[perm<DIR>(k)] + g);
        cell_store<CBS, BR>(a, cb, q, T::cbi(sl, pp));
    }

    int r0 = 0;                                  // ring slot of cell s0 - 2
    for (int s0 = fStart; s0 < fB; s0 += NS) {
        // ring slots of the six stencil cells f-2 .. f+3 of this thread's face f = s0 + s
        int sl[6];
        {
            int r = r0 + s;
            if (r >= RS) r -= RS;
#pragma unroll
            for (int k = 0; k < 6; ++k) { sl[k] = r; r = (r + 1 == RS) ? 0 : r + 1; }
        }
        // ---------------- cell phase: cell s0 + 3 + s from the staged state (each thread reads back
        // its own copies, no barrier needed for the staging buffer)
        if constexpr (TMA) { mbar_wait(&bars[0], phq); phq ^= 1u; } else cp_async_wait<0>();
        {
            const int si = T::sti(s, p, TMA ? ((s0 + 3) & 1) : 0);
            double q[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) q[k] = stq[perm<DIR>(k) * T::FST + si];     // staged in physical plane order
            cell_store<CBS, BR>(a, cb, q, T::cbi(sl[5], p));
        }
        __syncthreads();

        // ---------------- asynchronous copies that fly under the face phase:
        // group A = this chunk's RK operands (y sweep), group B = the next chunk's state
        const bool upd_ok = (s0 + s > fStart) && (s0 + s < fB) && (s0 + s >= G) && (s0 + s < a.N - G) && line_ok;
        if constexpr (TMA) {
            if (t == 0) {
                if constexpr (DIR == 1) tma_issue_operands<NT, YW>(a, &tmap, stu, &bars[1], s0);
                if (s0 + NS < fB) tma_issue_state<DIR, NT, YW>(a, &tmap, stq, &bars[0], s0 + NS + 3);
            }
        } else {
            if constexpr (DIR == 1) {
                if (upd_ok) {
                    const size_t gu = (size_t)line + (size_t)a.pitch * (s0 + s);
#pragma unroll
                    for (int m = 0; m < 6; ++m) {
                        cp_async8(stu + m * T::FST + t, a.rhs[m] + gu);
                        if (a.stage <= 3) cp_async8(stu + (6 + m) * T::FST + t, a.base[m] + gu);
                        if (a.stage >= 3) cp_async8(stu + (12 + m) * T::FST + t, a.acc[m] + gu);
                    }
                }
            }
            cp_async_commit();
            if (s0 + NS < fB) {
                const size_t gn = cell_index<DIR>(a, s0 + NS + 3 + s, line);
#pragma unroll
                for (int k = 0; k < 8; ++k) cp_async8(stq + k * T::FST + T::sti(s, p, 0), a.qc[k] + gn);
            }
            cp_async_commit();
        }

        // ---------------- face phase: face f between cells f (stencil slot 2) and f+1 (slot 3)
        const int f = s0 + s;
        if (f < fB) {
            const int iL = T::cbi(sl[2], p), iR = T::cbi(sl[3], p);
            FaceCoef fc;
            if constexpr (BR == 0)
                face_coefficients(cb[CQ_Q0 * CBS + iL], cb[CQ_Q0 * CBS + iR], cb[CQ_V1 * CBS + iL], cb[CQ_V1 * CBS + iR],
                                  cb[CQ_V2 * CBS + iL], cb[CQ_V2 * CBS + iR], cb[CQ_V3 * CBS + iL], cb[CQ_V3 * CBS + iR],
                                  cb[CQ_B1 * CBS + iL], cb[CQ_B1 * CBS + iR], cb[(CQ_Q0 + 4) * CBS + iL],
                                  cb[(CQ_Q0 + 4) * CBS + iR], cb[(CQ_Q0 + 5) * CBS + iL], cb[(CQ_Q0 + 5) * CBS + iR],
                                  cb[(CQ_Q0 + 6) * CBS + iL], cb[(CQ_Q0 + 6) * CBS + iR], a.gam, a.g2, fc);
            else
                face_coefficients_S(cb[CQ_Q0 * CBS + iL], cb[CQ_Q0 * CBS + iR], cb[CQ_RG0 * CBS + iL], cb[CQ_RG0 * CBS + iR],
                                    cb[CQ_V1 * CBS + iL], cb[CQ_V1 * CBS + iR], cb[CQ_V2 * CBS + iL], cb[CQ_V2 * CBS + iR],
                                    cb[CQ_V3 * CBS + iL], cb[CQ_V3 * CBS + iR], cb[CQ_B1 * CBS + iL], cb[CQ_B1 * CBS + iR],
                                    cb[(CQ_Q0 + 4) * CBS + iL], cb[(CQ_Q0 + 4) * CBS + iR], cb[(CQ_Q0 + 5) * CBS + iL],
                                    cb[(CQ_Q0 + 5) * CBS + iR], cb[CQ_P * CBS + iL], cb[CQ_P * CBS + iR], a.gam, fc);
            double amax[7];
            face_amax(cb[CQ_V1 * CBS + iL], cb[CQ_LF * CBS + iL], cb[CQ_LA * CBS + iL], cb[CQ_LS * CBS + iL],
                      cb[CQ_V1 * CBS + iR], cb[CQ_LF * CBS + iR], cb[CQ_LA * CBS + iR], cb[CQ_LS * CBS + iR], amax);
            auto load = [&](int k, double F[7], double x[7]) {
                const int ii = T::cbi(sl[k], p);
#pragma unroll
                for (int m = 0; m < 7; ++m) {
                    F[m] = cb[(CQ_F0 + m) * CBS + ii];
                    x[m] = cb[(CQ_Q0 + m) * CBS + ii];
                }
            };
            double fh[7], Fs[7];
            RegStash st;
            if constexpr (BR == 0) {
                BackCoef bc;
                split_back(fc, bc);
                face_flux(fc, amax, a.eps, load, st, Fs);
                back_project(bc, Fs, fh);
            } else {
                BackCoefS bc;
                split_back_S(fc, bc);
                face_flux_br<1>(fc, amax, a.eps, load, st, Fs);
                back_project_S(bc, Fs, fh);
            }
            const int io = T::fbi(s + 1, p);
#pragma unroll
            for (int m = 0; m < 7; ++m) fb[m * FBS + io] = fh[m];
            // CT seed flux, Weno5_2D.jl:1005-1008: x sweep -> fsy (tangential comp 2),
            // y sweep -> gsx (tangential comp 3 of the rotated frame)
            if (!a.is1d && line < a.NP && (seg == 0 || f >= fA)) {
                constexpr int TV = DIR == 0 ? CQ_V2 : CQ_V3;
                const double bv = cb[CQ_B1 * CBS + iL] * cb[TV * CBS + iL] + cb[CQ_B1 * CBS + iR] * cb[TV * CBS + iR];
                const size_t g = DIR == 0 ? (size_t)f + (size_t)a.pitch * line : (size_t)line + (size_t)a.pitch * f;
                a.bs[g] = fma(0.5, bv, fh[DIR == 0 ? 4 : 5]);
            }
        }
        __syncthreads();

        // ---------------- update phase: cell c = s0 + s with faces c-1 (fb slot s) and c (slot s+1)
        {
            const int c = s0 + s;
            if constexpr (DIR == 1) {                            // the RK operands have landed
                if constexpr (TMA) { mbar_wait(&bars[1], phu); phu ^= 1u; } else cp_async_wait<1>();
            }
            if (upd_ok) {
                const int i0 = T::fbi(s, p), i1 = T::fbi(s + 1, p);
                const size_t g = DIR == 0 ? (size_t)c + (size_t)a.pitch * line : (size_t)line + (size_t)a.pitch * c;
                if constexpr (DIR == 0) {
                    constexpr int fi[6] = {0, 1, 2, 3, 5, 6};      // rho,Mx,My,Mz,Bz,E in fhat
#pragma unroll
                    for (int m = 0; m < 6; ++m) a.rhs[m][g] = fb[fi[m] * FBS + i1] - fb[fi[m] * FBS + i0];
                    if (a.is1d) a.rhs[6][g] = fb[4 * FBS + i1] - fb[4 * FBS + i0];   // By
                } else {
                    // rotated frame (rho,My,Mz,Mx,Bz,Bx,E): physical rho,Mx,My,Mz,Bz,E
                    constexpr int fi[6] = {0, 3, 1, 2, 4, 6};
                    const int ic = T::cbi(sl[2], p);
                    const double cx = a.c_dx * dt, cy = a.c_dy * dt;
#pragma unroll
                    for (int m = 0; m < 6; ++m) {
                        const double dfy = fb[fi[m] * FBS + i1] - fb[fi[m] * FBS + i0];
                        const double qcen = cb[(CQ_Q0 + fi[m]) * CBS + ic];
                        const double rxm = stu[m * T::FST + t];
                        double b = a.stage <= 3 ? stu[(6 + m) * T::FST + t] : 0.0;
                        const double acm = a.stage >= 3 ? stu[(12 + m) * T::FST + t] : 0.0;
                        if (a.stage == 2) a.acc[m][g] = qcen - b;                       // -q0 + q1
                        if (a.stage == 3) a.acc[m][g] = fma(2.0, qcen, acm);            // + 2 q2
                        if (a.stage == 4) b = (acm + qcen) * (1.0 / 3.0);               // Weno5_2D.jl:279
                        a.qn[m][g] = fma(-cx, rxm, fma(-cy, dfy, b));
                    }
                }
            }
        }
        __syncthreads();

        // the last face of this chunk is the left face of the next chunk's first cell; no barrier
        // needed after this: the next writes to fb happen behind the next cell-phase barrier
        for (int e = t; e < 7 * NPB; e += NT) {
            const int m = e / NPB, pp = e % NPB;
            fb[m * FBS + T::fbi(0, pp)] = fb[m * FBS + T::fbi(NS, pp)];
        }
        r0 += NS;
        while (r0 >= RS) r0 -= RS;
    }
}

// ------------------------------------------------------------------ warp-specialised flux kernel
// Same arithmetic as flux_kernel, different division of labour.  In flux_kernel every warp walks through
// cell phase -> barrier -> face phase -> barrier -> update phase; with two 252-register warps per SM
// sub-partition the FP64 pipe idles whenever one of them is in the (latency-bound, FP64-poor) cell or update
// phase: measured 72-74 % pipe utilisation although the face phase alone sustains 93 %.  Here a CTA has two
// warpgroups with the register file split by `setmaxnreg`:
//   warps 0-3  FACE warps (FREG registers): nothing but the face phase, back to back;
//   warps 4-7  HELPER warps (HREG registers): TMA issue, cell phase of the NEXT chunk, flux difference /
//              RK update of the PREVIOUS chunk, i.e. everything that is not dense FP64.
// Each sub-partition holds one face and one helper warp of each of the two resident CTAs, so the FP64 pipe
// always has two face warps to draw from and the helpers fill in.  Hand-over through four mbarriers
// (helper_done / face_done, two each so that a parity can never be lapped); the face warps never meet a
// CTA-wide barrier and may drift against each other by a march step.
// Ring of 2 NS + 5 cell slots (cells of chunk n+1 are written while chunk n is read), double-buffered face
// fluxes.  2-D, E branch, TMA staging (the 1-D, S-branch and no-tensor-map paths keep flux_kernel).
template <int DIR> struct WsTile {
    static constexpr int NT = 128, YW = 16;               // threads per role; columns of the y tile
    static constexpr int NS = DIR == 0 ? 64 : NT / YW;
    static constexpr int NPB = DIR == 0 ? NT / 64 : YW;
    static constexpr int RS = DIR == 0 ? (2 * NS + 5 + 15) / 16 * 16 : 2 * NS + 5;
    static constexpr int CBS = RS * NPB;
    static constexpr int FBS = (NS + 1) * NPB;
    static constexpr int BOX_X = DIR == 0 ? NS + 2 : NPB, BOX_Y = DIR == 0 ? NPB : NS;
    static constexpr int FST = BOX_X * BOX_Y;
    static constexpr int OFF_FB = CQ_COUNT * CBS;                              // [2][7][FBS]
    static constexpr int OFF_STQ = (OFF_FB + 14 * FBS + 15) / 16 * 16;         // [8][FST]   (TMA destination)
    static constexpr int OFF_STU = OFF_STQ + (8 * FST + 15) / 16 * 16;         // [18][FST]  (y sweep)
    static constexpr int OFF_BAR = OFF_STU + (DIR == 0 ? 0 : 18 * FST);
    static constexpr int SMEM = (OFF_BAR + 8) * (int)sizeof(double) + 16;
    __device__ static __forceinline__ int sti(int s, int p, int e) { return DIR == 0 ? p * BOX_X + s + e : s * NPB + p; }
    __device__ static __forceinline__ int cbi(int slot, int p) { return DIR == 0 ? p * RS + slot : slot * NPB + p; }
    __device__ static __forceinline__ int fbi(int slot, int p) { return DIR == 0 ? p * (NS + 1) + slot : slot * NPB + p; }
};

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void helper_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int DIR, int FREG, int HREG>
__global__ void __launch_bounds__(256, 2) flux_ws_kernel(const FluxArgs a, const __grid_constant__ FluxMaps tmap)
{
    using T = WsTile<DIR>;
    constexpr int NT = T::NT, NS = T::NS, NPB = T::NPB, RS = T::RS, CBS = T::CBS, FBS = T::FBS;
    extern __shared__ __align__(128) double sm[];
    double *cb = sm;
    double *fb = sm + T::OFF_FB;
    double *stq = sm + T::OFF_STQ;
    double *stu = sm + T::OFF_STU;
    // [0] state TMA, [1] operand TMA, [2],[3] helper_done, [4],[5] face_done (index = march step & 1)
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + T::OFF_BAR);

    const double dt = *a.dtp;
    if (dt == 0.0) return;                       // advance() past tmax: leave the state untouched

    const bool helper = threadIdx.x >= NT;
    const int t = threadIdx.x & (NT - 1);
    const int s = DIR == 0 ? (t % NS) : (t / NPB);
    const int p = DIR == 0 ? (t / NS) : (t % NPB);
    // x sweep of a y-slab whose ghost rows are still arriving (halo push of the previous stage): the strips that
    // hold ghost rows are mapped to the LAST blocks of the grid and wait for the neighbours' flag words themselves,
    // so the interior of the sweep hides the exchange (no separate wait kernel, no split launch)
    int strip = blockIdx.x, seg = blockIdx.y;
    if (DIR == 0 && a.wait_flags != nullptr) {
        const int L = blockIdx.x + gridDim.x * blockIdx.y;
        strip = L / gridDim.y + 2;
        seg = L % gridDim.y;
        if (strip >= (int)gridDim.x) strip -= gridDim.x;
    }
    const int line = strip * NPB + p;
    const int fA = 2 + seg * a.seg_len;
    const int fB = min(fA + a.seg_len, a.N - 3);
    const int fStart = seg == 0 ? fA : fA - 1;
    const int nsteps = (fB - fStart + NS - 1) / NS;

    const bool line_ok = line >= G && line < a.NP - G;

    if (threadIdx.x == 0) {
        if (DIR == 0 && a.wait_flags != nullptr) {
            const int l0 = strip * NPB, l1 = l0 + NPB - 1;
            const bool need[2] = {a.wait_lo && l0 < G, a.wait_hi && l1 >= a.NP - G};
            const long long t0 = clock64();
            for (int side = 0; side < 2; ++side) {
                if (!need[side]) continue;
                unsigned long long v;
                do {
                    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.wait_flags + side) : "memory");
                    if (v < a.wait_seq && clock64() - t0 > 20000000000ll) { a.wait_err[6] = 2.0; break; }   // neighbour died
                } while (v < a.wait_seq);
            }
        }
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
        mbar_init(&bars[2], NT); mbar_init(&bars[3], NT); mbar_init(&bars[4], NT); mbar_init(&bars[5], NT);
        mbar_fence_init();
    }
    __syncthreads();

    if (!helper) {
        // ============================================================ FACE warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(FREG));
        int r0 = 0;                              // ring slot of cell s0 - 2
        for (int n = 0; n < nsteps; ++n) {
            const int s0 = fStart + n * NS;
            const int b = n & 1;
            mbar_wait(&bars[2 + b], (unsigned)(n >> 1) & 1u);          // cells of this chunk ready, fb[b] free
            int sl[6];
            {
                int r = r0 + s;
                if (r >= RS) r -= RS;
#pragma unroll
                for (int k = 0; k < 6; ++k) { sl[k] = r; r = (r + 1 == RS) ? 0 : r + 1; }
            }
            const int f = s0 + s;
            if (f < fB) {
                const int iL = T::cbi(sl[2], p), iR = T::cbi(sl[3], p);
                FaceCoef fc;
                face_coefficients(cb[CQ_Q0 * CBS + iL], cb[CQ_Q0 * CBS + iR], cb[CQ_V1 * CBS + iL], cb[CQ_V1 * CBS + iR],
                                  cb[CQ_V2 * CBS + iL], cb[CQ_V2 * CBS + iR], cb[CQ_V3 * CBS + iL], cb[CQ_V3 * CBS + iR],
                                  cb[CQ_B1 * CBS + iL], cb[CQ_B1 * CBS + iR], cb[(CQ_Q0 + 4) * CBS + iL],
                                  cb[(CQ_Q0 + 4) * CBS + iR], cb[(CQ_Q0 + 5) * CBS + iL], cb[(CQ_Q0 + 5) * CBS + iR],
                                  cb[(CQ_Q0 + 6) * CBS + iL], cb[(CQ_Q0 + 6) * CBS + iR], a.gam, a.g2, fc);
                double amax[7];
                face_amax(cb[CQ_V1 * CBS + iL], cb[CQ_LF * CBS + iL], cb[CQ_LA * CBS + iL], cb[CQ_LS * CBS + iL],
                          cb[CQ_V1 * CBS + iR], cb[CQ_LF * CBS + iR], cb[CQ_LA * CBS + iR], cb[CQ_LS * CBS + iR], amax);
                auto load = [&](int k, double F[7], double x[7]) {
                    const int ii = T::cbi(sl[k], p);
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
                const int io = b * 7 * FBS + T::fbi(s + 1, p);
#pragma unroll
                for (int m = 0; m < 7; ++m) fb[m * FBS + io] = fh[m];
                // CT seed flux, Weno5_2D.jl:1005-1008
                if (line < a.NP && (seg == 0 || f >= fA)) {
                    constexpr int TV = DIR == 0 ? CQ_V2 : CQ_V3;
                    const double bv = cb[CQ_B1 * CBS + iL] * cb[TV * CBS + iL] + cb[CQ_B1 * CBS + iR] * cb[TV * CBS + iR];
                    const size_t g = DIR == 0 ? (size_t)f + (size_t)a.pitch * line : (size_t)line + (size_t)a.pitch * f;
                    a.bs[g] = fma(0.5, bv, fh[DIR == 0 ? 4 : 5]);
                }
            }
            mbar_arrive(&bars[4 + b]);
            r0 += NS;
            if (r0 >= RS) r0 -= RS;
        }
    } else {
        // ============================================================ HELPER warps
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(HREG));
        unsigned phq = 0, phu = 0;
        if (t == 0) {
            tma_issue_state<DIR, NT, T::YW>(a, &tmap, stq, &bars[0], fStart + 3, strip);
            if constexpr (DIR == 1) tma_issue_operands<NT, T::YW>(a, &tmap, stu, &bars[1], fStart);
        }
        // the 5 leading stencil cells of the segment, straight from global memory -> ring slots 0..4
        for (int e = t; e < 5 * NPB; e += NT) {
            const int sl = DIR == 0 ? e % 5 : e / NPB;
            const int pp = DIR == 0 ? e / 5 : e % NPB;
            const size_t g = cell_index<DIR>(a, fStart - 2 + sl, strip * NPB + pp);
            double q[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) q[k] = __ldg(a.qc[perm<DIR>(k)] + g);
            cell_store<CBS, 0>(a, cb, q, T::cbi(sl, pp));
        }
        // iteration n: flux difference / RK update of chunk n-1, cells of chunk n+1
        int rc = 5;                              // ring slot of the first new cell of chunk n+1
        int ru = 2;                              // ring slot of the first centre cell of chunk n-1 (used from n = 1)
        for (int n = -1; n <= nsteps; ++n) {
            if (n >= 1) {
                const int m1 = n - 1, bu = m1 & 1;
                mbar_wait(&bars[4 + bu], (unsigned)(m1 >> 1) & 1u);    // faces of chunk n-1 are in fb[bu]
                if constexpr (DIR == 1) { mbar_wait(&bars[1], phu); phu ^= 1u; }
                const int s0 = fStart + m1 * NS;
                const int c = s0 + s;
                const bool upd_ok = (c > fStart) && (c < fB) && (c >= G) && (c < a.N - G) && line_ok;
                const double *fbu = fb + bu * 7 * FBS;
                if (upd_ok) {
                    const int i0 = T::fbi(s, p), i1 = T::fbi(s + 1, p);
                    const size_t g = DIR == 0 ? (size_t)c + (size_t)a.pitch * line : (size_t)line + (size_t)a.pitch * c;
                    if constexpr (DIR == 0) {
                        constexpr int fi[6] = {0, 1, 2, 3, 5, 6};      // rho,Mx,My,Mz,Bz,E in fhat
#pragma unroll
                        for (int m = 0; m < 6; ++m) a.rhs[m][g] = fbu[fi[m] * FBS + i1] - fbu[fi[m] * FBS + i0];
                    } else {
                        constexpr int fi[6] = {0, 3, 1, 2, 4, 6};      // rotated frame -> physical rho,Mx,My,Mz,Bz,E
                        int rr = ru + s;
                        if (rr >= RS) rr -= RS;
                        const int ic = T::cbi(rr, p);
                        const double cx = a.c_dx * dt, cy = a.c_dy * dt;
#pragma unroll
                        for (int m = 0; m < 6; ++m) {
                            const double dfy = fbu[fi[m] * FBS + i1] - fbu[fi[m] * FBS + i0];
                            const double qcen = cb[(CQ_Q0 + fi[m]) * CBS + ic];
                            const double rxm = stu[m * T::FST + t];
                            double bb = a.stage <= 3 ? stu[(6 + m) * T::FST + t] : 0.0;
                            const double acm = a.stage >= 3 ? stu[(12 + m) * T::FST + t] : 0.0;
                            if (a.stage == 2) a.acc[m][g] = qcen - bb;                      // -q0 + q1
                            if (a.stage == 3) a.acc[m][g] = fma(2.0, qcen, acm);            // + 2 q2
                            if (a.stage == 4) bb = (acm + qcen) * (1.0 / 3.0);              // Weno5_2D.jl:279
                            a.qn[m][g] = fma(-cx, rxm, fma(-cy, dfy, bb));
                        }
                    }
                }
                // the last face of chunk n-1 is the left face of chunk n's first cell
                if (n < nsteps)
                    for (int e = t; e < 7 * NPB; e += NT) {
                        const int m = e / NPB, pp = e % NPB;
                        fb[(bu ^ 1) * 7 * FBS + m * FBS + T::fbi(0, pp)] = fbu[m * FBS + T::fbi(NS, pp)];
                    }
                ru += NS;
                if (ru >= RS) ru -= RS;
            }
            helper_sync();                       // every helper is done with stu and with the ring cells of chunk n-1
            if constexpr (DIR == 1) {
                if (t == 0 && n >= 1 && n < nsteps) tma_issue_operands<NT, T::YW>(a, &tmap, stu, &bars[1], fStart + n * NS);
            }
            if (n + 1 < nsteps) {
                mbar_wait(&bars[0], phq); phq ^= 1u;
                const int c0 = fStart + (n + 1) * NS + 3;               // first new cell of chunk n+1
                const int si = T::sti(s, p, c0 & 1);
                double q[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) q[k] = stq[perm<DIR>(k) * T::FST + si];
                int rr = rc + s;
                if (rr >= RS) rr -= RS;
                cell_store<CBS, 0>(a, cb, q, T::cbi(rr, p));
                rc += NS;
                if (rc >= RS) rc -= RS;
                helper_sync();                   // stq consumed
                if (t == 0 && n + 2 < nsteps) tma_issue_state<DIR, NT, T::YW>(a, &tmap, stq, &bars[0], c0 + NS, strip);
            }
            if (n + 1 < nsteps) mbar_arrive(&bars[2 + ((n + 1) & 1)]);
        }
    }
}

// ------------------------------------------------------------------ TMEM-stash flux kernel (3 warps per sub-partition)
// What caps the FP64 pipe at ~73 % is thread-level parallelism: a face needs ~250 live registers, so a sub-partition
// holds two warps, and a dependent DFMA issues only every 8 cycles.  Blackwell's tensor memory is idle in this
// kernel (no MMA anywhere): 256 KB per SM that tcgen05.st / tcgen05.ld move to and from registers, each warp owning
// the 32 lanes of its quadrant -- a per-thread stash with register-file-like addressing.  This variant parks the
// 28 minus-side differences (needed only after the stencil has streamed by) and the 17 back-projection coefficients
// there: 45 doubles = 90 of the CTA's 128 TMEM columns per thread.  That brings the kernel from 248 to <= 168
// registers WITHOUT local-memory spills (which is what sank the 168-register build of round 1: its spills
// overflowed the L1 left over next to 200 KB of shared memory), so THREE 128-thread CTAs are resident per SM:
// 3 warps per sub-partition.  Same phases and arithmetic as flux_kernel.
__device__ __forceinline__ void tm_st(unsigned taddr, double v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"((unsigned)__double2loint(v)),
                 "r"((unsigned)__double2hiint(v))
                 : "memory");
}
template <int ND> __device__ __forceinline__ void tm_ld(unsigned taddr, double v[ND]);
template <> __device__ __forceinline__ void tm_ld<4>(unsigned taddr, double v[4])
{
    unsigned r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
}
template <> __device__ __forceinline__ void tm_ld<8>(unsigned taddr, double v[8])
{
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int TM_COLS = 128;                     // TMEM columns per CTA (power of two >= 90)
constexpr int TM_Q0 = 0;                         // [7 fields][4] minus-side differences: columns 0..55
constexpr int TM_BACK0 = 56;                     // 17 (+ 7 padding) back-projection coefficients: columns 56..103

// the face flux with the stash in TMEM (difference-first form, see face_flux_diff); every lane of the warp must
// call it (tcgen05.ld / .st are warp-collective)
template <class Loader>
__device__ __forceinline__ void face_flux_tmem(const FaceCoef &c, const double amax[7], double eps, Loader load, unsigned tb, double Fs[7])
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
            if (k == 1) comb[m] = -Fb[m];
            if (k == 2) comb[m] = fma(7.0, Fb[m], comb[m]);
            if (k == 3) comb[m] = fma(7.0, Fb[m], comb[m]);
            if (k == 4) comb[m] = comb[m] - Fb[m];
            Fa[m] = Fb[m]; qa[m] = qb[m];
        }
        project(c, dF, dw);
        project(c, dq, dv);
#pragma unroll
        for (int m = 0; m < 7; ++m) {
            if (k <= 4) P[k - 1][m] = fma(amax[m], dv[m], dw[m]);
            if (k >= 2) tm_st(tb + TM_Q0 + (m * 4 + (k - 2)) * 2, fma(-amax[m], dv[m], dw[m]));
        }
    };
    kstep(1); kstep(2); kstep(3); kstep(4);
    kstep(5);
    double first[7];
    project(c, comb, first);
#pragma unroll
    for (int m = 0; m < 7; ++m) {
        const double second6 = weno_phi6(P[0][m], P[1][m], P[2][m], P[3][m], epsq);
        Fs[m] = fma(first[m], 1.0 / 12.0, (-1.0 / 6.0) * second6);
    }
    tm_wait_st();
#pragma unroll
    for (int m = 0; m < 7; ++m) {
        double Q[4];
        tm_ld<4>(tb + TM_Q0 + m * 8, Q);
        const double third6 = weno_phi6(Q[3], Q[2], Q[1], Q[0], epsq);
        Fs[m] = fma(1.0 / 6.0, third6, Fs[m]);
    }
}

template <int DIR> struct TmTile {
    static constexpr int NT = 128, YW = 16;
    static constexpr int NS = DIR == 0 ? 64 : NT / YW;
    static constexpr int NPB = DIR == 0 ? NT / 64 : YW;
    static constexpr int RS = DIR == 0 ? (NS + 5 + 15) / 16 * 16 : NS + 5;      // x: 80 (bank-conflict-free wrap), y: 13
    static constexpr int CBS = RS * NPB;
    static constexpr int FBS = (NS + 1) * NPB;
    static constexpr int BOX_X = DIR == 0 ? NS + 2 : NPB, BOX_Y = DIR == 0 ? NPB : NS;
    static constexpr int FST = BOX_X * BOX_Y;
    static constexpr int OFFQ = (CQ_COUNT * CBS + 7 * FBS + 15) / 16 * 16;
    static constexpr int STQ = 8 * FST;
    static constexpr int STU = DIR == 0 ? 0 : 18 * FST;
    static constexpr int SMEM = (OFFQ + (STQ + 15) / 16 * 16 + STU) * (int)sizeof(double) + 32;
    __device__ static __forceinline__ int sti(int s, int p, int e) { return DIR == 0 ? p * BOX_X + s + e : s * NPB + p; }
    __device__ static __forceinline__ int cbi(int slot, int p) { return DIR == 0 ? p * RS + slot : slot * NPB + p; }
    __device__ static __forceinline__ int fbi(int slot, int p) { return DIR == 0 ? p * (NS + 1) + slot : slot * NPB + p; }
};

template <int DIR>
__global__ void __launch_bounds__(128, 3) flux_tm_kernel(const FluxArgs a, const __grid_constant__ FluxMaps tmap)
{
    using T = TmTile<DIR>;
    constexpr int NT = T::NT, NS = T::NS, NPB = T::NPB, RS = T::RS, CBS = T::CBS, FBS = T::FBS;
    extern __shared__ __align__(128) double sm[];
    double *cb = sm;
    double *fb = sm + CQ_COUNT * CBS;
    double *stq = sm + T::OFFQ;
    double *stu = stq + (T::STQ + 15) / 16 * 16;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(stu + T::STU);   // [0] state, [1] operands
    unsigned *tmem_base = reinterpret_cast<unsigned *>(bars + 2);

    const double dt = *a.dtp;
    if (dt == 0.0) return;

    const int t = threadIdx.x;
    const int s = DIR == 0 ? (t % NS) : (t / NPB);
    const int p = DIR == 0 ? (t / NS) : (t % NPB);
    const int line = blockIdx.x * NPB + p;
    const int seg = blockIdx.y;
    const int fA = 2 + seg * a.seg_len;
    const int fB = min(fA + a.seg_len, a.N - 3);
    const int fStart = seg == 0 ? fA : fA - 1;
    const bool line_ok = line >= G && line < a.NP - G;

    // tensor memory: warp 0 allocates the CTA's columns; every warp addresses the 32 lanes of its own quadrant
    if (t < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base)), "r"(TM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (t == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned tb = *tmem_base + ((unsigned)((t >> 5) * 32) << 16);

    unsigned phq = 0, phu = 0;
    if (t == 0) tma_issue_state<DIR, NT, T::YW>(a, &tmap, stq, &bars[0], fStart + 3);
    for (int e = t; e < 5 * NPB; e += NT) {
        const int sl = DIR == 0 ? e % 5 : e / NPB;
        const int pp = DIR == 0 ? e / 5 : e % NPB;
        const size_t g = cell_index<DIR>(a, fStart - 2 + sl, blockIdx.x * NPB + pp);
        double q[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) q[k] = __ldg(a.qc[perm<DIR>(k)] + g);
        cell_store<CBS, 0>(a, cb, q, T::cbi(sl, pp));
    }

    int r0 = 0;
    for (int s0 = fStart; s0 < fB; s0 += NS) {
        int sl[6];
        {
            int r = r0 + s;
            if (r >= RS) r -= RS;
#pragma unroll
            for (int k = 0; k < 6; ++k) { sl[k] = r; r = (r + 1 == RS) ? 0 : r + 1; }
        }
        // ---------------- cell phase
        mbar_wait(&bars[0], phq); phq ^= 1u;
        {
            const int si = T::sti(s, p, (s0 + 3) & 1);
            double q[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) q[k] = stq[perm<DIR>(k) * T::FST + si];
            cell_store<CBS, 0>(a, cb, q, T::cbi(sl[5], p));
        }
        __syncthreads();

        const bool upd_ok = (s0 + s > fStart) && (s0 + s < fB) && (s0 + s >= G) && (s0 + s < a.N - G) && line_ok;
        if (t == 0) {
            if constexpr (DIR == 1) tma_issue_operands<NT, T::YW>(a, &tmap, stu, &bars[1], s0);
            if (s0 + NS < fB) tma_issue_state<DIR, NT, T::YW>(a, &tmap, stq, &bars[0], s0 + NS + 3);
        }

        // ---------------- face phase: every lane computes (warp-collective TMEM accesses); faces past the segment
        // work on whatever the ring holds and store nothing
        const int f = s0 + s;
        {
            const int iL = T::cbi(sl[2], p), iR = T::cbi(sl[3], p);
            double amax[7];
            face_amax(cb[CQ_V1 * CBS + iL], cb[CQ_LF * CBS + iL], cb[CQ_LA * CBS + iL], cb[CQ_LS * CBS + iL],
                      cb[CQ_V1 * CBS + iR], cb[CQ_LF * CBS + iR], cb[CQ_LA * CBS + iR], cb[CQ_LS * CBS + iR], amax);
            FaceCoef fc;
            face_coefficients(cb[CQ_Q0 * CBS + iL], cb[CQ_Q0 * CBS + iR], cb[CQ_V1 * CBS + iL], cb[CQ_V1 * CBS + iR],
                              cb[CQ_V2 * CBS + iL], cb[CQ_V2 * CBS + iR], cb[CQ_V3 * CBS + iL], cb[CQ_V3 * CBS + iR],
                              cb[CQ_B1 * CBS + iL], cb[CQ_B1 * CBS + iR], cb[(CQ_Q0 + 4) * CBS + iL],
                              cb[(CQ_Q0 + 4) * CBS + iR], cb[(CQ_Q0 + 5) * CBS + iL], cb[(CQ_Q0 + 5) * CBS + iR],
                              cb[(CQ_Q0 + 6) * CBS + iL], cb[(CQ_Q0 + 6) * CBS + iR], a.gam, a.g2, fc);
            {
                // the back-projection coefficients are cold until the stencil has streamed by: park them in TMEM
                BackCoef bc;
                split_back(fc, bc);
                const double *bv = reinterpret_cast<const double *>(&bc);
#pragma unroll
                for (int k = 0; k < 17; ++k) tm_st(tb + TM_BACK0 + 2 * k, bv[k]);
            }
            auto load = [&](int k, double F[7], double x[7]) {
                const int ii = T::cbi(sl[k], p);
#pragma unroll
                for (int m = 0; m < 7; ++m) {
                    F[m] = cb[(CQ_F0 + m) * CBS + ii];
                    x[m] = cb[(CQ_Q0 + m) * CBS + ii];
                }
            };
            double fh[7], Fs[7];
            face_flux_tmem(fc, amax, a.eps, load, tb, Fs);
            {
                BackCoef bc;
                double *bv = reinterpret_cast<double *>(&bc);
                tm_ld<8>(tb + TM_BACK0, bv);
                tm_ld<8>(tb + TM_BACK0 + 16, bv + 8);
                double last[4];
                tm_ld<4>(tb + TM_BACK0 + 32, last);
                bv[16] = last[0];
                back_project(bc, Fs, fh);
            }
            if (f < fB) {
                const int io = T::fbi(s + 1, p);
#pragma unroll
                for (int m = 0; m < 7; ++m) fb[m * FBS + io] = fh[m];
                if (line < a.NP && (seg == 0 || f >= fA)) {
                    constexpr int TV = DIR == 0 ? CQ_V2 : CQ_V3;
                    const double bv2 = cb[CQ_B1 * CBS + iL] * cb[TV * CBS + iL] + cb[CQ_B1 * CBS + iR] * cb[TV * CBS + iR];
                    const size_t g = DIR == 0 ? (size_t)f + (size_t)a.pitch * line : (size_t)line + (size_t)a.pitch * f;
                    a.bs[g] = fma(0.5, bv2, fh[DIR == 0 ? 4 : 5]);
                }
            }
        }
        __syncthreads();

        // ---------------- update phase
        {
            const int c = s0 + s;
            if constexpr (DIR == 1) { mbar_wait(&bars[1], phu); phu ^= 1u; }
            if (upd_ok) {
                const int i0 = T::fbi(s, p), i1 = T::fbi(s + 1, p);
                const size_t g = DIR == 0 ? (size_t)c + (size_t)a.pitch * line : (size_t)line + (size_t)a.pitch * c;
                if constexpr (DIR == 0) {
                    constexpr int fi[6] = {0, 1, 2, 3, 5, 6};
#pragma unroll
                    for (int m = 0; m < 6; ++m) a.rhs[m][g] = fb[fi[m] * FBS + i1] - fb[fi[m] * FBS + i0];
                } else {
                    constexpr int fi[6] = {0, 3, 1, 2, 4, 6};
                    const int ic = T::cbi(sl[2], p);
                    const double cx = a.c_dx * dt, cy = a.c_dy * dt;
#pragma unroll
                    for (int m = 0; m < 6; ++m) {
                        const double dfy = fb[fi[m] * FBS + i1] - fb[fi[m] * FBS + i0];
                        const double qcen = cb[(CQ_Q0 + fi[m]) * CBS + ic];
                        const double rxm = stu[m * T::FST + t];
                        double b = a.stage <= 3 ? stu[(6 + m) * T::FST + t] : 0.0;
                        const double acm = a.stage >= 3 ? stu[(12 + m) * T::FST + t] : 0.0;
                        if (a.stage == 2) a.acc[m][g] = qcen - b;
                        if (a.stage == 3) a.acc[m][g] = fma(2.0, qcen, acm);
                        if (a.stage == 4) b = (acm + qcen) * (1.0 / 3.0);
                        a.qn[m][g] = fma(-cx, rxm, fma(-cy, dfy, b));
                    }
                }
            }
        }
        __syncthreads();
        for (int e = t; e < 7 * NPB; e += NT) {
            const int m = e / NPB, pp = e % NPB;
            fb[m * FBS + T::fbi(0, pp)] = fb[m * FBS + T::fbi(NS, pp)];
        }
        r0 += NS;
        while (r0 >= RS) r0 -= RS;
    }
    __syncthreads();
    if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_base), "r"(TM_COLS));
}

// Launch variants, chosen once per process by WENO5_FLUX_VARIANT (development knob; the default is the
// fastest measured on B200):
//   2: flux_kernel, cp.async (LDGSTS) staging     4: flux_kernel, TMA tensor-map staging
//   5: flux_ws_kernel, warp-specialised, TMA staging
//   6: flux_tm_kernel, TMEM stash, 3 CTAs per SM, TMA staging
// All: x tile 2 rows x 64 faces, y tile 8 face rows x 16 columns per march step, 2 CTAs per SM.
// 1-D grids, the entropy branch and contexts without tensor maps always take flux_kernel with cp.async.
static int flux_variant()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("WENO5_FLUX_VARIANT");
        v = e ? atoi(e) : W5_DEFAULT_VARIANT;
        if (v != 2 && v != 4 && v != 5 && v != 6) v = W5_DEFAULT_VARIANT;
    }
    return v;
}

// Function attributes belong to the device's context: set the dynamic shared-memory limit once per device
// (a `static bool` would leave every device but the first at the 48 KB default).
template <auto kernel> static void ensure_smem(int bytes)      // one flag word per kernel instantiation
{
    static unsigned long long done = 0;          // bit d: attribute set on device d (devices >= 64: set every time)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && (__atomic_load_n(&done, __ATOMIC_ACQUIRE) >> dev & 1ull)) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (dev < 64) __atomic_fetch_or(&done, 1ull << dev, __ATOMIC_RELEASE);
}

template <int DIR, int NT, int YW, int MINB, bool TMA, int BR = 0>
static void launch_flux_t(const FluxArgs &a, cudaStream_t s)
{
    using T = Tile<DIR, NT, YW, BR == 0 ? CQ_COUNT : CQ_COUNT_S>;
    static_assert(T::SMEM * MINB <= 227 * 1024, "shared memory budget");
    ensure_smem<flux_kernel<DIR, NT, YW, MINB, TMA, BR>>(T::SMEM);
    dim3 grid((a.NP + T::NPB - 1) / T::NPB, a.nseg, 1);
    static const FluxMaps dummy = {};
    const FluxMaps *tm = reinterpret_cast<const FluxMaps *>(DIR == 0 ? a.tmap_x : a.tmap_y);
    flux_kernel<DIR, NT, YW, MINB, TMA, BR><<<grid, NT, T::SMEM, s>>>(a, TMA && tm ? *tm : dummy);
}

// register split of the two warpgroups (face + helper <= 256 = 2 x the 128 of __launch_bounds__(256, 2));
// WENO5_WS_SPLIT selects among the compiled splits (development knob)
template <int DIR, int FREG, int HREG> static void launch_flux_ws_t(const FluxArgs &a, cudaStream_t s)
{
    using T = WsTile<DIR>;
    static_assert((T::SMEM + 1024) * 2 <= 227 * 1024, "shared memory budget (2 CTAs per SM)");
    static_assert(FREG + HREG <= 256 && FREG % 8 == 0 && HREG % 8 == 0 && HREG >= 24, "register split of the two warpgroups");
    ensure_smem<flux_ws_kernel<DIR, FREG, HREG>>(T::SMEM);
    dim3 grid((a.NP + T::NPB - 1) / T::NPB, a.nseg, 1);
    const FluxMaps *tm = reinterpret_cast<const FluxMaps *>(DIR == 0 ? a.tmap_x : a.tmap_y);
    flux_ws_kernel<DIR, FREG, HREG><<<grid, 2 * T::NT, T::SMEM, s>>>(a, *tm);
}

template <int DIR> static void launch_flux_ws(const FluxArgs &a, cudaStream_t s)
{
    static const int split = getenv("WENO5_WS_SPLIT") ? atoi(getenv("WENO5_WS_SPLIT")) : 0;
    if (split == 1) launch_flux_ws_t<DIR, 216, 40>(a, s);
    else if (split == 2) launch_flux_ws_t<DIR, 232, 24>(a, s);
    else launch_flux_ws_t<DIR, 224, 32>(a, s);       // measured: 15.05 ms/step (216/40: 15.25, 208/48: 15.22)
}

template <int DIR> static void launch_flux_tm(const FluxArgs &a, cudaStream_t s)
{
    using T = TmTile<DIR>;
    static_assert((T::SMEM + 1024) * 3 <= 227 * 1024, "shared memory budget (3 CTAs per SM)");
    ensure_smem<flux_tm_kernel<DIR>>(T::SMEM);
    dim3 grid((a.NP + T::NPB - 1) / T::NPB, a.nseg, 1);
    const FluxMaps *tm = reinterpret_cast<const FluxMaps *>(DIR == 0 ? a.tmap_x : a.tmap_y);
    flux_tm_kernel<DIR><<<grid, T::NT, T::SMEM, s>>>(a, *tm);
}

template <int DIR>
static void launch_flux(const FluxArgs &a, cudaStream_t s)
{
    const bool tma_ok = a.tmap_x != nullptr && a.tmap_y != nullptr && !a.is1d;
    const int v = flux_variant();
    // entropy branch (dual-energy mode): its state planes are not consecutive in the pool, so it takes the
    // cp.async staging, which addresses every plane through its own pointer
    if (a.branch == 1) launch_flux_t<DIR, 128, 16, 2, false, 1>(a, s);
    else if (tma_ok && v == 6) launch_flux_tm<DIR>(a, s);
    else if (tma_ok && v == 5) launch_flux_ws<DIR>(a, s);
    else if (tma_ok && v == 4) launch_flux_t<DIR, 128, 16, 2, true>(a, s);
    else launch_flux_t<DIR, 128, 16, 2, false>(a, s);
}

void launch_flux_x(const FluxArgs &a, cudaStream_t s) { launch_flux<0>(a, s); }
void launch_flux_y(const FluxArgs &a, cudaStream_t s) { launch_flux<1>(a, s); }

// faces per march step along the sweep / lines per CTA (host-side planning; the same for every variant)
int flux_chunk(int dir) { return dir == 0 ? 64 : 8; }
int flux_lines_per_cta(int dir) { return dir == 0 ? 2 : 16; }
int flux_uses_tma() { return flux_variant() != 2; }
int flux_resident_ctas() { return (flux_variant() == 6 ? 3 : 2) * 148; }
int flux_waits_in_kernel() { return flux_variant() == 5 && !getenv("WENO5_NO_INKERNEL_WAIT"); }
void flux_box(int dir, int box[2])
{
    if (dir == 0) { box[0] = 64 + 2; box[1] = 2; } else { box[0] = 16; box[1] = 8; }
}

// ------------------------------------------------------------------ 1-D stage update
// Weno5_1D.jl:152-205: q_{k+1} = base - c dt/dx Q for the 7 flux-advanced components
// (Bx constant).  order of the 7: rho,Mx,My,Mz,Bz,E,By
struct Upd1dArgs {
    const double *rhs[7]; const double *qc[7]; const double *base[7]; double *acc[7]; double *qn[7];
    const double *dtp; double c_dx; int stage; int N;
};

__global__ void update_1d_kernel(const Upd1dArgs a)
{
    const double dt = *a.dtp;
    if (dt == 0.0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < G || i >= a.N - G) return;
    const double cx = a.c_dx * dt;
#pragma unroll
    for (int m = 0; m < 7; ++m) {
        const double qcen = a.qc[m][i];
        double b;
        if (a.stage <= 3) {
            b = a.base[m][i];
            if (a.stage == 2) a.acc[m][i] = qcen - b;
            if (a.stage == 3) a.acc[m][i] = fma(2.0, qcen, a.acc[m][i]);
        } else {
            b = (a.acc[m][i] + qcen) * (1.0 / 3.0);
        }
        a.qn[m][i] = fma(-cx, a.rhs[m][i], b);
    }
}

void launch_update_1d(const FluxArgs &a, double *const qn7[7], const double *const base7[7],
                      double *const acc7[7], cudaStream_t s)
{
    Upd1dArgs u;
    // current-state planes in the order rho,Mx,My,Mz,Bz,E,By
    const int cur[7] = {C_RHO, C_MX, C_MY, C_MZ, C_BZ, C_E, C_BY};
    for (int m = 0; m < 7; ++m) {
        u.rhs[m] = a.rhs[m]; u.qc[m] = a.qc[cur[m]]; u.base[m] = base7[m]; u.acc[m] = acc7[m]; u.qn[m] = qn7[m];
    }
    u.dtp = a.dtp; u.c_dx = a.c_dx; u.stage = a.stage; u.N = a.N;
    update_1d_kernel<<<(a.N + 255) / 256, 256, 0, s>>>(u);
}

// ------------------------------------------------------------------ constrained transport
// flux arrays with the reference's boundary rules applied on outflow sides
// (Weno5_2D.jl:1704-1729): rows/cols 1,2 := 3; Nx-1,Nx := Nx-2; then the rim := 0
__device__ __forceinline__ double ld_flux(const double *f, int i, int j, const CtArgs &a)
{
    if ((a.oxlo && i <= 1) || (a.oxhi && i >= a.Nx) || (a.oylo && j <= 1) || (a.oyhi && j >= a.Ny)) return 0.0;
    if (a.oxlo) i = max(i, 3);
    if (a.oxhi) i = min(i, a.Nx - 2);
    if (a.oylo) j = max(j, 3);
    if (a.oyhi) j = min(j, a.Ny - 2);
    return f[(size_t)i + (size_t)a.pitch * j];
}

__device__ __forceinline__ bool ct_outside(int i, int j, const CtArgs &a)
{
    return (a.oxlo && i < 2) || (a.oxhi && i > a.Nx - 1) || (a.oylo && j < 2) || (a.oyhi && j > a.Ny - 1);
}

// corner EMF, Weno5_2D.jl:328-332 (zero outside 2..N-1 on outflow sides)
__device__ __forceinline__ double corner(int i, int j, const CtArgs &a)
{
    if (ct_outside(i, j, a)) return 0.0;
    return 0.5 * (ld_flux(a.gsx, i, j, a) + ld_flux(a.gsx, i + 1, j, a) - ld_flux(a.fsy, i, j, a) -
                  ld_flux(a.fsy, i, j + 1, a));
}

__global__ void ct_faces_kernel(const CtArgs a)
{
    const double dt = *a.dtp;
    if (dt == 0.0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > a.Nx || j > a.Ny) return;
    const bool wbx = i >= a.bx_i0 && i <= a.bx_i1 && j >= a.bx_j0 && j <= a.bx_j1;
    const bool wby = i >= a.by_i0 && i <= a.by_i1 && j >= a.by_j0 && j <= a.by_j1;
    if (!wbx && !wby) return;
    const size_t g = (size_t)i + (size_t)a.pitch * j;
    const double cy = a.c_dy * dt, cx = a.c_dx * dt;
    // Away from outflow sides none of the boundary rules of ld_flux / ct_outside applies to the three corners
    // of this face pair: plain loads (the kernel was instruction-bound on the clamping logic: 435 -> ~90
    // instructions per thread; same arithmetic, bit-identical results)
    const bool outflow = a.oxlo | a.oxhi | a.oylo | a.oyhi;
    const bool lean = !outflow || (i >= 4 && i <= a.Nx - 3 && j >= 4 && j <= a.Ny - 3);
    double o, oj_, oi_;
    bool inner;
    if (lean) {
        const size_t p = (size_t)a.pitch;
        const double g00 = a.gsx[g], f00 = a.fsy[g];
        o = 0.5 * (g00 + a.gsx[g + 1] - f00 - a.fsy[g + p]);
        oj_ = 0.5 * (a.gsx[g - p] + a.gsx[g - p + 1] - a.fsy[g - p] - f00);
        oi_ = 0.5 * (a.gsx[g - 1] + g00 - a.fsy[g - 1] - a.fsy[g - 1 + p]);
        inner = true;
    } else {
        o = corner(i, j, a);
        // Oxi/Oxj are only filled on 2..N-1 (Weno5_2D.jl:334-339)
        inner = !ct_outside(i, j, a);
        oj_ = (inner && wbx) ? corner(i, j - 1, a) : 0.0;
        oi_ = (inner && wby) ? corner(i - 1, j, a) : 0.0;
    }
    if (wbx) {
        const double oj = inner ? oj_ : 0.0;
        const double cur = a.bxc[g];
        double b;
        if (a.stage <= 3) {
            b = a.bx0[g];
            if (a.stage == 2 && !a.no_acc) a.accbx[g] = cur - b;
            if (a.stage == 3 && !a.no_acc) a.accbx[g] = fma(2.0, cur, a.accbx[g]);
        } else {
            b = (a.accbx[g] + cur) * (1.0 / 3.0);
        }
        a.bxn[g] = b - cy * (o - oj);                       // Weno5_2D.jl:150
    }
    if (wby) {
        const double oi = inner ? oi_ : 0.0;
        const double cur = a.byc[g];
        double b;
        if (a.stage <= 3) {
            b = a.by0[g];
            if (a.stage == 2 && !a.no_acc) a.accby[g] = cur - b;
            if (a.stage == 3 && !a.no_acc) a.accby[g] = fma(2.0, cur, a.accby[g]);
        } else {
            b = (a.accby[g] + cur) * (1.0 / 3.0);
        }
        a.byn[g] = b - cx * (oi - o);                       // Weno5_2D.jl:151
    }
}

void launch_ct_faces(const CtArgs &a, cudaStream_t s)
{
    dim3 block(64, 4), grid((a.Nx + 63) / 64, (a.Ny + 3) / 4);
    ct_faces_kernel<<<grid, block, 0, s>>>(a);
}

// interpolateB_face2center, Weno5_2D.jl:344-356, interior cells (ghosts come from bc_fill)
__global__ void ct_centers_kernel(const double *__restrict__ bxn, const double *__restrict__ byn,
                                  double *__restrict__ Bx, double *__restrict__ By, int NXI, int NYI, int pitch,
                                  const double *dtp)
{
    if (*dtp == 0.0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x + G;
    const int j = blockIdx.y * blockDim.y + threadIdx.y + G;
    if (i >= NXI - G || j >= NYI - G) return;
    const size_t g = (size_t)i + (size_t)pitch * j;
    Bx[g] = (-bxn[g - 2] + 9.0 * bxn[g - 1] + 9.0 * bxn[g] - bxn[g + 1]) / 16.0;
    By[g] = (-byn[g - 2 * (size_t)pitch] + 9.0 * byn[g - pitch] + 9.0 * byn[g] - byn[g + pitch]) / 16.0;
}

void launch_ct_centers(const double *bxn, const double *byn, double *Bx, double *By, const Geom &g,
                       const double *dtp, cudaStream_t s)
{
    dim3 block(64, 4), grid((g.NXI - 2 * G + 63) / 64, (g.NYI - 2 * G + 3) / 4);
    ct_centers_kernel<<<grid, block, 0, s>>>(bxn, byn, Bx, By, g.NXI, g.NYI, g.pitch, dtp);
}

// boundaries! applied to the face fields themselves (Weno5_2D.jl:313 and :39), in place:
// targets lie outside the source range so there is no read-after-write hazard.
__global__ void face_bc_kernel(double *bx, double *by, int Nx, int Ny, int pitch, int sxlo, int sxhi, int sylo,
                               int syhi, const double *dtp)
{
    if (dtp && *dtp == 0.0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > Nx || j > Ny) return;
    const int nx = Nx - 2 * 3, ny = Ny - 2 * 3;
    // new(i,j) = 0 on the outflow rim, else old(map(i), map(j)); map clamps to 3..N-2 (outflow) or
    // wraps by the interior size (periodic); mapped indices are fixed points of the map
    int si = i, sj = j;
    bool zero = false;
    if (i <= 3) {
        if (sxlo == SIDE_OUTFLOW) { si = 3; zero |= i == 1; }
        else if (sxlo == SIDE_PERIODIC) si = i + nx;
    } else if (i >= Nx - 2) {
        if (sxhi == SIDE_OUTFLOW) { si = Nx - 2; zero |= i == Nx; }
        else if (sxhi == SIDE_PERIODIC) si = i - nx;
    }
    if (j <= 3) {
        if (sylo == SIDE_OUTFLOW) { sj = 3; zero |= j == 1; }
        else if (sylo == SIDE_PERIODIC) sj = j + ny;
    } else if (j >= Ny - 2) {
        if (syhi == SIDE_OUTFLOW) { sj = Ny - 2; zero |= j == Ny; }
        else if (syhi == SIDE_PERIODIC) sj = j - ny;
    }
    if (si == i && sj == j && !zero) return;
    const size_t g = (size_t)i + (size_t)pitch * j, sg = (size_t)si + (size_t)pitch * sj;
    bx[g] = zero ? 0.0 : bx[sg];
    by[g] = zero ? 0.0 : by[sg];
}

void launch_face_bc(double *bx, double *by, const Geom &g, const int side[4], const double *dtp, cudaStream_t s)
{
    dim3 block(64, 4), grid((g.Nx + 63) / 64, (g.Ny + 3) / 4);
    face_bc_kernel<<<grid, block, 0, s>>>(bx, by, g.Nx, g.Ny, g.pitch, side[0], side[1], side[2], side[3], dtp);
}

// ------------------------------------------------------------------ ghost fill of cell fields
struct BcArgs { double *pl[NCELL]; int n; int NXI, NYI, pitch; int sxlo, sxhi, sylo, syhi; const double *dtp; };

__device__ __forceinline__ int bc_map(int i, int N, int lo, int hi, bool &skip)
{
    const int n = N - 2 * G;
    if (i < G) {
        if (lo == SIDE_OUTFLOW) return G;
        if (lo == SIDE_PERIODIC) return i + n;
        skip = true; return i;
    }
    if (i >= N - G) {
        if (hi == SIDE_OUTFLOW) return N - G - 1;
        if (hi == SIDE_PERIODIC) return i - n;
        skip = true; return i;
    }
    return i;
}

// every ghost cell copies the interior cell its (x-fill then y-fill) chain ends at
// (Weno5_2D.jl:1686-1702); ghost rows of slab-internal sides are left to the exchange.
__global__ void bc_fill_kernel(const BcArgs a)
{
    if (a.dtp && *a.dtp == 0.0) return;
    const int nyg = a.NYI > 1 ? 2 * G : 0;
    const long nrowcells = (long)nyg * a.NXI;                   // ghost rows, all columns
    const long ncolcells = (long)(a.NYI - nyg) * 2 * G;         // interior rows, ghost columns
    const long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nrowcells + ncolcells) return;
    int i, j;
    if (id < nrowcells) {
        const int r = (int)(id / a.NXI);
        i = (int)(id % a.NXI);
        j = r < G ? r : a.NYI - 2 * G + r;
    } else {
        const long e = id - nrowcells;
        const int r = (int)(e / (2 * G)), c = (int)(e % (2 * G));
        j = (a.NYI > 1 ? G : 0) + r;
        i = c < G ? c : a.NXI - 2 * G + c;
    }
    bool skip = false;
    const int si = bc_map(i, a.NXI, a.sxlo, a.sxhi, skip);
    const int sj = a.NYI > 1 ? bc_map(j, a.NYI, a.sylo, a.syhi, skip) : j;
    if (skip) return;
    const size_t g = (size_t)i + (size_t)a.pitch * j, sg = (size_t)si + (size_t)a.pitch * sj;
    for (int k = 0; k < a.n; ++k) a.pl[k][g] = a.pl[k][sg];
}

void launch_bc_fill(double *const planes[], int nplanes, const Geom &g, const int side[4], const double *dtp,
                    cudaStream_t s)
{
    BcArgs a;
    a.n = nplanes;
    for (int k = 0; k < nplanes; ++k) a.pl[k] = planes[k];
    a.NXI = g.NXI; a.NYI = g.NYI; a.pitch = g.pitch;
    a.sxlo = side[0]; a.sxhi = side[1]; a.sylo = side[2]; a.syhi = side[3];
    a.dtp = dtp;
    const int nyg = g.NYI > 1 ? 2 * G : 0;
    const long total = (long)nyg * g.NXI + (long)(g.NYI - nyg) * 2 * G;
    bc_fill_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a);
}

// ------------------------------------------------------------------ halo packing for the slab exchange
// One contiguous message per neighbour and stage instead of one per plane: rows [NYI-2G, NYI-G) of
// every plane -> up buffer, rows [G, 2G) -> down buffer (pack); received rows -> ghost rows (unpack).
struct PackArgs { double *pl[NCELL]; int n; int NYI, pitch; double *up, *down; int has_up, has_down; };

__global__ void halo_pack_kernel(const PackArgs a, int unpack)
{
    const size_t per = (size_t)G * a.pitch;                 // doubles per plane and direction
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= per * a.n) return;
    const int k = (int)(e / per);
    const size_t r = e % per;
    double *p = a.pl[k];
    if (!unpack) {
        if (a.has_up) a.up[e] = p[(size_t)(a.NYI - 2 * G) * a.pitch + r];
        if (a.has_down) a.down[e] = p[(size_t)G * a.pitch + r];
    } else {
        if (a.has_up) p[(size_t)(a.NYI - G) * a.pitch + r] = a.up[e];
        if (a.has_down) p[r] = a.down[e];
    }
}

void launch_halo_pack(double *const planes[], int n, const Geom &g, double *up, double *down, int has_up, int has_down,
                      int unpack, cudaStream_t s)
{
    PackArgs a;
    a.n = n;
    for (int k = 0; k < n; ++k) a.pl[k] = planes[k];
    a.NYI = g.NYI; a.pitch = g.pitch; a.up = up; a.down = down; a.has_up = has_up; a.has_down = has_down;
    const size_t total = (size_t)G * g.pitch * n;
    halo_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a, unpack);
}

// ------------------------------------------------------------------ pipelined host step: rows of the input image -> slab planes
// The host arrays of weno5_pipe_step are uploaded once, row range by row range, into an image of the whole grid
// (public layout, unpadded); a slab takes the rows of its window from there.
struct GatherArgs { const double *src[11]; double *dst[11]; int n, Nx, spitch, dpitch, rows; };

__global__ void gather_rows_kernel(const GatherArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (i >= a.Nx) return;
    const size_t so = (size_t)a.spitch * r + i, dof = (size_t)a.dpitch * r + i;
    for (int k = 0; k < a.n; ++k) a.dst[k][dof] = a.src[k][so];
}

void launch_gather_rows(const double *const src[], double *const dst[], int n, int Nx, int spitch, int dpitch, int rows,
                        cudaStream_t s)
{
    if (rows <= 0) return;
    GatherArgs a;
    a.n = n;
    for (int k = 0; k < n; ++k) { a.src[k] = src[k]; a.dst[k] = dst[k]; }
    a.Nx = Nx; a.spitch = spitch; a.dpitch = dpitch; a.rows = rows;
    gather_rows_kernel<<<dim3((Nx + 255) / 256, rows), 256, 0, s>>>(a);
}

// ------------------------------------------------------------------ halo push over peer memory (NVLink)
// One process per GPU; at weno5_comm_init every rank maps its neighbours' plane pools (CUDA IPC).  A stage's
// exchange is then ONE kernel that stores this slab's 4 outermost interior rows of every plane straight into
// the neighbours' ghost rows and, once the last block is done, raises a sequence number in the neighbour's
// flag word; the neighbour's stream waits for that number with a one-thread kernel before it touches its ghost
// rows.  No pack / unpack passes, no staging buffers, no NCCL kernel competing for SMs.
__global__ void halo_push_kernel(const PushArgs a)
{
    const size_t per = (size_t)G * a.pitch;                 // doubles per plane and direction
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < per * a.n) {
        const int k = (int)(e / per);
        const size_t r = e % per;
        const double *src = a.src[k];
        if (a.peer_up) a.peer_up[(size_t)a.plane_idx[k] * a.plane_up + r] = src[(size_t)(a.NYI - 2 * G) * a.pitch + r];
        if (a.peer_down)
            a.peer_down[(size_t)a.plane_idx[k] * a.plane_down + (size_t)(a.NYI_down - G) * a.pitch + r] = src[(size_t)G * a.pitch + r];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(a.counter, 1u) + 1u;
        if (done == gridDim.x) {
            *a.counter = 0u;
            __threadfence_system();
            if (a.flag_up) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.flag_up), "l"(a.seq) : "memory");
            if (a.flag_down) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.flag_down), "l"(a.seq) : "memory");
        }
    }
}

void launch_halo_push(const PushArgs &a, cudaStream_t s)
{
    const size_t total = (size_t)G * a.pitch * a.n;
    halo_push_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a);
}

// flags[0]: sequence number of the last push from the lower neighbour, flags[1]: from the upper one.
// Gives up after ~10 s (a neighbour died): raises the bad-state flag of the control block instead of hanging the GPU.
__global__ void halo_wait_kernel(const unsigned long long *flags, int need_down, int need_up, unsigned long long seq, double *ctl)
{
    const long long t0 = clock64();
    for (int side = 0; side < 2; ++side) {
        if (!(side == 0 ? need_down : need_up)) continue;
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + side) : "memory");
            if (v < seq && clock64() - t0 > 20000000000ll) { ctl[6] = 2.0; return; }
        } while (v < seq);
    }
}

// importer bookkeeping: word [3] of a rank's flag words counts the processes that currently map its memory
__global__ void peer_count_kernel(unsigned long long *counter, long long delta)
{
    atomicAdd_system(counter, (unsigned long long)delta);
    __threadfence_system();
}

void launch_peer_count(unsigned long long *counter, long long delta, cudaStream_t s) { peer_count_kernel<<<1, 1, 0, s>>>(counter, delta); }

void launch_halo_wait(const unsigned long long *flags, int need_down, int need_up, unsigned long long seq, double *ctl,
                      cudaStream_t s)
{
    halo_wait_kernel<<<1, 1, 0, s>>>(flags, need_down, need_up, seq, ctl);
}

// ------------------------------------------------------------------ entropy variable
// ES_Switch residue (Weno5_2D.jl:1737-1765): S = E2S(q); with roundtrip E = S2E(q,S)
struct EntArgs { double *q[9]; int NXI, NYI, pitch; double gam; int roundtrip; int all_cells; const double *dtp; };

__global__ void entropy_kernel(const EntArgs a)
{
    if (a.dtp && *a.dtp == 0.0) return;
    const int off = a.all_cells ? 0 : G;
    const int i = blockIdx.x * blockDim.x + threadIdx.x + off;
    const int j = a.NYI > 1 ? blockIdx.y * blockDim.y + threadIdx.y + off : 0;
    if (i >= a.NXI - off || (a.NYI > 1 && j >= a.NYI - off)) return;
    const size_t g = (size_t)i + (size_t)a.pitch * j;
    const double rho = a.q[C_RHO][g], mx = a.q[C_MX][g], my = a.q[C_MY][g], mz = a.q[C_MZ][g];
    const double bx = a.q[C_BX][g], by = a.q[C_BY][g], bz = a.q[C_BZ][g];
    double S, E2;
    e2s_s2e(rho, mx, my, mz, bx, by, bz, a.q[C_E][g], a.gam, S, E2);
    a.q[C_S][g] = S;
    if (a.roundtrip) a.q[C_E][g] = E2;
}

// interpolateB_face2center (:344-356) and the ES_Switch residue (:1737-1765) of the same stage in one pass
// over the interior cells: the new Bx, By never travel to HBM and back between the two (96 B instead of
// 112 B per cell and one launch less).  Same arithmetic, same order as the two separate kernels.
__global__ void ct_centers_entropy_kernel(const EntArgs a, const double *__restrict__ bxn, const double *__restrict__ byn)
{
    if (a.dtp && *a.dtp == 0.0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x + G;
    const int j = blockIdx.y * blockDim.y + threadIdx.y + G;
    if (i >= a.NXI - G || j >= a.NYI - G) return;
    const size_t g = (size_t)i + (size_t)a.pitch * j;
    const size_t pitch = (size_t)a.pitch;
    const double bx = (-bxn[g - 2] + 9.0 * bxn[g - 1] + 9.0 * bxn[g] - bxn[g + 1]) / 16.0;
    const double by = (-byn[g - 2 * pitch] + 9.0 * byn[g - pitch] + 9.0 * byn[g] - byn[g + pitch]) / 16.0;
    a.q[C_BX][g] = bx;
    a.q[C_BY][g] = by;
    const double rho = a.q[C_RHO][g], mx = a.q[C_MX][g], my = a.q[C_MY][g], mz = a.q[C_MZ][g];
    const double bz = a.q[C_BZ][g];
    double S, E2;
    e2s_s2e(rho, mx, my, mz, bx, by, bz, a.q[C_E][g], a.gam, S, E2);
    a.q[C_S][g] = S;
    if (a.roundtrip) a.q[C_E][g] = E2;
}

void launch_ct_centers_entropy(const double *bxn, const double *byn, double *const q9[9], const Geom &g, double gam,
                               int roundtrip, const double *dtp, cudaStream_t s)
{
    EntArgs a;
    for (int k = 0; k < 9; ++k) a.q[k] = q9[k];
    a.NXI = g.NXI; a.NYI = g.NYI; a.pitch = g.pitch; a.gam = gam; a.roundtrip = roundtrip;
    a.all_cells = 0; a.dtp = dtp;
    dim3 block(64, 4), grid((g.NXI - 2 * G + 63) / 64, (g.NYI - 2 * G + 3) / 4);
    ct_centers_entropy_kernel<<<grid, block, 0, s>>>(a, bxn, byn);
}

// Stage 4 only: the same pass also yields the wave-speed maxima of the NEW state, i.e. compute_global_vmax
// (Weno5_2D.jl:1942-1984) of the next step's timestep() call -- the thread holds rho, M, B, S of its cell
// anyway; the x-face averages need the right-hand neighbour, which comes through shared memory.  A block owns
// 63 columns, its 64th thread evaluates the neighbour column redundantly (not written); at the last interior
// column the neighbour is the ghost cell bc_fill is about to create, i.e. the image of cell G (periodic) or the
// column itself (outflow).  Saves the separate 64 B/cell vmax pass once per step.  Same expressions as
// vmax_kernel, bit-identical maxima.
__global__ void ct_centers_entropy_vmax_kernel(const EntArgs a, const double *__restrict__ bxn, const double *__restrict__ byn,
                                               int xhi_side, unsigned long long *out)
{
    if (a.dtp && *a.dtp == 0.0) return;
    __shared__ double sh[7][4][64];
    __shared__ unsigned long long sx[8], sy[8];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.x * 63 + tx + G;
    const int j = blockIdx.y * blockDim.y + ty + G;
    const int iend = a.NXI - G;                        // first ghost column
    const bool row_ok = j < a.NYI - G;
    int ie = i;
    if (i == iend) ie = xhi_side == SIDE_PERIODIC ? G : iend - 1;
    const bool eval = row_ok && i <= iend;
    const bool own = row_ok && i < iend && tx < 63;
    double rho = 1.0, mx = 0.0, my = 0.0, bx = 0.0, by = 0.0, bz = 0.0, S = 0.0;
    if (eval) {
        const size_t g = (size_t)ie + (size_t)a.pitch * j;
        const size_t pitch = (size_t)a.pitch;
        bx = (-bxn[g - 2] + 9.0 * bxn[g - 1] + 9.0 * bxn[g] - bxn[g + 1]) / 16.0;
        by = (-byn[g - 2 * pitch] + 9.0 * byn[g - pitch] + 9.0 * byn[g] - byn[g + pitch]) / 16.0;
        rho = a.q[C_RHO][g]; mx = a.q[C_MX][g]; my = a.q[C_MY][g];
        const double mz = a.q[C_MZ][g];
        bz = a.q[C_BZ][g];
        double E2;
        e2s_s2e(rho, mx, my, mz, bx, by, bz, a.q[C_E][g], a.gam, S, E2);
        if (own) {
            a.q[C_BX][g] = bx;
            a.q[C_BY][g] = by;
            a.q[C_S][g] = S;
            if (a.roundtrip) a.q[C_E][g] = E2;
        }
    }
    sh[0][ty][tx] = rho; sh[1][ty][tx] = mx; sh[2][ty][tx] = my; sh[3][ty][tx] = bx;
    sh[4][ty][tx] = by; sh[5][ty][tx] = bz; sh[6][ty][tx] = S;
    __syncthreads();
    double vxm = 0.0, vym = 0.0;
    if (own) {
        const int t1 = tx + 1;
        const double rf = (rho + sh[0][ty][t1]) / 2;
        const double ri = fast_rcp(rf);
        const double vx = (mx + sh[1][ty][t1]) / 2 * ri;
        const double vy = (my + sh[2][ty][t1]) / 2 * ri;
        const double Bx = (bx + sh[3][ty][t1]) / 2;
        const double By = (by + sh[4][ty][t1]) / 2;
        const double Bz = (bz + sh[5][ty][t1]) / 2;
        const double Sf = (S + sh[6][ty][t1]) / 2;
        const double pres = Sf * pow_gm1(rf, a.gam);
        const double cs2 = a.gam * fabs(pres * ri);
        const double B2 = Bx * Bx + By * By + Bz * Bz;
        const double bbn2 = B2 * ri, bnx2 = Bx * Bx * ri, bny2 = By * By * ri;
        const double sm = bbn2 + cs2;
        const double rootx = fast_sqrt(fmax(0.0, sm * sm - 4 * bnx2 * cs2));
        const double rooty = fast_sqrt(fmax(0.0, sm * sm - 4 * bny2 * cs2));
        vxm = fabs(vx) + fast_sqrt(0.5 * (sm + rootx));
        vym = fabs(vy) + fast_sqrt(0.5 * (sm + rooty));
        if (sm != sm) { vxm = sm; vym = sm; }
    }
    unsigned long long bx_ = (unsigned long long)__double_as_longlong(vxm);
    unsigned long long by_ = (unsigned long long)__double_as_longlong(vym);
    if (vxm != vxm || vxm < 0.0) bx_ = ~0ull;
    if (vym != vym || vym < 0.0) by_ = ~0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bx_ = max(bx_, __shfl_xor_sync(0xffffffffu, bx_, o));
        by_ = max(by_, __shfl_xor_sync(0xffffffffu, by_, o));
    }
    const int tid = ty * 64 + tx;
    if ((tid & 31) == 0) { sx[tid >> 5] = bx_; sy[tid >> 5] = by_; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) { bx_ = max(bx_, sx[w]); by_ = max(by_, sy[w]); }
        atomicMax(out, bx_);
        atomicMax(out + 1, by_);
    }
}

void launch_ct_centers_entropy_vmax(const double *bxn, const double *byn, double *const q9[9], const Geom &g, double gam,
                                    int roundtrip, const double *dtp, int xhi_side, unsigned long long *vmax_out,
                                    cudaStream_t s)
{
    EntArgs a;
    for (int k = 0; k < 9; ++k) a.q[k] = q9[k];
    a.NXI = g.NXI; a.NYI = g.NYI; a.pitch = g.pitch; a.gam = gam; a.roundtrip = roundtrip;
    a.all_cells = 0; a.dtp = dtp;
    dim3 block(64, 4), grid((g.NXI - 2 * G + 62) / 63, (g.NYI - 2 * G + 3) / 4);
    ct_centers_entropy_vmax_kernel<<<grid, block, 0, s>>>(a, bxn, byn, xhi_side, vmax_out);
}

void launch_entropy(double *const q9[9], const Geom &g, double gam, int roundtrip, int all_cells,
                    const double *dtp, cudaStream_t s)
{
    EntArgs a;
    for (int k = 0; k < 9; ++k) a.q[k] = q9[k];
    a.NXI = g.NXI; a.NYI = g.NYI; a.pitch = g.pitch; a.gam = gam; a.roundtrip = roundtrip;
    a.all_cells = all_cells; a.dtp = dtp;
    dim3 block(64, 4), grid((g.NXI + 63) / 64, g.NYI > 1 ? (g.NYI + 3) / 4 : 1);
    entropy_kernel<<<grid, block, 0, s>>>(a);
}

// ------------------------------------------------------------------ dual-energy switch
// ES_Switch, Weno5_2D.jl:1734-1768, with the reference's commented-out criterion (:1740) live:
// dS = |E2S(q_E) - S(q_S)|; cells with dS >= threshold take the E state (with S = E2S(q_E)), all others the
// S state; q[:,:,9] = S2E(q) of the selected state with the B of its own candidate (:1765).
__global__ void es_switch_kernel(const SwitchArgs a)
{
    if (a.dtp && *a.dtp == 0.0) return;
    const bool d1 = a.NYI == 1;                                  // 1-D: Weno5_1D.jl:238-258
    const int i = blockIdx.x * blockDim.x + threadIdx.x + G;
    const int j = d1 ? 0 : blockIdx.y * blockDim.y + threadIdx.y + G;
    if (i >= a.NXI - G || (!d1 && j >= a.NYI - G)) return;
    const size_t g = (size_t)i + (size_t)a.pitch * j;
    double e[8], s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { e[k] = a.ce[k][g]; s[k] = a.cs[k][g]; }
    const double SE = e2s(e[0], e[1], e[2], e[3], e[6], e[7], e[4], e[5], a.gam);
    const bool bad = fabs(SE - s[5]) >= a.threshold;
    // flags_only (the 1-D reference as shipped, :252-254): the energy state everywhere, the flags are the output
    const bool takeE = bad || a.flags_only;
    const double rho = takeE ? e[0] : s[0], mx = takeE ? e[1] : s[1], my = takeE ? e[2] : s[2], mz = takeE ? e[3] : s[3];
    const double bz = takeE ? e[4] : s[4], S = takeE ? SE : s[5], bx = takeE ? e[6] : s[6], by = takeE ? e[7] : s[7];
    a.qn[0][g] = rho; a.qn[1][g] = mx; a.qn[2][g] = my; a.qn[3][g] = mz; a.qn[4][g] = bz;
    a.qn[5][g] = S;
    a.qn[6][g] = a.flags_only ? e[5] : s2e(rho, mx, my, mz, bx, by, bz, S, a.gam);
    if (a.qn_by) a.qn_by[g] = by;                                // 1-D: By is a flux-advanced component
    a.flag[g] = bad ? 1.0 : 0.0;
}

void launch_es_switch(const SwitchArgs &a, cudaStream_t s)
{
    dim3 block(64, 4), grid((a.NXI - 2 * G + 63) / 64, a.NYI > 1 ? (a.NYI - 2 * G + 3) / 4 : 1);
    if (a.NYI == 1) block = dim3(256, 1), grid = dim3((a.NXI - 2 * G + 255) / 256, 1);
    es_switch_kernel<<<grid, block, 0, s>>>(a);
}

// fsy[i,j], fsy[i-1,j], gsx[i,j], gsx[i,j-1] of a switched cell (i,j) come from the E sweep (:1754-1758); the
// reference's loop starts at 2, so the outermost public row/column never switches.
__global__ void flux_select_kernel(const double *__restrict__ flag, const double *__restrict__ fsyE,
                                   const double *__restrict__ gsxE, double *__restrict__ fsyS, double *__restrict__ gsxS,
                                   int NXI, int NYI, int pitch, const double *dtp)
{
    if (dtp && *dtp == 0.0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= NXI - 1 || j >= NYI - 1) return;
    const size_t g = (size_t)i + (size_t)pitch * j;
    auto bad = [&](int ii, int jj) { return ii >= 2 && jj >= 2 && flag[(size_t)ii + (size_t)pitch * jj] != 0.0; };
    const bool here = bad(i, j);
    if (here || bad(i + 1, j)) fsyS[g] = fsyE[g];
    if (here || bad(i, j + 1)) gsxS[g] = gsxE[g];
}

void launch_flux_select(const double *flag, const double *fsyE, const double *gsxE, double *fsyS, double *gsxS,
                        const Geom &g, const double *dtp, cudaStream_t s)
{
    dim3 block(64, 4), grid((g.NXI + 63) / 64, (g.NYI + 3) / 4);
    flux_select_kernel<<<grid, block, 0, s>>>(flag, fsyE, gsxE, fsyS, gsxS, g.NXI, g.NYI, g.pitch, dtp);
}

// ------------------------------------------------------------------ CFL reduction
// compute_global_vmax, Weno5_2D.jl:1942-1984 (x-face averages for both directions, pressure
// from S); 1-D twin Weno5_1D.jl:1038-1074.  Non-negative doubles order like their bit
// patterns, so the grid-wide maximum is an integer atomicMax; NaN maps to the largest
// pattern and is caught by dt_kernel.
struct VmaxArgs { const double *q[9]; int NXI, NYI, pitch; double gam; unsigned long long *out; int j0, j1; };

__global__ void vmax_kernel(const VmaxArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + G;
    const int j = a.NYI > 1 ? blockIdx.y * blockDim.y + threadIdx.y + a.j0 : 0;
    double vxm = 0.0, vym = 0.0;
    if (i < a.NXI - G && (a.NYI == 1 || j < a.j1)) {
        const size_t g = (size_t)i + (size_t)a.pitch * j;
        // same quantities as the reference loop (Weno5_2D.jl:1950-1979); the six divisions by the face
        // density share one branch-free reciprocal and the square roots are branch-free (~1 ulp)
        const double rho = (a.q[C_RHO][g] + a.q[C_RHO][g + 1]) / 2;
        const double ri = fast_rcp(rho);
        const double vx = (a.q[C_MX][g] + a.q[C_MX][g + 1]) / 2 * ri;
        const double vy = (a.q[C_MY][g] + a.q[C_MY][g + 1]) / 2 * ri;
        const double Bx = (a.q[C_BX][g] + a.q[C_BX][g + 1]) / 2;
        const double By = (a.q[C_BY][g] + a.q[C_BY][g + 1]) / 2;
        const double Bz = (a.q[C_BZ][g] + a.q[C_BZ][g + 1]) / 2;
        const double S = (a.q[C_S][g] + a.q[C_S][g + 1]) / 2;
        const double pres = S * pow_gm1(rho, a.gam);
        const double cs2 = a.gam * fabs(pres * ri);
        const double B2 = Bx * Bx + By * By + Bz * Bz;
        const double bbn2 = B2 * ri, bnx2 = Bx * Bx * ri, bny2 = By * By * ri;
        const double sm = bbn2 + cs2;
        const double rootx = fast_sqrt(fmax(0.0, sm * sm - 4 * bnx2 * cs2));
        const double rooty = fast_sqrt(fmax(0.0, sm * sm - 4 * bny2 * cs2));
        vxm = fabs(vx) + fast_sqrt(0.5 * (sm + rootx));
        vym = fabs(vy) + fast_sqrt(0.5 * (sm + rooty));
        if (sm != sm) { vxm = sm; vym = sm; }          // keep a NaN visible whatever the clamps above did
    }
    unsigned long long bx_ = (unsigned long long)__double_as_longlong(vxm);
    unsigned long long by_ = (unsigned long long)__double_as_longlong(vym);
    if (vxm != vxm || vxm < 0.0) bx_ = ~0ull;
    if (vym != vym || vym < 0.0) by_ = ~0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bx_ = max(bx_, __shfl_xor_sync(0xffffffffu, bx_, o));
        by_ = max(by_, __shfl_xor_sync(0xffffffffu, by_, o));
    }
    __shared__ unsigned long long sx[8], sy[8];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if ((tid & 31) == 0) { sx[tid >> 5] = bx_; sy[tid >> 5] = by_; }
    __syncthreads();
    if (tid == 0) {
        const int nw = (blockDim.x * blockDim.y + 31) / 32;
        for (int w = 1; w < nw; ++w) { bx_ = max(bx_, sx[w]); by_ = max(by_, sy[w]); }
        atomicMax(a.out, bx_);
        atomicMax(a.out + 1, by_);
    }
}

void launch_vmax(const double *const q9[9], const Geom &g, double gam, unsigned long long *vmax_bits,
                 cudaStream_t s)
{
    launch_vmax_rows(q9, g, gam, vmax_bits, G, g.NYI - G, 1, s);
}

// rows [j0, j1) (internal indices) only; with zero_first == 0 the maxima accumulate into vmax_bits
void launch_vmax_rows(const double *const q9[9], const Geom &g, double gam, unsigned long long *vmax_bits, int j0, int j1,
                      int zero_first, cudaStream_t s)
{
    VmaxArgs a;
    for (int k = 0; k < 9; ++k) a.q[k] = q9[k];
    a.NXI = g.NXI; a.NYI = g.NYI; a.pitch = g.pitch; a.gam = gam; a.out = vmax_bits;
    a.j0 = g.NYI > 1 ? j0 : 0; a.j1 = g.NYI > 1 ? j1 : 1;
    if (zero_first) cudaMemsetAsync(vmax_bits, 0, 2 * sizeof(unsigned long long), s);
    dim3 block(64, 4), grid((g.NXI - 2 * G + 63) / 64, g.NYI > 1 ? (a.j1 - a.j0 + 3) / 4 : 1);
    vmax_kernel<<<grid, block, 0, s>>>(a);
}

// timestep(), Weno5_2D.jl:1930-1940; 1-D: dt = courFac*dx/vmax (Weno5_1D.jl:55).
// With tmax > 0 the step is shortened to land on tmax and becomes 0 once t >= tmax.
// With a communicator whose ranks map each other's flag words (peer memory, weno5_comm_init) the max-all-reduce
// of the two wave speeds happens right here: lane r stores this rank's two bit patterns and the sequence number into
// rank r's slot array (over NVLink for r != own rank), then waits for rank r's entry in the local array, and a warp
// maximum finishes the reduction -- no NCCL kernel, no hop to a second stream and back.  Slots are double-buffered by
// sequence parity: a rank can be at most one reduction ahead of the slowest (it needs everybody's entry to finish).
__global__ void dt_kernel(unsigned long long *vb, double *ctl, double dx, double dy, double cfl, int is1d, int clear, const ShareArgs sh)
{
    const int lane = threadIdx.x;
    unsigned long long b0 = vb[0], b1 = vb[1];
    if (sh.n > 1) {
        const int par = (int)(sh.seq & 1ull);
        unsigned long long m0 = 0ull, m1 = 0ull;
        if (lane < sh.n) {
            unsigned long long *dst = sh.slots[lane] + (size_t)(par * 32 + sh.rank) * 4;
            dst[0] = b0; dst[1] = b1;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst + 2), "l"(sh.seq) : "memory");
            const unsigned long long *src = sh.slots[sh.rank] + (size_t)(par * 32 + lane) * 4;
            const long long t0 = clock64();
            unsigned long long v;
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(src + 2) : "memory");
                if (v < sh.seq && clock64() - t0 > 20000000000ll) { ctl[6] = 2.0; break; }      // a rank never arrived
            } while (v < sh.seq);
            m0 = *(volatile const unsigned long long *)(src + 0);
            m1 = *(volatile const unsigned long long *)(src + 1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m0 = max(m0, __shfl_xor_sync(0xffffffffu, m0, o));
            m1 = max(m1, __shfl_xor_sync(0xffffffffu, m1, o));
        }
        b0 = m0; b1 = m1;
    }
    __syncwarp();
    if (lane != 0) return;
    const double vx = __longlong_as_double((long long)b0);
    const double vy = __longlong_as_double((long long)b1);
    if (clear) { vb[0] = 0ull; vb[1] = 0ull; }             // accumulator of the fused stage-4 reduction: ready for the next step
    else if (sh.n > 1) { vb[0] = b0; vb[1] = b1; }
    const double t = ctl[1], tmax = ctl[4];
    if (tmax > 0.0 && t >= tmax) { ctl[0] = 0.0; return; } // finished: the kernels of further steps return at once
    ctl[2] = vx; ctl[3] = vy;
    double dt;
    if (is1d) {
        dt = cfl * dx / vx;
    } else {
        const double dtx = dx / vx, dty = dy / vy;
        dt = cfl / (1 / dtx + 1 / dty);
    }
    if (!(dt > 0.0) || !(dt < 1e300)) { ctl[6] = 1.0; dt = 0.0; }
    if (tmax > 0.0 && t + dt > tmax) dt = tmax - t;
    ctl[0] = dt;
}

void launch_dt_from_vmax(unsigned long long *vmax_bits, double *ctl, double dx, double dy, double cfl,
                         int is1d, int clear, const ShareArgs *share, cudaStream_t s)
{
    ShareArgs none;
    none.n = 0; none.rank = 0; none.seq = 0;
    dt_kernel<<<1, 32, 0, s>>>(vmax_bits, ctl, dx, dy, cfl, is1d, clear, share ? *share : none);
}

__global__ void advance_time_kernel(double *ctl)
{
    if (ctl[0] > 0.0) { ctl[1] += ctl[0]; ctl[5] += 1.0; }
}

void launch_advance_time(double *ctl, cudaStream_t s) { advance_time_kernel<<<1, 1, 0, s>>>(ctl); }

// ------------------------------------------------------------------ start-up and diagnostics
// interpolateB_center2face, Weno5_2D.jl:1669-1682 (public i = 2..Nx-2, j = 2..Ny-2; 0 elsewhere)
__global__ void center2face_kernel(const double *__restrict__ Bx, const double *__restrict__ By,
                                   double *__restrict__ bxb, double *__restrict__ byb, int Nx, int Ny, int pitch)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > Nx || j > Ny) return;
    const size_t g = (size_t)i + (size_t)pitch * j;
    double fx = 0.0, fy = 0.0;
    if (i >= 2 && i <= Nx - 2 && j >= 2 && j <= Ny - 2) {
        fx = (-Bx[g - 1] + 9 * Bx[g] + 9 * Bx[g + 1] - Bx[g + 2]) / 16;
        fy = (-By[g - pitch] + 9 * By[g] + 9 * By[g + pitch] - By[g + 2 * (size_t)pitch]) / 16;
    }
    bxb[g] = fx;
    byb[g] = fy;
}

void launch_center2face(const double *Bx, const double *By, double *bxb, double *byb, const Geom &g,
                        cudaStream_t s)
{
    dim3 block(64, 4), grid((g.Nx + 63) / 64, (g.Ny + 3) / 4);
    center2face_kernel<<<grid, block, 0, s>>>(Bx, By, bxb, byb, g.Nx, g.Ny, g.pitch);
}

// compute_divB, Weno5_2D.jl:1986-1999 (public i = 3..Nx-3, j = 3..Ny-3)
__global__ void divb_kernel(const double *__restrict__ bxb, const double *__restrict__ byb,
                            double *__restrict__ divb, int Nx, int Ny, int pitch, double dx, double dy)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > Nx || j > Ny) return;
    const size_t g = (size_t)i + (size_t)pitch * j;
    double d = 0.0;
    if (i >= 3 && i <= Nx - 3 && j >= 3 && j <= Ny - 3)
        d = (bxb[g] - bxb[g - 1]) / dx + (byb[g] - byb[g - pitch]) / dy;
    divb[g] = d;
}

void launch_divb(const double *bxb, const double *byb, double *divb, const Geom &g, double dx, double dy,
                 cudaStream_t s)
{
    dim3 block(64, 4), grid((g.Nx + 63) / 64, (g.Ny + 3) / 4);
    divb_kernel<<<grid, block, 0, s>>>(bxb, byb, divb, g.Nx, g.Ny, g.pitch, dx, dy);
}

// interior sums per plane (deterministic: one block per plane, fixed reduction tree)
struct TotArgs { const double *pl[11]; int NXI, NYI, pitch; double *out; };

__global__ void totals_kernel(const TotArgs a)
{
    const double *f = a.pl[blockIdx.x];
    const int nxi = a.NXI - 2 * G, nyi = a.NYI > 1 ? a.NYI - 2 * G : 1;
    const long n = (long)nxi * nyi;
    double acc = 0.0;
    for (long e = threadIdx.x; e < n; e += blockDim.x) {
        const int i = (int)(e % nxi) + G, j = a.NYI > 1 ? (int)(e / nxi) + G : 0;
        acc += f[(size_t)i + (size_t)a.pitch * j];
    }
    __shared__ double sh[1024];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) a.out[blockIdx.x] = sh[0];
}

void launch_totals(const double *const planes[], int nplanes, const Geom &g, double *sums, cudaStream_t s)
{
    TotArgs a;
    for (int k = 0; k < nplanes; ++k) a.pl[k] = planes[k];
    a.NXI = g.NXI; a.NYI = g.NYI; a.pitch = g.pitch; a.out = sums;
    totals_kernel<<<nplanes, 1024, 0, s>>>(a);
}

// ------------------------------------------------------------------ snapshot hook (Arr2Img, JColorMaps.jl:92-146)
// Order-preserving map double -> unsigned 64-bit (all finite values and infinities), so that grid-wide
// minima / maxima are integer atomics.
__device__ __forceinline__ unsigned long long ord_key(double v)
{
    const long long b = __double_as_longlong(v);
    return b >= 0 ? (unsigned long long)b ^ 0x8000000000000000ull : ~(unsigned long long)b;
}
__device__ __forceinline__ double ord_val(unsigned long long k)
{
    return __longlong_as_double((k & 0x8000000000000000ull) ? (long long)(k ^ 0x8000000000000000ull) : (long long)~k);
}

// keys[0] = min over v > 0, keys[1] = min over finite v, keys[2] = max over finite v (public Nx x Ny region)
__global__ void img_range_kernel(const double *__restrict__ f, int Nx, int Ny, int pitch, int row0, unsigned long long *keys)
{
    unsigned long long kp = ~0ull, kmin = ~0ull, kmax = 0ull;
    const int n = Nx * Ny;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int i = e % Nx + 1, j = e / Nx + row0;
        const double v = f[(size_t)i + (size_t)pitch * j];
        const unsigned long long k = ord_key(v);
        if (v > 0.0) kp = min(kp, k);
        if (isfinite(v)) { kmin = min(kmin, k); kmax = max(kmax, k); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        kp = min(kp, __shfl_xor_sync(0xffffffffu, kp, o));
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&keys[0], kp);
        atomicMin(&keys[1], kmin);
        atomicMax(&keys[2], kmax);
    }
}

// colour index (1-based, Int16) of every pixel of the TRANSPOSED field: out[(j-1) + Ny (i-1)].
// rng[0..1] receive the range that was used (before the log10).
__global__ void img_index_kernel(const double *__restrict__ f, int Nx, int Ny, int pitch, int row0, double lo, double hi,
                                 int log_scale, int ncolours, const unsigned long long *keys, short *out, double *rng)
{
    double minpos = 0.0;
    if (log_scale) minpos = ord_val(keys[0]);
    if (lo == 0.0 && hi == 0.0) {                                   // automatic range over the finite pixels
        if (log_scale) {                                             // after non-positive pixels became minpos
            const double mx = ord_val(keys[2]);
            lo = minpos;
            hi = mx > 0.0 ? mx : minpos;
        } else {
            lo = ord_val(keys[1]);
            hi = ord_val(keys[2]);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { rng[0] = lo; rng[1] = hi; }
    if (log_scale) { lo = log10(lo); hi = log10(hi); }
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Nx * Ny) return;
    const int i = e % Nx + 1, j = e / Nx + row0;
    double v = f[(size_t)i + (size_t)pitch * j];
    if (log_scale) {
        if (v <= 0.0) v = minpos;
        v = log10(v);
    }
    if (isnan(v)) v = lo;
    double r = rint((v - lo) / (hi - lo) * (double)(ncolours - 1)) + 1.0;
    r = fmin(fmax(r, 1.0), (double)ncolours);                       // also absorbs +-inf; NaN (0/0 range) -> 1
    if (isnan(r)) r = 1.0;
    out[(size_t)(j - row0) + (size_t)Ny * (i - 1)] = (short)r;
}

void launch_arr2img(const double *field, const Geom &g, double lo, double hi, int log_scale, int ncolours,
                    unsigned long long *keys, short *out, double *rng, cudaStream_t s)
{
    const unsigned long long init[3] = {~0ull, ~0ull, 0ull};
    cudaMemcpyAsync(keys, init, sizeof(init), cudaMemcpyHostToDevice, s);
    const int n = g.Nx * g.Ny, row0 = g.is1d ? 0 : 1;
    img_range_kernel<<<min((n + 255) / 256, 1184), 256, 0, s>>>(field, g.Nx, g.Ny, g.pitch, row0, keys);
    img_index_kernel<<<(n + 255) / 256, 256, 0, s>>>(field, g.Nx, g.Ny, g.pitch, row0, lo, hi, log_scale, ncolours, keys, out, rng);
}

}  // namespace w5
