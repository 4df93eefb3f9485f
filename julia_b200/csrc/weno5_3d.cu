// 3-D extension of libweno5mhd (include/weno5mhd.h "3-D extension"): kernels and C ABI.
//
// The arithmetic and the stage sequence live in weno5_3d_body.cuh (host/device, so that the same code is checked
// against the CPU oracle without a GPU); this file wraps the bodies in sm_100a kernels, owns the device memory of a
// handle and exports the entry points.
//
// Per RK stage: sweep3<0>, sweep3<1>, sweep3<2> (each: cell phase -> face phase -> flux difference inside one CTA
// tile staged in shared memory; x is the fastest thread coordinate in all three so every global access is coalesced;
// the y and z sweeps walk the volume with strides Nx and Nx*Ny instead of transposing it, rotate_state
// Weno5_2D.jl:1897 does not exist) -> emf3 -> face_update3 -> update3 (RK combination, face->centre interpolation,
// E2S/S2E) -> cell_bc3 (-> face_bc3 after stage 4).  Per step additionally vmax3 -> dt3 -> advance_time3.
// The step loop of weno5_advance_3d never synchronises with the host (dt, t and the maxima live in device memory).
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/weno5mhd.h"
#include "weno5_3d_body.cuh"

namespace w5 { int api_fail(int code, const char *msg); }      // weno5_api.cu: sets the thread-local error text

using namespace w5;
using namespace w5::d3;

static int fail3(int code, const char *fmt, ...)
{
    char buf[480];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return w5::api_fail(code, buf);
}

#define CU3(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail3(WENO5_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

namespace {

struct DevGuard3 {
    int prev = -1;
    cudaError_t set(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        return cudaSetDevice(dev);
    }
    ~DevGuard3() { if (prev >= 0) cudaSetDevice(prev); }
};

// ------------------------------------------------------------------ kernels
template <int DIR>
__global__ void __launch_bounds__(NT3, W5_3D_MINB) sweep3_kernel(const SweepArgs a)
{
    extern __shared__ double smem3[];
    if (a.dtp[0] == 0.0) return;                           // finished (tmax reached): nothing to do
    sweep_cells<DIR>(a, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, smem3);
    __syncthreads();
    sweep_faces<DIR>(a, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, smem3);
    __syncthreads();
    sweep_update<DIR>(a, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, smem3);
}

// y / z sweep with one thread per line (weno5_3d_body.cuh, sweep_march): no barrier, thread-private ring
template <int DIR>
__global__ void __launch_bounds__(MarchGeo::NT, 1) sweep3_march_kernel(const SweepArgs a, int nseg)
{
    extern __shared__ double smem3[];
    if (a.dtp[0] == 0.0) return;
    sweep_march<DIR>(a, nseg, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, smem3);
}

// x sweep with LPL lanes per line (weno5_3d_body.cuh, xwarp_*): warp-private ring, __syncwarp between the phases
template <int LPL>
__global__ void __launch_bounds__(XWarpGeo<LPL>::NT, 1) sweep3_xwarp_kernel(const SweepArgs a, int nseg)
{
    extern __shared__ double smem3[];
    if (a.dtp[0] == 0.0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *wsm = smem3 + (size_t)warp * XWarpGeo<LPL>::PER_WARP;
    const XWarpPos<LPL> P(a.d, nseg, blockIdx.x, blockIdx.y, blockIdx.z, warp, lane);
    for (int chunk = 0; chunk < P.nchunks; ++chunk) {
        xwarp_cells<LPL>(a, nseg, blockIdx.x, blockIdx.y, blockIdx.z, warp, lane, chunk, wsm);
        __syncwarp();
        xwarp_faces<LPL>(a, nseg, blockIdx.x, blockIdx.y, blockIdx.z, warp, lane, chunk, wsm);
        __syncwarp();
        xwarp_update<LPL>(a, nseg, blockIdx.x, blockIdx.y, blockIdx.z, warp, lane, chunk, wsm);
    }
}

template <int LPL> void launch_xwarp(const SweepArgs &a, int nseg, cudaStream_t s)
{
    int g[3];
    XWarpGeo<LPL>::grid(a.d, nseg, g);
    sweep3_xwarp_kernel<LPL><<<dim3(g[0], g[1], g[2]), XWarpGeo<LPL>::NT, XWarpGeo<LPL>::SMEM, s>>>(a, nseg);
}

// elementwise kernels: one thread per cell of the box [lo, N-lo) in every axis; block (64, 2, 2)
#define W5_CELL_IJK(d, lo)                                                     \
    const int i = blockIdx.x * blockDim.x + threadIdx.x + (lo);               \
    const int j = blockIdx.y * blockDim.y + threadIdx.y + (lo);               \
    const int k = blockIdx.z * blockDim.z + threadIdx.z + (lo);               \
    const bool inside = i < (d).Nx - (lo) && j < (d).Ny - (lo) && k < (d).Nz - (lo)

__global__ void emf3_kernel(const CtArgs3 a)
{
    W5_CELL_IJK(a.d, 0);
    if (!inside || a.dtp[0] == 0.0) return;
    emf_cell(a, i, j, k);
}

__global__ void face_update3_kernel(const CtArgs3 a)
{
    W5_CELL_IJK(a.d, 0);
    const double dt = a.dtp[0];
    if (!inside || dt == 0.0) return;
    face_update_cell(a, dt, i, j, k);
}

__global__ void ct3_kernel(const CtArgs3 a)
{
    W5_CELL_IJK(a.d, 0);
    const double dt = a.dtp[0];
    if (!inside || dt == 0.0) return;
    ct_cell(a, dt, i, j, k);
}

__global__ void update3_kernel(const UpdArgs3 a)
{
    W5_CELL_IJK(a.d, NB3);
    const double dt = a.dtp[0];
    if (!inside || dt == 0.0) return;
    update_cell(a, dt, i, j, k);
}

struct Vols { double *v[9]; int n; };

__global__ void cell_bc3_kernel(const Vols p, const Dims d, const double *dtp)
{
    W5_CELL_IJK(d, 0);
    if (!inside || (dtp && dtp[0] == 0.0)) return;
    for (int c = 0; c < p.n; ++c) cell_bc(p.v[c], d, i, j, k);
}

__global__ void face_bc3_kernel(const Vols p, const Dims d, const double *dtp)
{
    W5_CELL_IJK(d, 0);
    if (!inside || (dtp && dtp[0] == 0.0)) return;
    for (int c = 0; c < p.n; ++c) face_bc_cell(p.v[c], d, i, j, k);
}

struct CVols { const double *v[9]; };

// grid-wide maxima by integer atomicMax on the bit patterns of non-negative doubles (NaN -> all ones)
__global__ void vmax3_kernel(const CVols q, const Dims d, double gam, unsigned long long *out)
{
    W5_CELL_IJK(d, NB3);
    double v[3] = {0.0, 0.0, 0.0};
    if (inside) vmax_cell(q.v, d, gam, i, j, k, v);
    unsigned long long b[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        b[m] = (unsigned long long)__double_as_longlong(v[m]);
        if (v[m] != v[m] || v[m] < 0.0) b[m] = ~0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b[m] = max(b[m], __shfl_xor_sync(0xffffffffu, b[m], o));
    }
    __shared__ unsigned long long sm[3][8];
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    if ((tid & 31) == 0) { sm[0][tid >> 5] = b[0]; sm[1][tid >> 5] = b[1]; sm[2][tid >> 5] = b[2]; }
    __syncthreads();
    if (tid < 3) {
        unsigned long long r = 0ull;
        const int nw = (blockDim.x * blockDim.y * blockDim.z + 31) / 32;
        for (int w = 0; w < nw; ++w) r = max(r, sm[tid][w]);
        atomicMax(out + tid, r);
    }
}

__global__ void dt3_kernel(unsigned long long *vb, double *ctl, double dx, double dy, double dz, double cfl)
{
    double vm[3];
    for (int m = 0; m < 3; ++m) { vm[m] = __longlong_as_double((long long)vb[m]); vb[m] = 0ull; }
    dt_from_vmax(vm, ctl, dx, dy, dz, cfl);
}

__global__ void advance_time3_kernel(double *ctl)
{
    if (ctl[0] > 0.0) { ctl[1] += ctl[0]; ctl[6] += 1.0; }
}

__global__ void divb3_kernel(const double *bx, const double *by, const double *bz, double *out, const Dims d, double ix,
                             double iy, double iz)
{
    W5_CELL_IJK(d, 0);
    if (inside) divb_cell(bx, by, bz, out, d, ix, iy, iz, i, j, k);
}

__global__ void center2face3_kernel(const double *Bx, const double *By, const double *Bz, double *bx, double *by,
                                    double *bz, const Dims d)
{
    W5_CELL_IJK(d, 0);
    if (inside) center2face_cell(Bx, By, Bz, bx, by, bz, d, i, j, k);
}

struct TotVols { const double *v[12]; };

// interior sums of 12 volumes; one block per (j,k) row group, double atomicAdd of block sums
__global__ void totals3_kernel(const TotVols p, const Dims d, double *sums)
{
    W5_CELL_IJK(d, NB3);
    __shared__ double sm[12][8];
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    for (int c = 0; c < 12; ++c) {
        double x = inside ? p.v[c][vidx(d, i, j, k)] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((tid & 31) == 0) sm[c][tid >> 5] = x;
    }
    __syncthreads();
    if (tid < 12) {
        double r = 0.0;
        const int nw = (blockDim.x * blockDim.y * blockDim.z + 31) / 32;
        for (int w = 0; w < nw; ++w) r += sm[tid][w];
        atomicAdd(sums + tid, r);
    }
}

dim3 cell_grid(const Dims &d, int lo, dim3 block)
{
    return dim3((d.Nx - 2 * lo + block.x - 1) / block.x, (d.Ny - 2 * lo + block.y - 1) / block.y,
                (d.Nz - 2 * lo + block.z - 1) / block.z);
}
const dim3 kBlock(64, 2, 2);

}  // namespace

// ------------------------------------------------------------------ handle
struct weno5_ctx3d {
    weno5_params3d P;
    int device = 0;
    Bufs3 b;
    size_t vol = 0;                 // doubles per volume
    cudaStream_t stream = nullptr;
    double *pool = nullptr;
    double *scratch = nullptr;      // one volume (divB)
    double *sums = nullptr;         // 12 doubles
    unsigned long long *vbits = nullptr;
    bool have_state = false;
    // sweep kernels: x by sweep3_xwarp_kernel (a warp per line), y and z by sweep3_march_kernel (a thread per line).
    // WENO5_3D_SWEEP (tuning / cross-checks): tile = the CTA-tile kernel in all directions, xtile = for x only,
    // xmarch = x by the thread-per-line kernel (uncoalesced)
    bool fused_ct = false;          // WENO5_3D_CT=fused: EMFs and face update in one kernel (measured slower: 25.1 vs 24.3 ms/step at 256^3)
    bool xwarp = true;
    int lpl = 32;                   // lanes per line of the x sweep (8, 16, 32)
    bool march[3] = {false, true, true};
    int nseg[3] = {1, 1, 1};        // segments per line of the marching sweeps
    unsigned long long launches = 0;
    // sweep timing (weno5_profile_3d)
    bool profiling = false;
    std::vector<cudaEvent_t> ev;    // pairs
    size_t ev_used = 0;
};

namespace {

struct DevExec {
    weno5_ctx3d *c;
    bool gate = true;               // kernels return at once when the device time step is 0 (tmax reached)
    template <int DIR> void sweep(const SweepArgs &a)
    {
        int g[3];
        SweepGeo<DIR>::grid(a.d, g);
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (c->profiling && c->ev_used < 2 * 4096) {        // at most 4096 timed launches between two reads
            if (c->ev_used + 2 > c->ev.size()) {
                cudaEvent_t x, y;
                cudaEventCreate(&x); cudaEventCreate(&y);
                c->ev.push_back(x); c->ev.push_back(y);
            }
            e0 = c->ev[c->ev_used]; e1 = c->ev[c->ev_used + 1];
            c->ev_used += 2;
            cudaEventRecord(e0, c->stream);
        }
        if (DIR == 0 && c->xwarp) {
            if (c->lpl == 8) launch_xwarp<8>(a, c->nseg[0], c->stream);
            else if (c->lpl == 16) launch_xwarp<16>(a, c->nseg[0], c->stream);
            else launch_xwarp<32>(a, c->nseg[0], c->stream);
        } else if (c->march[DIR]) {
            MarchGeo::grid(a.d, DIR, c->nseg[DIR], g);
            sweep3_march_kernel<DIR><<<dim3(g[0], g[1], g[2]), MarchGeo::NT, MarchGeo::SMEM, c->stream>>>(a, c->nseg[DIR]);
        } else {
            sweep3_kernel<DIR><<<dim3(g[0], g[1], g[2]), NT3, SweepGeo<DIR>::SMEM, c->stream>>>(a);
        }
        if (e1) cudaEventRecord(e1, c->stream);
        c->launches++;
    }
    void ct(const CtArgs3 &a)
    {
        if (c->fused_ct) {
            ct3_kernel<<<cell_grid(a.d, 0, kBlock), kBlock, 0, c->stream>>>(a);
            c->launches++;
        } else {                                           // default: EMF volumes written and read back
            emf3_kernel<<<cell_grid(a.d, 0, kBlock), kBlock, 0, c->stream>>>(a);
            face_update3_kernel<<<cell_grid(a.d, 0, kBlock), kBlock, 0, c->stream>>>(a);
            c->launches += 2;
        }
    }
    void update(const UpdArgs3 &a) { update3_kernel<<<cell_grid(a.d, NB3, kBlock), kBlock, 0, c->stream>>>(a); c->launches++; }
    void cell_bc(double *const v[], int n, const Dims &d)
    {
        Vols p;
        p.n = n;
        for (int k = 0; k < n; ++k) p.v[k] = v[k];
        cell_bc3_kernel<<<cell_grid(d, 0, kBlock), kBlock, 0, c->stream>>>(p, d, gate ? c->b.ctl : nullptr);
        c->launches++;
    }
    void face_bc(double *const v[3], const Dims &d)
    {
        Vols p;
        p.n = 3;
        for (int k = 0; k < 3; ++k) p.v[k] = v[k];
        face_bc3_kernel<<<cell_grid(d, 0, kBlock), kBlock, 0, c->stream>>>(p, d, gate ? c->b.ctl : nullptr);
        c->launches++;
    }
};

void enqueue_timestep3(weno5_ctx3d *c)
{
    CVols q;
    for (int k = 0; k < 9; ++k) q.v[k] = c->b.Q[0][k];
    vmax3_kernel<<<cell_grid(c->b.d, NB3, kBlock), kBlock, 0, c->stream>>>(q, c->b.d, c->b.gam, c->vbits);
    dt3_kernel<<<1, 1, 0, c->stream>>>(c->vbits, c->b.ctl, c->b.dx, c->b.dy, c->b.dz, c->b.cfl);
    c->launches += 2;
}

int check_bad(weno5_ctx3d *c, const double ctl[8])
{
    if (ctl[7] != 0.0) {
        cudaMemsetAsync(c->b.ctl + 7, 0, sizeof(double), c->stream);
        return fail3(WENO5_ERR_NUMERIC, "non-finite wave speed or time step (vmax = %g, %g, %g)", ctl[2], ctl[3], ctl[4]);
    }
    return 0;
}

}  // namespace

// ------------------------------------------------------------------ C ABI
extern "C" int weno5_destroy_3d(weno5_ctx3d *c)
{
    if (!c) return 0;
    DevGuard3 dg;
    dg.set(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
    if (c->pool) cudaFree(c->pool);
    if (c->vbits) cudaFree(c->vbits);
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
    return 0;
}

extern "C" int weno5_create_3d(const weno5_params3d *P, int device, weno5_ctx3d **out)
{
    if (!P || !out) return fail3(WENO5_ERR_ARG, "null argument");
    *out = nullptr;
    if (P->nx < 4 || P->ny < 4 || P->nz < 4)
        return fail3(WENO5_ERR_ARG, "grid too small: %d x %d x %d (need >= 4 cells per axis)", P->nx, P->ny, P->nz);
    if (!(P->dx > 0) || !(P->dy > 0) || !(P->dz > 0) || !(P->gamma > 1.0)) return fail3(WENO5_ERR_ARG, "bad dx/dy/dz/gamma");
    for (int bc : {P->bc_x, P->bc_y, P->bc_z})
        if (bc != WENO5_BC_OUTFLOW && bc != WENO5_BC_PERIODIC) return fail3(WENO5_ERR_ARG, "unknown boundary condition %d", bc);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail3(WENO5_ERR_CUDA, "no CUDA device: libweno5mhd has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail3(WENO5_ERR_ARG, "device %d out of range (0..%d)", device, ndev - 1);
    cudaDeviceProp prop;
    CU3(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail3(WENO5_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    DevGuard3 dg;
    CU3(dg.set(device));

    weno5_ctx3d *c = new weno5_ctx3d();
    struct Cleanup { weno5_ctx3d *c; ~Cleanup() { if (c) weno5_destroy_3d(c); } } cleanup{c};
    c->P = *P;
    c->device = device;
    Bufs3 &b = c->b;
    b.d.Nx = P->nx + 2 * NB3; b.d.Ny = P->ny + 2 * NB3; b.d.Nz = P->nz + 2 * NB3;
    b.d.bc[0] = P->bc_x; b.d.bc[1] = P->bc_y; b.d.bc[2] = P->bc_z;
    b.dx = P->dx; b.dy = P->dy; b.dz = P->dz; b.gam = P->gamma; b.eps = P->eps; b.cfl = P->cfl;
    b.roundtrip = P->es_roundtrip;
    const size_t n = (size_t)b.d.Nx * b.d.Ny * b.d.Nz;
    c->vol = (n + 31) / 32 * 32;                           // volumes start on 256-byte boundaries
    // 4 states x 9 + 4 x 3 face fields + 5 rhs + 6 transport fluxes + 3 EMFs + scratch = 63 volumes, + ctl, sums
    const size_t nvol = 4 * 9 + 4 * 3 + 5 + 6 + 3 + 1;
    CU3(cudaMalloc(&c->pool, (nvol * c->vol + 32) * sizeof(double)));
    CU3(cudaMemset(c->pool, 0, (nvol * c->vol + 32) * sizeof(double)));
    double *cur = c->pool;
    auto take = [&]() { double *p = cur; cur += c->vol; return p; };
    // the 9 components of Q[0] and the 3 of B[0] are consecutive: upload / download are single copies per array
    for (int s = 0; s < 4; ++s) {
        for (int k = 0; k < 9; ++k) b.Q[s][k] = take();
        for (int k = 0; k < 3; ++k) b.B[s][k] = take();
    }
    for (int k = 0; k < 5; ++k) b.rhs[k] = take();
    for (int k = 0; k < 6; ++k) b.fl[k] = take();
    for (int k = 0; k < 3; ++k) b.O[k] = take();
    c->scratch = take();
    b.ctl = cur; c->sums = cur + 16;
    CU3(cudaMalloc(&c->vbits, 4 * sizeof(unsigned long long)));
    CU3(cudaMemset(c->vbits, 0, 4 * sizeof(unsigned long long)));
    CU3(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    // function attributes belong to the device: set them for every handle
    CU3(cudaFuncSetAttribute(sweep3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepGeo<0>::SMEM));
    CU3(cudaFuncSetAttribute(sweep3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepGeo<1>::SMEM));
    CU3(cudaFuncSetAttribute(sweep3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepGeo<2>::SMEM));
    CU3(cudaFuncSetAttribute(sweep3_xwarp_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XWarpGeo<8>::SMEM));
    CU3(cudaFuncSetAttribute(sweep3_xwarp_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XWarpGeo<16>::SMEM));
    CU3(cudaFuncSetAttribute(sweep3_xwarp_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XWarpGeo<32>::SMEM));
    CU3(cudaFuncSetAttribute(sweep3_march_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MarchGeo::SMEM));
    CU3(cudaFuncSetAttribute(sweep3_march_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MarchGeo::SMEM));
    CU3(cudaFuncSetAttribute(sweep3_march_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MarchGeo::SMEM));
    {
        const char *e = getenv("WENO5_3D_SWEEP");
        const bool tile = e && strcmp(e, "tile") == 0, xtile = e && strcmp(e, "xtile") == 0;
        const bool xmarch = e && strcmp(e, "xmarch") == 0;
        c->xwarp = !(tile || xtile || xmarch);
        c->march[0] = xmarch;
        c->march[1] = c->march[2] = !tile;
        const char *ce = getenv("WENO5_3D_CT");
        c->fused_ct = ce && strcmp(ce, "fused") == 0;
        const char *ns = getenv("WENO5_3D_NSEG");                         // tuning knob: segments per line
        for (int dir = 0; dir < 3; ++dir) {
            c->nseg[dir] = ns ? atoi(ns) : MarchGeo::segments(b.d, dir, prop.multiProcessorCount);
            if (c->nseg[dir] < 1) c->nseg[dir] = 1;
        }
        if (c->xwarp) {
            int nsx, lpl;
            xwarp_plan(b.d, prop.multiProcessorCount, nsx, lpl);
            if (!ns) c->nseg[0] = nsx;
            const char *le = getenv("WENO5_3D_LPL");                      // tuning knob: lanes per line of the x sweep
            c->lpl = le ? atoi(le) : lpl;
            if (c->lpl != 8 && c->lpl != 16) c->lpl = 32;
        }
    }
    cleanup.c = nullptr;
    *out = c;
    return WENO5_OK;
}

#define ENTER3(c)                                            \
    if (!(c)) return fail3(WENO5_ERR_ARG, "null handle");    \
    DevGuard3 dg_;                                           \
    CU3(dg_.set((c)->device))

static int copy_vols(weno5_ctx3d *c, double *const dev[], const double *host_in, double *host_out, int n)
{
    const size_t N = (size_t)c->b.d.Nx * c->b.d.Ny * c->b.d.Nz;
    for (int k = 0; k < n; ++k) {
        if (host_in) CU3(cudaMemcpyAsync(dev[k], host_in + k * N, N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        if (host_out) CU3(cudaMemcpyAsync(host_out + k * N, dev[k], N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    return 0;
}

extern "C" int weno5_upload_3d(weno5_ctx3d *c, const double *q, const double *bxb, const double *byb, const double *bzb)
{
    ENTER3(c);
    if (!q) return fail3(WENO5_ERR_ARG, "q is null");
    if ((bxb != nullptr) != (byb != nullptr) || (bxb != nullptr) != (bzb != nullptr))
        return fail3(WENO5_ERR_ARG, "pass all three face fields or none");
    Bufs3 &b = c->b;
    if (int r = copy_vols(c, b.Q[0], q, nullptr, 9)) return r;
    DevExec ex{c, false};
    ex.cell_bc(b.Q[0], 9, b.d);                            // boundaries!(q), driver line Weno5_2D.jl:39
    if (bxb) {
        const double *h[3] = {bxb, byb, bzb};
        for (int m = 0; m < 3; ++m)
            if (int r = copy_vols(c, &b.B[0][m], h[m], nullptr, 1)) return r;
    } else {
        center2face3_kernel<<<cell_grid(b.d, 0, kBlock), kBlock, 0, c->stream>>>(b.Q[0][4], b.Q[0][5], b.Q[0][6], b.B[0][0],
                                                                                b.B[0][1], b.B[0][2], b.d);
        c->launches++;
    }
    ex.face_bc(b.B[0], b.d);                               // the face-array rules of boundaries!
    CU3(cudaMemsetAsync(b.ctl, 0, 8 * sizeof(double), c->stream));
    CU3(cudaStreamSynchronize(c->stream));
    CU3(cudaGetLastError());
    c->have_state = true;
    return WENO5_OK;
}

extern "C" int weno5_download_3d(weno5_ctx3d *c, double *q, double *bxb, double *byb, double *bzb)
{
    ENTER3(c);
    if (!c->have_state) return fail3(WENO5_ERR_STATE, "no state uploaded");
    if (q) {
        if (int r = copy_vols(c, c->b.Q[0], nullptr, q, 9)) return r;
    }
    double *h[3] = {bxb, byb, bzb};
    for (int m = 0; m < 3; ++m) {
        if (!h[m]) continue;
        if (int r = copy_vols(c, &c->b.B[0][m], nullptr, h[m], 1)) return r;
    }
    CU3(cudaStreamSynchronize(c->stream));
    return WENO5_OK;
}

extern "C" int weno5_timestep_3d(weno5_ctx3d *c, double *dt, double vmax[3])
{
    ENTER3(c);
    if (!c->have_state) return fail3(WENO5_ERR_STATE, "no state uploaded");
    CU3(cudaMemsetAsync(c->b.ctl + 5, 0, sizeof(double), c->stream));      // a plain query: no tmax logic
    enqueue_timestep3(c);
    double ctl[8];
    CU3(cudaMemcpyAsync(ctl, c->b.ctl, sizeof(ctl), cudaMemcpyDeviceToHost, c->stream));
    CU3(cudaStreamSynchronize(c->stream));
    CU3(cudaGetLastError());
    if (int r = check_bad(c, ctl)) return r;
    if (dt) *dt = ctl[0];
    if (vmax) { vmax[0] = ctl[2]; vmax[1] = ctl[3]; vmax[2] = ctl[4]; }
    return WENO5_OK;
}

extern "C" int weno5_rk4_step_3d(weno5_ctx3d *c, double dt)
{
    ENTER3(c);
    if (!c->have_state) return fail3(WENO5_ERR_STATE, "no state uploaded");
    if (!(dt > 0.0)) return fail3(WENO5_ERR_ARG, "dt must be positive");
    CU3(cudaMemcpyAsync(c->b.ctl, &dt, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    DevExec ex{c};
    enqueue_step3(c->b, ex);
    CU3(cudaStreamSynchronize(c->stream));
    CU3(cudaGetLastError());
    return WENO5_OK;
}

extern "C" int weno5_advance_3d(weno5_ctx3d *c, int nsteps, double tmax, double *t, int *nsteps_done)
{
    ENTER3(c);
    if (!c->have_state) return fail3(WENO5_ERR_STATE, "no state uploaded");
    if (nsteps < 0) return fail3(WENO5_ERR_ARG, "nsteps < 0");
    if (tmax <= 0.0 && nsteps > 1000000) return fail3(WENO5_ERR_ARG, "nsteps without tmax is limited to 1e6 per call");
    double head[8] = {0.0, t ? *t : 0.0, 0.0, 0.0, 0.0, tmax > 0.0 ? tmax : 0.0, 0.0, 0.0};
    CU3(cudaMemcpyAsync(c->b.ctl, head, sizeof(head), cudaMemcpyHostToDevice, c->stream));
    DevExec ex{c};
    double ctl[8];
    for (int s = 0; s < nsteps; ++s) {
        enqueue_timestep3(c);
        enqueue_step3(c->b, ex);
        advance_time3_kernel<<<1, 1, 0, c->stream>>>(c->b.ctl);
        c->launches++;
        if ((s & 63) == 63) {                              // bounded launch queue; stop early once tmax is reached
            CU3(cudaMemcpyAsync(ctl, c->b.ctl, sizeof(ctl), cudaMemcpyDeviceToHost, c->stream));
            CU3(cudaStreamSynchronize(c->stream));
            if (ctl[7] != 0.0 || (tmax > 0.0 && ctl[1] >= tmax)) break;
        }
    }
    CU3(cudaMemcpyAsync(ctl, c->b.ctl, sizeof(ctl), cudaMemcpyDeviceToHost, c->stream));
    CU3(cudaStreamSynchronize(c->stream));
    CU3(cudaGetLastError());
    if (t) *t = ctl[1];
    if (nsteps_done) *nsteps_done = (int)ctl[6];
    return check_bad(c, ctl);
}

extern "C" int weno5_divb_3d(weno5_ctx3d *c, double *divb)
{
    ENTER3(c);
    if (!c->have_state) return fail3(WENO5_ERR_STATE, "no state uploaded");
    if (!divb) return fail3(WENO5_ERR_ARG, "divb is null");
    const Bufs3 &b = c->b;
    divb3_kernel<<<cell_grid(b.d, 0, kBlock), kBlock, 0, c->stream>>>(b.B[0][0], b.B[0][1], b.B[0][2], c->scratch, b.d,
                                                                     1.0 / b.dx, 1.0 / b.dy, 1.0 / b.dz);
    c->launches++;
    double *dev[1] = {c->scratch};
    if (int r = copy_vols(c, dev, nullptr, divb, 1)) return r;
    CU3(cudaStreamSynchronize(c->stream));
    CU3(cudaGetLastError());
    return WENO5_OK;
}

extern "C" int weno5_totals_3d(weno5_ctx3d *c, double sums[12])
{
    ENTER3(c);
    if (!c->have_state) return fail3(WENO5_ERR_STATE, "no state uploaded");
    if (!sums) return fail3(WENO5_ERR_ARG, "sums is null");
    TotVols p;
    for (int k = 0; k < 9; ++k) p.v[k] = c->b.Q[0][k];
    for (int k = 0; k < 3; ++k) p.v[9 + k] = c->b.B[0][k];
    CU3(cudaMemsetAsync(c->sums, 0, 12 * sizeof(double), c->stream));
    totals3_kernel<<<cell_grid(c->b.d, NB3, kBlock), kBlock, 0, c->stream>>>(p, c->b.d, c->sums);
    c->launches++;
    CU3(cudaMemcpyAsync(sums, c->sums, 12 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU3(cudaStreamSynchronize(c->stream));
    CU3(cudaGetLastError());
    return WENO5_OK;
}

extern "C" int weno5_classical_rk4_step_3d(const weno5_params3d *P, const double *q0, const double *bxb0, const double *byb0,
                                           const double *bzb0, double dt, double *q4, double *bxb4, double *byb4, double *bzb4)
{
    weno5_ctx3d *c = nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }     // the caller's current device, as the 2-D one-shot call
    if (int r = weno5_create_3d(P, dev, &c)) return r;
    int r = weno5_upload_3d(c, q0, bxb0, byb0, bzb0);
    if (!r) r = weno5_rk4_step_3d(c, dt);
    if (!r) r = weno5_download_3d(c, q4, bxb4, byb4, bzb4);
    weno5_destroy_3d(c);
    return r;
}

extern "C" unsigned long long weno5_launch_count_3d(const weno5_ctx3d *c) { return c ? c->launches : 0; }

extern "C" int weno5_profile_3d(weno5_ctx3d *c, int enable)
{
    ENTER3(c);
    c->profiling = enable != 0;
    c->ev_used = 0;
    return WENO5_OK;
}

extern "C" int weno5_profile_read_3d(weno5_ctx3d *c, double sweep_ms[3], unsigned long long *sweep_launches)
{
    ENTER3(c);
    CU3(cudaStreamSynchronize(c->stream));
    double ms[3] = {0.0, 0.0, 0.0};
    for (size_t k = 0; k + 1 < c->ev_used; k += 2) {       // launches are enqueued x, y, z, x, y, z, ...
        float x = 0.f;
        CU3(cudaEventElapsedTime(&x, c->ev[k], c->ev[k + 1]));
        ms[(k / 2) % 3] += x;
    }
    if (sweep_ms) { sweep_ms[0] = ms[0]; sweep_ms[1] = ms[1]; sweep_ms[2] = ms[2]; }
    if (sweep_launches) *sweep_launches = c->ev_used / 2;
    c->ev_used = 0;
    return WENO5_OK;
}
