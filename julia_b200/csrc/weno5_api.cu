// C ABI and host orchestration of libweno5mhd (see include/weno5mhd.h).
//
// One handle = one GPU = one y-slab.  All work of a handle is queued on its own stream; the
// time step, the simulation time and the wave-speed maxima live in device memory (`ctl`), so
// the step loop of weno5_advance() never synchronises with the host.
//
// Stage sequence (replaces classical_RK4_step, Weno5_2D.jl:125-317, E branch):
//   flux_x  : d fhat_x (6 fields) + fsy                       <- q_k
//   flux_y  : d fhat_y, fused RK update of rho,M,Bz,E + gsx   <- q_k, rhs, base
//   ct_faces: corner EMF, bxb/byb RK update                   <- fsy, gsx
//   ct_centers: Bx,By from the faces (4th order)
//   entropy : S = E2S(q) (and E = S2E(q,S) if es_roundtrip)   [every stage or last only]
//   bc_fill : ghosts of the 8(9) cell fields; NCCL exchange of 4 ghost rows on slab cuts
// RK storage: Q0 (step input, overwritten by the result in stage 4), QA/QB ping-pong,
// ACC = -q0 + q1 + 2 q2 accumulated while q1, q2 are read anyway.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/weno5mhd.h"
#include "weno5_kernels.cuh"

using namespace w5;

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// error text for the other translation units of the library (weno5_3d.cu)
namespace w5 { int api_fail(int code, const char *msg) { return fail(code, "%s", msg); } }

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(WENO5_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// Entry points run on the handle's device and give the caller's current device back on return (a host that
// shares the thread with CUDA.jl or torch keeps its own device selection).
struct DevGuard {
    int prev = -1;
    cudaError_t set(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        return cudaSetDevice(dev);
    }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ------------------------------------------------------------------ NCCL through dlopen
// The library has no link-time NCCL dependency: single-GPU users (and Julia) need none, and a
// host process that already loaded NCCL (e.g. torch) shares that copy.
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8, ncclUint64 = 5, ncclUint8 = 1 };
enum { ncclMax = 2, ncclMin = 3 };

struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load()
{
    if (g_nccl.h) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) { h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (h) break; }
    const char *env = getenv("WENO5_NCCL_LIB");
    if (!h && env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    for (const char *n : names) { if (h) break; h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); }
    if (!h) return fail(WENO5_ERR_NCCL, "cannot load libnccl.so.2 (set WENO5_NCCL_LIB): %s", dlerror());
#define SYM(field, name)                                                                  \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                            \
    if (!g_nccl.field) return fail(WENO5_ERR_NCCL, "libnccl lacks symbol %s", name);
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.h = h;
    return 0;
}

#define NC(call)                                                                                  \
    do {                                                                                          \
        int r_ = (call);                                                                          \
        if (r_ != ncclSuccess)                                                                    \
            return fail(WENO5_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?"); \
    } while (0)

// ------------------------------------------------------------------ handle
struct State {                 // 9 cell planes + 2 face planes
    double *q[NCELL];
    double *bx, *by;
};

struct weno5_ctx {
    weno5_params P;
    int device;
    int ny_local, j_offset;
    Geom g;
    cudaStream_t stream = nullptr;
    double *pool = nullptr;          // one allocation for every plane
    State Q0, QA, QB;
    double *acc[6], *accbx, *accby, *acc_by1d;
    double *rhs[7];
    double *fsy, *gsx;
    double *scratch;                 // one plane (divB)
    double *ctl = nullptr;           // device: dt, t, vxmax, vymax, tmax, steps, bad
    unsigned long long *vmax_bits = nullptr;
    unsigned long long *vmax_bits_img = nullptr;   // 3 keys of the snapshot range reduction
    unsigned long long *vmax_next = nullptr;       // maxima of the state stage 4 is writing (fused reduction, weno5_advance)
    bool fuse_vmax = false;                        // stage 4 also reduces the wave speeds of the new state
    // one captured step (timestep + 4 stages + clock) replayed by weno5_advance: launch-bound small grids
    // (512^2: ~25 launches of 5-40 us each) stop paying the per-launch CPU cost
    cudaGraphExec_t step_graph = nullptr;
    int graph_key = -1;
    unsigned long long graph_launches = 0;
    double *sums = nullptr;
    double *h_ctl = nullptr;         // pinned mirror
    int side[4];                     // SideBC of x-lo, x-hi, y-lo, y-hi for this slab
    bool uploaded = false;
    // multi-GPU: NCCL runs on its own stream so that the ghost-row exchange of stage k overlaps the
    // interior x sweep of stage k+1
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_main = nullptr, ev_comm = nullptr;
    bool xchg_pending = false;
    double *halo = nullptr;          // 4 message buffers: send up, send down, recv from up, recv from down
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    int up = -1, down = -1;          // neighbour ranks (-1: none)
    // halo push over peer memory (default; NCCL send/recv is the fallback): the neighbours' plane pools and flag
    // words mapped through CUDA IPC, a sequence number per exchange
    bool peer_ok = false;
    double *peer_pool_up = nullptr, *peer_pool_down = nullptr;
    unsigned long long *peer_flags_up = nullptr, *peer_flags_down = nullptr;
    int peer_ny_up = 0, peer_ny_down = 0;
    unsigned long long *sync_flags = nullptr;    // [0] last push from below, [1] from above, [2] block counter, [8..] all-reduce slots
    unsigned long long *share_slots[SHARE_MAX_RANKS] = {};   // every rank's all-reduce slot array (own included), or null
    std::vector<void *> ipc_opened;              // everything cudaIpcOpenMemHandle returned (closed in weno5_destroy)
    std::vector<unsigned long long *> peer_flag_maps;   // the flag words among them: [3] counts who maps that rank
    bool share_ok = false;                       // the dt all-reduce runs over peer memory inside dt_kernel
    unsigned long long share_seq = 0;
    unsigned long long xseq = 0, peer_wait_seq = 0;
    // instrumentation
    unsigned long long launches = 0;
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
    size_t ev_used = 0;
    double prof_ms = 0.0;
    unsigned long long prof_launches = 0;
    // launch decomposition
    int segx_len, nsegx, segy_len, nsegy;
    // TMA: one 3-D tensor map (x, y, plane) over the whole pool per sweep direction (box shapes differ)
    CUtensorMap tmap_x[4], tmap_y[4];        // boxes of 8, 6, 4, 2 planes
    bool have_tmap = false;
    size_t nplanes = 0;
    // dual-energy mode (weno5_set_es_switch): both branches are swept and ES_Switch selects per cell
    bool dual = false;
    bool dual_flags_only = false;    // both branches run, the state stays the energy branch's, the flags are the output
    double es_threshold = 1e-5;
    double *dual_pool = nullptr;
    double *candE[8], *candS[8];     // candidate states: rho,Mx,My,Mz,Bz,E|S,Bx,By
    double *accS[6];                 // RK accumulators of the S run (rho,Mx,My,Mz,Bz duplicate the E run's)
    double *fsyS, *gsxS, *tbx, *tby, *flag;
    unsigned long long switched = 0; // cells that took the E state in the last stage (diagnostic)
};

// cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time libcuda dependency
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_tensor_map(CUtensorMap *tm, double *pool, const Geom &g, size_t nplanes, const int box[2], int zbox)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) {
            cudaGetLastError();
            return false;
        }
        fn = (EncodeTiledFn)p;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)g.pitch, (cuuint64_t)g.NYI, (cuuint64_t)nplanes};
    const cuuint64_t strides[2] = {(cuuint64_t)g.pitch * sizeof(double), (cuuint64_t)g.plane * sizeof(double)};
    const cuuint32_t bx[3] = {(cuuint32_t)box[0], (cuuint32_t)box[1], (cuuint32_t)zbox};
    const cuuint32_t es[3] = {1, 1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, pool, dims, strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Cut every line of a sweep into segments (one CTA marches one segment of one strip of lines).
// Measured on B200 (scripts/wave_sweep.sh): the step time is flat between 4 and 12 waves of the 296
// resident CTAs and rises below 4 (tail) and for very short segments (prologue: 5 leading cells,
// one redundant face, pipeline start-up) -- wave quantisation itself costs little because the CTAs
// of a partial last wave have the FP64 pipe of their SM to themselves.
static void choose_segments(int nfaces, int NS, int strips, int &seg_len, int &nseg)
{
    const int chunks = (nfaces + NS - 1) / NS;
    static const int waves = getenv("WENO5_WAVES") ? atoi(getenv("WENO5_WAVES")) : 4;
    static const int minper = getenv("WENO5_MINPER") ? atoi(getenv("WENO5_MINPER")) : 4;
    const int resident = flux_resident_ctas();
    int want = (waves * resident + strips - 1) / strips;
    if (want < 1) want = 1;
    // among want .. 2*want segments per line take the count whose last wave of CTAs is fullest: cost model
    // (waves of 296 resident CTAs) x (chunks per segment + 1 for the prologue); at 4096^2 this picks 8 segments
    // for the y sweep (6.95 waves) instead of 5 (4.34 waves), measured 1.5 % of the step
    int best = want;
    long best_cost = -1;
    for (int k = want; k <= 2 * want; ++k) {
        const int perk = (chunks + k - 1) / k;
        if (perk < minper && k > want) break;
        const long nwaves = ((long)k * strips + resident - 1) / resident;
        const long cost = nwaves * (perk + 1);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = k; }
    }
    int per = (chunks + best - 1) / best;
    if (per < minper) per = minper;
    if (per > chunks) per = chunks;
    // every segment but the first starts one face early (its first cell needs both faces): with segments of
    // per * NS - 1 owned faces that extra face fits into `per` march steps instead of costing a step of its own
    // (measured on an 8192 x 1024 slab: 35 -> 34 steps per segment incl. prologue)
    seg_len = per < chunks ? per * NS - 1 : per * NS;
    nseg = (nfaces + seg_len - 1) / seg_len;
}

extern "C" const char *weno5_last_error(void) { return g_err; }
extern "C" int weno5_version(void) { return 100; }

extern "C" int weno5_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) ++ok;
    }
    return ok;
}

extern "C" int weno5_destroy(weno5_ctx *c);

static State carve_state(double *&cur, size_t plane)
{
    State s;
    for (int k = 0; k < NCELL; ++k) { s.q[k] = cur; cur += plane; }
    s.bx = cur; cur += plane;
    s.by = cur; cur += plane;
    return s;
}

// side_lo/side_hi >= 0 override the y sides derived from (j_offset, ny_local) -- used by the
// pipelined host step, whose slabs are cut at arbitrary rows
static int create_ctx(const weno5_params *P, int device, int ny_local, int j_offset, int side_lo, int side_hi, weno5_ctx **out)
{
    if (!P || !out) return fail(WENO5_ERR_ARG, "null argument");
    const bool is1d = P->ny == 1;
    // a periodic ghost layer (4 cells internally) must map onto interior cells; the reference's own
    // configuration is 16 x 4 cells (Weno5_2D.jl:11-12)
    if (P->nx < 4 || (!is1d && (P->ny < 4 || ny_local < 4)))
        return fail(WENO5_ERR_ARG, "grid too small: nx=%d ny=%d ny_local=%d (need >= 4, or ny = 1)", P->nx, P->ny, ny_local);
    if (is1d && (ny_local != 1 || j_offset != 0)) return fail(WENO5_ERR_ARG, "1-D grids cannot be decomposed");
    if (!is1d && side_lo < 0 && (j_offset < 0 || j_offset + ny_local > P->ny)) return fail(WENO5_ERR_ARG, "slab outside the grid");
    if (!(P->dx > 0) || (!is1d && !(P->dy > 0)) || !(P->gamma > 1.0)) return fail(WENO5_ERR_ARG, "bad dx/dy/gamma");
    for (int b : {P->bc_x, P->bc_y})
        if (b != WENO5_BC_OUTFLOW && b != WENO5_BC_PERIODIC) return fail(WENO5_ERR_ARG, "unknown boundary condition %d", b);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(WENO5_ERR_CUDA, "no CUDA device: libweno5mhd has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(WENO5_ERR_ARG, "device %d out of range (0..%d)", device, ndev - 1);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(WENO5_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    DevGuard dg_; CU(dg_.set(device));

    weno5_ctx *c = new weno5_ctx();
    // any failure from here on releases what has been allocated so far (weno5_destroy takes a half-built handle)
    struct Cleanup { weno5_ctx *c; ~Cleanup() { if (c) weno5_destroy(c); } } cleanup{c};
    c->P = *P;
    c->device = device;
    c->ny_local = ny_local;
    c->j_offset = j_offset;
    Geom &g = c->g;
    g.is1d = is1d;
    g.NXI = P->nx + 2 * G;
    g.NYI = is1d ? 1 : ny_local + 2 * G;
    g.pitch = (g.NXI + 15) / 16 * 16;
    g.Nx = P->nx + 2 * WENO5_NBND;
    g.Ny = is1d ? 1 : ny_local + 2 * WENO5_NBND;
    g.plane = (size_t)g.pitch * g.NYI;

    c->side[0] = c->side[1] = P->bc_x == WENO5_BC_PERIODIC ? SIDE_PERIODIC : SIDE_OUTFLOW;
    const int ybc = P->bc_y == WENO5_BC_PERIODIC ? SIDE_PERIODIC : SIDE_OUTFLOW;
    const bool whole = ny_local == P->ny;
    c->side[2] = (j_offset == 0) ? (whole || ybc == SIDE_OUTFLOW ? ybc : SIDE_NONE) : SIDE_NONE;
    c->side[3] = (j_offset + ny_local == P->ny) ? (whole || ybc == SIDE_OUTFLOW ? ybc : SIDE_NONE) : SIDE_NONE;
    if (is1d) c->side[2] = c->side[3] = SIDE_NONE;
    if (side_lo >= 0) c->side[2] = side_lo;
    if (side_hi >= 0) c->side[3] = side_hi;

    // planes: 3 states x 11 + acc 6+2+1 + rhs 7 + fsy,gsx + scratch + switch flags (dual-energy mode; in the pool so
    // that the halo push can address it)
    const size_t nplanes = 3 * 11 + 9 + 7 + 2 + 1 + 1;
    // + slack: a TMA row segment of the x sweep may run up to 66 cells past the end of the last row
    if (cudaMalloc(&c->pool, (nplanes * g.plane + 1024) * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        c->pool = nullptr;
        return fail(WENO5_ERR_CUDA, "cudaMalloc of %.2f GB failed", nplanes * g.plane * 8.0 / 1e9);
    }
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaMemsetAsync(c->pool, 0, (nplanes * g.plane + 1024) * sizeof(double), c->stream));
    double *cur = c->pool;
    c->Q0 = carve_state(cur, g.plane);
    c->QA = carve_state(cur, g.plane);
    c->QB = carve_state(cur, g.plane);
    for (int k = 0; k < 6; ++k) { c->acc[k] = cur; cur += g.plane; }
    c->accbx = cur; cur += g.plane;
    c->accby = cur; cur += g.plane;
    c->acc_by1d = cur; cur += g.plane;
    for (int k = 0; k < 7; ++k) { c->rhs[k] = cur; cur += g.plane; }
    c->fsy = cur; cur += g.plane;
    c->gsx = cur; cur += g.plane;
    c->scratch = cur; cur += g.plane;
    c->flag = cur; cur += g.plane;

    c->nplanes = nplanes;
    if (!is1d && flux_uses_tma()) {
        int bx[2], by[2];
        flux_box(0, bx);
        flux_box(1, by);
        static const int zs[4] = {8, 6, 4, 2};
        c->have_tmap = true;
        for (int k = 0; k < 4; ++k)
            c->have_tmap = c->have_tmap && make_tensor_map(&c->tmap_x[k], c->pool, g, nplanes, bx, zs[k]) &&
                           make_tensor_map(&c->tmap_y[k], c->pool, g, nplanes, by, zs[k]);
    }

    CU(cudaMalloc(&c->ctl, 8 * sizeof(double)));
    CU(cudaMemsetAsync(c->ctl, 0, 8 * sizeof(double), c->stream));
    CU(cudaMalloc(&c->vmax_bits, 8 * sizeof(unsigned long long)));
    c->vmax_bits_img = c->vmax_bits + 4;
    c->vmax_next = c->vmax_bits + 2;
    CU(cudaMalloc(&c->sums, 16 * sizeof(double)));
    CU(cudaMallocHost(&c->h_ctl, 16 * sizeof(double)));

    choose_segments(g.NXI - 5, flux_chunk(0), (g.NYI + flux_lines_per_cta(0) - 1) / flux_lines_per_cta(0),
                    c->segx_len, c->nsegx);
    if (!is1d)
        choose_segments(g.NYI - 5, flux_chunk(1), (g.NXI + flux_lines_per_cta(1) - 1) / flux_lines_per_cta(1),
                        c->segy_len, c->nsegy);
    CU(cudaStreamSynchronize(c->stream));
    cleanup.c = nullptr;
    *out = c;
    return 0;
}

extern "C" int weno5_create(const weno5_params *P, int device, int ny_local, int j_offset, weno5_ctx **out)
{
    return create_ctx(P, device, ny_local, j_offset, -1, -1, out);
}

extern "C" int weno5_destroy(weno5_ctx *c)
{
    if (!c) return 0;
    DevGuard dg_;
    dg_.set(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
    if (c->step_graph) cudaGraphExecDestroy(c->step_graph);
    // CUDA leaves freeing exported memory that another process still maps undefined.  Every importer is counted in
    // word [3] of the exporter's flag words: un-count and unmap what this rank mapped, then wait (at most 2 s: a
    // peer that died must not hang the survivors) until nobody maps this rank's memory any more.
    if (c->stream) {
        for (unsigned long long *f : c->peer_flag_maps) launch_peer_count(f + 3, -1, c->stream);
        cudaStreamSynchronize(c->stream);
    }
    for (void *p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    c->ipc_opened.clear();
    c->peer_flag_maps.clear();
    cudaGetLastError();
    if (c->sync_flags && c->stream) {
        for (int it = 0; it < 2000; ++it) {
            unsigned long long n = 0;
            if (cudaMemcpyAsync(&n, c->sync_flags + 3, sizeof(n), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) break;
            if (cudaStreamSynchronize(c->stream) != cudaSuccess || n == 0) break;
            usleep(1000);
        }
        cudaGetLastError();
    }
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    cudaFree(c->sync_flags);
    if (c->dual_pool) cudaFree(c->dual_pool);
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    if (c->ev_comm) cudaEventDestroy(c->ev_comm);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    cudaFree(c->halo);
    for (auto &e : c->ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    cudaFree(c->pool); cudaFree(c->ctl); cudaFree(c->vmax_bits); cudaFree(c->sums);
    cudaFreeHost(c->h_ctl);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

// ------------------------------------------------------------------ multi-GPU plumbing
extern "C" int weno5_comm_unique_id(unsigned char id[128])
{
    if (int r = nccl_load()) return r;
    ncclUniqueId u;
    NC(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return 0;
}

// Map the neighbours' plane pools and flag words (CUDA IPC; one process per GPU on one node).  Every rank
// publishes (IPC handle of its pool, IPC handle of its flag words, its rows) through an NCCL all-gather; if any
// rank cannot map a neighbour (same process, no peer access, WENO5_HALO=nccl) ALL ranks keep the NCCL
// send/recv exchange -- the decision is an all-reduce so that both sides of every cut use the same protocol.
struct PeerBlob { cudaIpcMemHandle_t pool, flags; int ny_local; int pid; };

static int setup_peer_halo(weno5_ctx *c)
{
    const int nranks = c->nranks;
    CU(cudaMalloc(&c->sync_flags, (8 + SHARE_SLOT_WORDS) * sizeof(unsigned long long)));
    CU(cudaMemset(c->sync_flags, 0, (8 + SHARE_SLOT_WORDS) * sizeof(unsigned long long)));
    PeerBlob mine;
    memset(&mine, 0, sizeof(mine));
    int ok = getenv("WENO5_HALO") && !strcmp(getenv("WENO5_HALO"), "nccl") ? 0 : 1;
    if (ok && (cudaIpcGetMemHandle(&mine.pool, c->pool) != cudaSuccess ||
               cudaIpcGetMemHandle(&mine.flags, c->sync_flags) != cudaSuccess)) { cudaGetLastError(); ok = 0; }
    mine.ny_local = c->ny_local;
    mine.pid = (int)getpid();
    unsigned char *d_all = nullptr;
    CU(cudaMalloc(&d_all, (size_t)(nranks + 1) * sizeof(PeerBlob)));
    CU(cudaMemcpy(d_all + (size_t)nranks * sizeof(PeerBlob), &mine, sizeof(mine), cudaMemcpyHostToDevice));
    NC(g_nccl.AllGather(d_all + (size_t)nranks * sizeof(PeerBlob), d_all, sizeof(PeerBlob), ncclUint8, c->comm, c->comm_stream));
    CU(cudaStreamSynchronize(c->comm_stream));
    std::vector<PeerBlob> all(nranks);
    CU(cudaMemcpy(all.data(), d_all, (size_t)nranks * sizeof(PeerBlob), cudaMemcpyDeviceToHost));
    // flag words of every rank (halo flags of the neighbours, all-reduce slots of everybody), mapped once each
    std::vector<unsigned long long *> flags_of(nranks, nullptr);
    auto open_flags = [&](int peer) -> unsigned long long * {
        if (peer == c->rank) return c->sync_flags;
        if (flags_of[peer]) return flags_of[peer];
        if (all[peer].pid == mine.pid) return nullptr;                           // IPC maps other processes only
        void *f = nullptr;
        if (cudaIpcOpenMemHandle(&f, all[peer].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        c->ipc_opened.push_back(f);
        c->peer_flag_maps.push_back((unsigned long long *)f);
        launch_peer_count((unsigned long long *)f + 3, 1, c->comm_stream);      // counted as an importer of that rank
        return flags_of[peer] = (unsigned long long *)f;
    };
    auto open = [&](int peer, double *&pool, unsigned long long *&flags, int &ny) {
        if (peer < 0) return true;
        if (peer == c->rank || all[peer].pid == mine.pid) return false;
        flags = open_flags(peer);
        if (!flags) return false;
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[peer].pool, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return false; }
        c->ipc_opened.push_back(p);
        pool = (double *)p; ny = all[peer].ny_local;
        return true;
    };
    if (ok) {
        ok = open(c->up, c->peer_pool_up, c->peer_flags_up, c->peer_ny_up) ? 1 : 0;
        if (ok && c->down >= 0) {
            if (c->down == c->up) {                       // two ranks, periodic: the same neighbour on both sides
                c->peer_pool_down = c->peer_pool_up; c->peer_flags_down = c->peer_flags_up; c->peer_ny_down = c->peer_ny_up;
            } else {
                ok = open(c->down, c->peer_pool_down, c->peer_flags_down, c->peer_ny_down) ? 1 : 0;
            }
        }
    }
    // all-reduce over peer memory: needs every rank's slot array
    int share = ok && nranks <= SHARE_MAX_RANKS && !(getenv("WENO5_ALLREDUCE") && !strcmp(getenv("WENO5_ALLREDUCE"), "nccl")) ? 1 : 0;
    for (int r = 0; share && r < nranks; ++r) {
        unsigned long long *f = open_flags(r);
        if (!f) share = 0; else c->share_slots[r] = f + 8;
    }
    // unanimous or not at all (bit 0: halo push, bit 1: all-reduce)
    unsigned long long *d_ok = reinterpret_cast<unsigned long long *>(d_all);
    const unsigned long long h_ok = (unsigned long long)(ok && share ? 3 : (ok ? 1 : 0));
    CU(cudaMemcpy(d_ok, &h_ok, sizeof(h_ok), cudaMemcpyHostToDevice));
    NC(g_nccl.AllReduce(d_ok, d_ok, 1, ncclUint64, ncclMin, c->comm, c->comm_stream));
    CU(cudaStreamSynchronize(c->comm_stream));
    unsigned long long agreed = 0;
    CU(cudaMemcpy(&agreed, d_ok, sizeof(agreed), cudaMemcpyDeviceToHost));
    CU(cudaFree(d_all));
    c->peer_ok = (agreed & 1ull) != 0;
    c->share_ok = agreed == 3ull;
    return 0;
}

extern "C" int weno5_comm_halo_mode(const weno5_ctx *c) { return c && c->comm ? (c->peer_ok ? (c->share_ok ? 3 : 2) : 1) : 0; }

extern "C" int weno5_comm_init(weno5_ctx *c, int rank, int nranks, const unsigned char id[128])
{
    if (!c || !id) return fail(WENO5_ERR_ARG, "null argument");
    if (c->g.is1d) return fail(WENO5_ERR_ARG, "1-D grids cannot be decomposed");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(WENO5_ERR_ARG, "bad rank %d of %d", rank, nranks);
    if (int r = nccl_load()) return r;
    DevGuard dg_; CU(dg_.set(c->device));
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    NC(g_nccl.CommInitRank(&c->comm, nranks, u, rank));
    CU(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
    CU(cudaMalloc(&c->halo, 4 * (size_t)NCELL * G * c->g.pitch * sizeof(double)));
    CU(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_comm, cudaEventDisableTiming));
    c->rank = rank;
    c->nranks = nranks;
    const bool per = c->P.bc_y == WENO5_BC_PERIODIC;
    c->up = rank + 1 < nranks ? rank + 1 : (per && nranks > 1 ? 0 : -1);
    c->down = rank > 0 ? rank - 1 : (per && nranks > 1 ? nranks - 1 : -1);
    if (int r = setup_peer_halo(c)) return r;
    return 0;
}

// make the main stream wait for an exchange that is still in flight on the NCCL stream
// the main stream waits for work queued on the NCCL stream (all-reduce of the wave speeds, NCCL halo exchange)
static int wait_comm_event(weno5_ctx *c)
{
    if (c->xchg_pending) {
        CU(cudaStreamWaitEvent(c->stream, c->ev_comm, 0));
        c->xchg_pending = false;
    }
    return 0;
}

static int wait_exchange(weno5_ctx *c)
{
    if (c->peer_wait_seq) {
        launch_halo_wait(c->sync_flags, c->down >= 0, c->up >= 0, c->peer_wait_seq, c->ctl, c->stream);
        c->peer_wait_seq = 0;
        c->launches += 1;
    }
    if (c->xchg_pending) {
        CU(cudaStreamWaitEvent(c->stream, c->ev_comm, 0));
        c->xchg_pending = false;
    }
    return 0;
}

// send the 4 outermost interior rows, receive the 4 ghost rows, for each plane: a push over peer memory on the main
// stream, or (fallback) packed NCCL messages on the NCCL stream.  With `async` (peer path only) nothing waits here:
// the next x sweep does, inside the kernel; whoever else touches ghost rows calls wait_exchange first.
static int exchange_rows(weno5_ctx *c, double *const planes[], int n, bool async = false)
{
    if (!c->comm || (c->up < 0 && c->down < 0)) return 0;
    // peer push + a flux kernel whose ghost-row strips wait for the flag words themselves: the wait is left to the
    // next stage's x sweep (enqueue_stage), hidden behind its interior rows; every other combination waits here
    const bool inkernel = c->peer_ok && c->have_tmap && !c->dual && flux_waits_in_kernel();
    if (!inkernel) async = false;
    const Geom &g = c->g;
    if (c->peer_ok) {
        // push the edge rows into the neighbours' ghost rows and raise their flag words; then (or later, async)
        // wait for the neighbours' pushes of the same sequence number
        if (int r = wait_exchange(c)) return r;           // at most one exchange in flight
        PushArgs a;
        memset(&a, 0, sizeof(a));
        a.n = n; a.NYI = g.NYI; a.pitch = g.pitch;
        for (int k = 0; k < n; ++k) { a.src[k] = planes[k]; a.plane_idx[k] = (int)((planes[k] - c->pool) / (ptrdiff_t)g.plane); }
        a.peer_up = c->up >= 0 ? c->peer_pool_up : nullptr;
        a.peer_down = c->down >= 0 ? c->peer_pool_down : nullptr;
        a.plane_up = (size_t)g.pitch * (c->peer_ny_up + 2 * G);
        a.plane_down = (size_t)g.pitch * (c->peer_ny_down + 2 * G);
        a.NYI_down = c->peer_ny_down + 2 * G;
        a.flag_up = c->up >= 0 ? c->peer_flags_up + 0 : nullptr;        // I am its lower neighbour
        a.flag_down = c->down >= 0 ? c->peer_flags_down + 1 : nullptr;  // I am its upper neighbour
        a.counter = reinterpret_cast<unsigned int *>(c->sync_flags + 2);
        a.seq = ++c->xseq;
        launch_halo_push(a, c->stream);
        c->launches += 1;
        c->peer_wait_seq = a.seq;
        if (!async) return wait_exchange(c);
        return 0;
    }
    const size_t cnt = (size_t)G * g.pitch;
    CU(cudaEventRecord(c->ev_main, c->stream));
    CU(cudaStreamWaitEvent(c->comm_stream, c->ev_main, 0));
    // one packed message per neighbour: [plane][4 rows][pitch]
    const size_t msg = (size_t)NCELL * G * g.pitch;
    double *s_up = c->halo, *s_down = c->halo + msg, *r_up = c->halo + 2 * msg, *r_down = c->halo + 3 * msg;
    launch_halo_pack(planes, n, g, s_up, s_down, c->up >= 0, c->down >= 0, 0, c->comm_stream);
    NC(g_nccl.GroupStart());
    if (c->up >= 0) NC(g_nccl.Send(s_up, cnt * n, ncclFloat64, c->up, c->comm, c->comm_stream));
    if (c->down >= 0) {
        NC(g_nccl.Recv(r_down, cnt * n, ncclFloat64, c->down, c->comm, c->comm_stream));
        NC(g_nccl.Send(s_down, cnt * n, ncclFloat64, c->down, c->comm, c->comm_stream));
    }
    if (c->up >= 0) NC(g_nccl.Recv(r_up, cnt * n, ncclFloat64, c->up, c->comm, c->comm_stream));
    NC(g_nccl.GroupEnd());
    launch_halo_pack(planes, n, g, r_up, r_down, c->up >= 0, c->down >= 0, 1, c->comm_stream);
    c->launches += 2;
    CU(cudaEventRecord(c->ev_comm, c->comm_stream));
    c->xchg_pending = true;
    c->launches += 1;
    if (!async) return wait_exchange(c);
    return 0;
}

// ------------------------------------------------------------------ transfers
// host component order is the reference's [rho,Mx,My,Mz,Bx,By,Bz,S,E]; on the device E sits before S
// so that the eight fields the sweeps read are planes 0..7
static inline int host_comp(int k) { return k == C_E ? 8 : (k == C_S ? 7 : k); }

static int copy_plane(weno5_ctx *c, double *dev, const double *host_in, double *host_out)
{
    const Geom &g = c->g;
    double *d = dev + 1 + (g.is1d ? 0 : g.pitch);          // public (1,1) -> internal (1,1)
    const size_t w = (size_t)g.Nx * sizeof(double);
    if (host_in) CU(cudaMemcpy2DAsync(d, g.pitch * sizeof(double), host_in, w, w, g.Ny, cudaMemcpyHostToDevice, c->stream));
    if (host_out) CU(cudaMemcpy2DAsync(host_out, w, d, g.pitch * sizeof(double), w, g.Ny, cudaMemcpyDeviceToHost, c->stream));
    return 0;
}

static int apply_boundaries(weno5_ctx *c, State &S, bool faces)
{
    launch_bc_fill(S.q, NCELL, c->g, c->side, nullptr, c->stream);
    c->launches++;
    if (c->comm) { if (int r = exchange_rows(c, S.q, NCELL)) return r; }
    if (faces && !c->g.is1d) {
        launch_face_bc(S.bx, S.by, c->g, c->side, nullptr, c->stream);
        c->launches++;
        if (c->comm) {
            double *f[2] = {S.bx, S.by};
            if (int r = exchange_rows(c, f, 2)) return r;
        }
    }
    CU(cudaGetLastError());
    return 0;
}

extern "C" int weno5_upload(weno5_ctx *c, const double *q, const double *bxb, const double *byb)
{
    if (!c || !q) return fail(WENO5_ERR_ARG, "null argument");
    if (c->g.is1d) return fail(WENO5_ERR_ARG, "use weno5_upload_1d for 1-D grids");
    if ((bxb == nullptr) != (byb == nullptr)) return fail(WENO5_ERR_ARG, "pass both face fields or neither");
    DevGuard dg_; CU(dg_.set(c->device));
    const size_t hp = (size_t)c->g.Nx * c->g.Ny;
    for (int k = 0; k < NCELL; ++k) if (int r = copy_plane(c, c->Q0.q[k], q + host_comp(k) * hp, nullptr)) return r;
    if (bxb) {
        if (int r = copy_plane(c, c->Q0.bx, bxb, nullptr)) return r;
        if (int r = copy_plane(c, c->Q0.by, byb, nullptr)) return r;
    } else {
        launch_center2face(c->Q0.q[C_BX], c->Q0.q[C_BY], c->Q0.bx, c->Q0.by, c->g, c->stream);
        c->launches++;
    }
    if (int r = apply_boundaries(c, c->Q0, true)) return r;
    CU(cudaStreamSynchronize(c->stream));
    c->uploaded = true;
    return 0;
}

extern "C" int weno5_download(weno5_ctx *c, double *q, double *bxb, double *byb)
{
    if (!c) return fail(WENO5_ERR_ARG, "null argument");
    if (c->g.is1d) return fail(WENO5_ERR_ARG, "use weno5_download_1d for 1-D grids");
    if (!c->uploaded) return fail(WENO5_ERR_STATE, "no state on the device yet");
    DevGuard dg_; CU(dg_.set(c->device));
    if (int r = wait_exchange(c)) return r;
    const size_t hp = (size_t)c->g.Nx * c->g.Ny;
    if (q) for (int k = 0; k < NCELL; ++k) if (int r = copy_plane(c, c->Q0.q[k], nullptr, q + host_comp(k) * hp)) return r;
    if (bxb) if (int r = copy_plane(c, c->Q0.bx, nullptr, bxb)) return r;
    if (byb) if (int r = copy_plane(c, c->Q0.by, nullptr, byb)) return r;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int weno5_upload_1d(weno5_ctx *c, const double *q)
{
    if (!c || !q) return fail(WENO5_ERR_ARG, "null argument");
    if (!c->g.is1d) return fail(WENO5_ERR_ARG, "handle is 2-D");
    DevGuard dg_; CU(dg_.set(c->device));
    for (int k = 0; k < NCELL; ++k) if (int r = copy_plane(c, c->Q0.q[k], q + (size_t)host_comp(k) * c->g.Nx, nullptr)) return r;
    if (int r = apply_boundaries(c, c->Q0, false)) return r;
    CU(cudaStreamSynchronize(c->stream));
    c->uploaded = true;
    return 0;
}

extern "C" int weno5_download_1d(weno5_ctx *c, double *q)
{
    if (!c || !q) return fail(WENO5_ERR_ARG, "null argument");
    if (!c->g.is1d) return fail(WENO5_ERR_ARG, "handle is 2-D");
    if (!c->uploaded) return fail(WENO5_ERR_STATE, "no state on the device yet");
    DevGuard dg_; CU(dg_.set(c->device));
    for (int k = 0; k < NCELL; ++k) if (int r = copy_plane(c, c->Q0.q[k], nullptr, q + (size_t)host_comp(k) * c->g.Nx)) return r;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int weno5_boundaries(weno5_ctx *c)
{
    if (!c) return fail(WENO5_ERR_ARG, "null argument");
    if (!c->uploaded) return fail(WENO5_ERR_STATE, "no state on the device yet");
    DevGuard dg_; CU(dg_.set(c->device));
    if (int r = apply_boundaries(c, c->Q0, true)) return r;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------ time step
// fused = true: the maxima were accumulated by stage 4 of the previous step (ct_centers_entropy_vmax) into
// vmax_next; no pass over the state is needed, and dt_kernel clears the accumulator for the next step
static int enqueue_timestep(weno5_ctx *c, bool fused = false)
{
    unsigned long long *vb = fused ? c->vmax_next : c->vmax_bits;
    if (!fused) {
        launch_vmax(c->Q0.q, c->g, c->P.gamma, vb, c->stream);
        c->launches += 1;
    }
    ShareArgs share;
    const bool peer_reduce = c->comm && c->nranks > 1 && c->share_ok;
    if (peer_reduce) {
        share.n = c->nranks; share.rank = c->rank; share.seq = ++c->share_seq;
        for (int r = 0; r < SHARE_MAX_RANKS; ++r) share.slots[r] = c->share_slots[r];
    } else if (c->comm && c->nranks > 1) {
        // all NCCL work of a handle is ordered on the NCCL stream (after a still-running exchange)
        CU(cudaEventRecord(c->ev_main, c->stream));
        CU(cudaStreamWaitEvent(c->comm_stream, c->ev_main, 0));
        NC(g_nccl.AllReduce(vb, vb, 2, ncclUint64, ncclMax, c->comm, c->comm_stream));
        CU(cudaEventRecord(c->ev_comm, c->comm_stream));
        c->xchg_pending = true;
        if (int r = wait_comm_event(c)) return r;       // (a halo push still in flight stays pending: dt needs no ghost rows)
        c->launches += 1;
    }
    launch_dt_from_vmax(vb, c->ctl, c->P.dx, c->P.dy, c->P.cfl, c->g.is1d, fused ? 1 : 0, peer_reduce ? &share : nullptr, c->stream);
    c->launches += 1;
    return 0;
}

extern "C" int weno5_timestep(weno5_ctx *c, double *dt, double *vxmax, double *vymax)
{
    if (!c) return fail(WENO5_ERR_ARG, "null argument");
    if (!c->uploaded) return fail(WENO5_ERR_STATE, "no state on the device yet");
    DevGuard dg_; CU(dg_.set(c->device));
    const double zero[2] = {0.0, 0.0};
    CU(cudaMemcpyAsync(c->ctl + 4, zero, sizeof(double), cudaMemcpyHostToDevice, c->stream));   // tmax off
    if (int r = enqueue_timestep(c)) return r;
    CU(cudaMemcpyAsync(c->h_ctl, c->ctl, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    if (c->h_ctl[6] != 0.0) {
        CU(cudaMemsetAsync(c->ctl + 6, 0, sizeof(double), c->stream));
        return fail(WENO5_ERR_NUMERIC, "non-finite wave speed (vxmax=%g vymax=%g)", c->h_ctl[2], c->h_ctl[3]);
    }
    if (dt) *dt = c->h_ctl[0];
    if (vxmax) *vxmax = c->h_ctl[2];
    if (vymax) *vymax = c->h_ctl[3];
    return 0;
}

// ------------------------------------------------------------------ one RK4 step
struct ProfScope {
    weno5_ctx *c; size_t idx; bool on;
    ProfScope(weno5_ctx *c_) : c(c_), idx(0), on(c_->profiling)
    {
        if (!on) return;
        if (c->ev_used == c->ev.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            c->ev.emplace_back(a, b);
        }
        idx = c->ev_used++;
        cudaEventRecord(c->ev[idx].first, c->stream);
    }
    ~ProfScope() { if (on) cudaEventRecord(c->ev[idx].second, c->stream); }
};

static int enqueue_stage(weno5_ctx *c, int stage, State &cur, State &nxt)
{
    static const double ck[4] = {0.5, 0.5, 1.0, 1.0 / 6.0};
    const weno5_params &P = c->P;
    const Geom &g = c->g;
    const double *dtp = c->ctl;
    // the six flux-advanced cell fields
    static const int six[6] = {C_RHO, C_MX, C_MY, C_MZ, C_BZ, C_E};

    FluxArgs a;
    memset(&a, 0, sizeof(a));
    for (int k = 0; k < 8; ++k) a.qc[k] = cur.q[k];
    for (int k = 0; k < 7; ++k) a.rhs[k] = c->rhs[k];
    a.dtp = dtp;
    a.c_dx = ck[stage - 1] / P.dx;
    a.c_dy = g.is1d ? 0.0 : ck[stage - 1] / P.dy;
    a.stage = stage;
    a.gam = P.gamma; a.eps = P.eps; a.g2 = (P.gamma - 2.0) / (P.gamma - 1.0);
    a.is1d = g.is1d;
    a.pitch = g.pitch;
    a.tmap_x = c->have_tmap ? c->tmap_x : nullptr;
    a.tmap_y = c->have_tmap ? c->tmap_y : nullptr;
    auto plane_of = [&](const double *p) { return (int)((p - c->pool) / (ptrdiff_t)g.plane); };
    for (int k = 0; k < 8; ++k) a.qc_plane[k] = plane_of(cur.q[k]);
    for (int k = 0; k < 7; ++k) a.rhs_plane[k] = plane_of(c->rhs[k]);

    // x sweep
    a.N = g.NXI; a.NP = g.NYI; a.bs = c->fsy;
    a.seg_len = c->segx_len; a.nseg = c->nsegx;
    if (c->peer_wait_seq && !c->xchg_pending && !g.is1d && c->have_tmap && flux_waits_in_kernel()) {
        // ghost rows of `cur` are still arriving over peer memory: the sweep's ghost-row strips run last and wait for
        // the neighbours' flag words inside the kernel
        a.wait_flags = c->sync_flags; a.wait_seq = c->peer_wait_seq;
        a.wait_lo = c->down >= 0; a.wait_hi = c->up >= 0; a.wait_err = c->ctl;
        c->peer_wait_seq = 0;
        { ProfScope ps(c); launch_flux_x(a, c->stream); }
        a.wait_flags = nullptr;
        c->launches++; c->prof_launches += c->profiling;
    } else {
        if (int r = wait_exchange(c)) return r;          // an exchange still in flight (NCCL path, other flux variants)
        { ProfScope ps(c); launch_flux_x(a, c->stream); }
        c->launches++; c->prof_launches += c->profiling;
    }

    if (g.is1d) {
        double *qn7[7]; const double *base7[7]; double *acc7[7];
        static const int seven[7] = {C_RHO, C_MX, C_MY, C_MZ, C_BZ, C_E, C_BY};
        for (int k = 0; k < 7; ++k) {
            qn7[k] = nxt.q[seven[k]];
            base7[k] = c->Q0.q[seven[k]];
            acc7[k] = k < 6 ? c->acc[k] : c->acc_by1d;
        }
        launch_update_1d(a, qn7, base7, acc7, c->stream);
        c->launches++;
        // Bx is constant in 1-D (Weno5_1D.jl:771)
        if (nxt.q[C_BX] != c->Q0.q[C_BX])
            CU(cudaMemcpyAsync(nxt.q[C_BX], c->Q0.q[C_BX], g.plane * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        // 1-D keeps E (Weno5_1D.jl:252-254); S = E2S(q) is only read by the next timestep
        if (stage == 4) { launch_entropy(nxt.q, g, P.gamma, 0, 0, dtp, c->stream); c->launches++; }
        launch_bc_fill(nxt.q, stage == 4 ? NCELL : 8, g, c->side, dtp, c->stream);
        c->launches++;
        CU(cudaGetLastError());
        return 0;
    }

    // y sweep + RK update
    a.N = g.NYI; a.NP = g.NXI; a.bs = c->gsx;
    a.seg_len = c->segy_len; a.nseg = c->nsegy;
    for (int k = 0; k < 6; ++k) {
        a.base[k] = c->Q0.q[six[k]]; a.acc[k] = c->acc[k]; a.qn[k] = nxt.q[six[k]];
        a.base_plane[k] = plane_of(a.base[k]); a.acc_plane[k] = plane_of(a.acc[k]);
    }
    { ProfScope ps(c); launch_flux_y(a, c->stream); }
    c->launches++; c->prof_launches += c->profiling;

    // constrained transport
    CtArgs t;
    memset(&t, 0, sizeof(t));
    t.fsy = c->fsy; t.gsx = c->gsx;
    t.bx0 = c->Q0.bx; t.by0 = c->Q0.by; t.bxc = cur.bx; t.byc = cur.by;
    t.accbx = c->accbx; t.accby = c->accby; t.bxn = nxt.bx; t.byn = nxt.by;
    t.dtp = dtp; t.c_dx = a.c_dx; t.c_dy = a.c_dy; t.stage = stage;
    t.Nx = g.Nx; t.Ny = g.Ny; t.pitch = g.pitch;
    t.oxlo = c->side[0] == SIDE_OUTFLOW; t.oxhi = c->side[1] == SIDE_OUTFLOW;
    t.oylo = c->side[2] == SIDE_OUTFLOW; t.oyhi = c->side[3] == SIDE_OUTFLOW;
    // outflow sides follow the reference's full-array update; elsewhere only faces whose corner
    // EMFs have complete stencils are written (they are what the 4th-order centring needs)
    t.bx_i0 = t.oxlo ? 1 : 2; t.bx_i1 = t.oxhi ? g.Nx : g.Nx - 2;
    t.bx_j0 = t.oylo ? 1 : 3; t.bx_j1 = t.oyhi ? g.Ny : g.Ny - 2;
    t.by_i0 = t.oxlo ? 1 : 3; t.by_i1 = t.oxhi ? g.Nx : g.Nx - 2;
    t.by_j0 = t.oylo ? 1 : 2; t.by_j1 = t.oyhi ? g.Ny : g.Ny - 2;
    // (a fused corner/face/centre kernel with halo recomputation in shared memory was measured 1.6 %
    // slower than these two launches and removed)
    launch_ct_faces(t, c->stream);
    // cell-centred B from the new faces and, when this stage derives S (every stage with the 2-D
    // reference's E -> S -> E round trip, else stage 4 only), the ES_Switch residue in the same pass
    const bool ent = P.es_roundtrip || stage == 4;
    if (ent && stage == 4 && c->fuse_vmax)
        launch_ct_centers_entropy_vmax(nxt.bx, nxt.by, nxt.q, g, P.gamma, P.es_roundtrip, dtp, c->side[1], c->vmax_next, c->stream);
    else if (ent) launch_ct_centers_entropy(nxt.bx, nxt.by, nxt.q, g, P.gamma, P.es_roundtrip, dtp, c->stream);
    else launch_ct_centers(nxt.bx, nxt.by, nxt.q[C_BX], nxt.q[C_BY], g, dtp, c->stream);
    c->launches += 2;
    if (stage == 4) {                                   // boundaries!(q4, bxb4, byb4), Weno5_2D.jl:313
        launch_face_bc(nxt.bx, nxt.by, g, c->side, dtp, c->stream);
        c->launches++;
    }
    const int nfill = ent ? NCELL : 8;
    launch_bc_fill(nxt.q, nfill, g, c->side, dtp, c->stream);
    c->launches++;
    if (c->comm) { if (int r = exchange_rows(c, nxt.q, nfill, true)) return r; }
    CU(cudaGetLastError());
    return 0;
}

// One RK stage with both branches live (Weno5_2D.jl:134-180): E sweeps -> E candidate, S sweeps -> S candidate,
// constrained transport for each (their cell-centred B enters E2S / S2E), ES_Switch, selection of the CT seed
// fluxes, constrained transport with the selected fluxes.  Single GPU, 2-D.
static int enqueue_stage_dual(weno5_ctx *c, int stage, State &cur, State &nxt)
{
    static const double ck[4] = {0.5, 0.5, 1.0, 1.0 / 6.0};
    const weno5_params &P = c->P;
    const Geom &g = c->g;
    const double *dtp = c->ctl;
    static const int sixE[6] = {C_RHO, C_MX, C_MY, C_MZ, C_BZ, C_E};
    static const int sixS[6] = {C_RHO, C_MX, C_MY, C_MZ, C_BZ, C_S};
    auto plane_of = [&](const double *p) { return (int)((p - c->pool) / (ptrdiff_t)g.plane); };
    if (int r = wait_exchange(c)) return r;

    for (int br = 0; br < 2; ++br) {
        const int *six = br == 0 ? sixE : sixS;
        FluxArgs a;
        memset(&a, 0, sizeof(a));
        for (int k = 0; k < 7; ++k) a.qc[k] = cur.q[k];
        a.qc[7] = cur.q[br == 0 ? C_E : C_S];
        for (int k = 0; k < 7; ++k) a.rhs[k] = c->rhs[k];
        a.dtp = dtp;
        a.c_dx = ck[stage - 1] / P.dx;
        a.c_dy = ck[stage - 1] / P.dy;
        a.stage = stage;
        a.gam = P.gamma; a.eps = P.eps; a.g2 = (P.gamma - 2.0) / (P.gamma - 1.0);
        a.pitch = g.pitch;
        a.branch = br;
        const bool tma = br == 0 && c->have_tmap;
        a.tmap_x = tma ? c->tmap_x : nullptr;
        a.tmap_y = tma ? c->tmap_y : nullptr;
        if (tma) {
            for (int k = 0; k < 8; ++k) a.qc_plane[k] = plane_of(cur.q[k]);
            for (int k = 0; k < 7; ++k) a.rhs_plane[k] = plane_of(c->rhs[k]);
        }
        a.N = g.NXI; a.NP = g.NYI; a.bs = br == 0 ? c->fsy : c->fsyS;
        a.seg_len = c->segx_len; a.nseg = c->nsegx;
        launch_flux_x(a, c->stream);
        a.N = g.NYI; a.NP = g.NXI; a.bs = br == 0 ? c->gsx : c->gsxS;
        a.seg_len = c->segy_len; a.nseg = c->nsegy;
        for (int k = 0; k < 6; ++k) {
            a.base[k] = c->Q0.q[six[k]];
            a.acc[k] = br == 0 ? c->acc[k] : c->accS[k];
            a.qn[k] = br == 0 ? c->candE[k] : c->candS[k];
            if (tma) { a.base_plane[k] = plane_of(a.base[k]); a.acc_plane[k] = plane_of(a.acc[k]); }
        }
        launch_flux_y(a, c->stream);
        c->launches += 2;
    }

    CtArgs t;
    memset(&t, 0, sizeof(t));
    t.bx0 = c->Q0.bx; t.by0 = c->Q0.by; t.bxc = cur.bx; t.byc = cur.by;
    t.accbx = c->accbx; t.accby = c->accby;
    t.dtp = dtp; t.c_dx = ck[stage - 1] / P.dx; t.c_dy = ck[stage - 1] / P.dy; t.stage = stage;
    t.Nx = g.Nx; t.Ny = g.Ny; t.pitch = g.pitch;
    t.oxlo = c->side[0] == SIDE_OUTFLOW; t.oxhi = c->side[1] == SIDE_OUTFLOW;
    t.oylo = c->side[2] == SIDE_OUTFLOW; t.oyhi = c->side[3] == SIDE_OUTFLOW;
    t.bx_i0 = t.oxlo ? 1 : 2; t.bx_i1 = t.oxhi ? g.Nx : g.Nx - 2;
    t.bx_j0 = t.oylo ? 1 : 3; t.bx_j1 = t.oyhi ? g.Ny : g.Ny - 2;
    t.by_i0 = t.oxlo ? 1 : 3; t.by_i1 = t.oxhi ? g.Nx : g.Nx - 2;
    t.by_j0 = t.oylo ? 1 : 2; t.by_j1 = t.oyhi ? g.Ny : g.Ny - 2;
    // candidate face fields only feed the candidates' cell-centred B (:150-153, :168-170)
    t.no_acc = 1; t.bxn = c->tbx; t.byn = c->tby;
    t.fsy = c->fsy; t.gsx = c->gsx;
    launch_ct_faces(t, c->stream);
    launch_ct_centers(c->tbx, c->tby, c->candE[6], c->candE[7], g, dtp, c->stream);
    t.fsy = c->fsyS; t.gsx = c->gsxS;
    launch_ct_faces(t, c->stream);
    launch_ct_centers(c->tbx, c->tby, c->candS[6], c->candS[7], g, dtp, c->stream);

    SwitchArgs w;
    for (int k = 0; k < 8; ++k) { w.ce[k] = c->candE[k]; w.cs[k] = c->candS[k]; }
    static const int sel[7] = {C_RHO, C_MX, C_MY, C_MZ, C_BZ, C_S, C_E};
    for (int k = 0; k < 7; ++k) w.qn[k] = nxt.q[sel[k]];
    w.flag = c->flag;
    w.NXI = g.NXI; w.NYI = g.NYI; w.pitch = g.pitch; w.gam = P.gamma; w.threshold = c->es_threshold; w.dtp = dtp;
    w.qn_by = nullptr; w.flags_only = c->dual_flags_only ? 1 : 0;
    launch_es_switch(w, c->stream);
    double *fl[1] = {c->flag};
    launch_bc_fill(fl, 1, g, c->side, dtp, c->stream);
    // y-slabs: the switch decisions of the neighbours' edge rows select the seed fluxes of this slab's ghost-row faces
    if (c->comm) { if (int r = exchange_rows(c, fl, 1)) return r; }
    launch_flux_select(c->flag, c->fsy, c->gsx, c->fsyS, c->gsxS, g, dtp, c->stream);

    // constrained transport with the selected seed fluxes (:175-180)
    t.no_acc = 0; t.bxn = nxt.bx; t.byn = nxt.by;
    launch_ct_faces(t, c->stream);
    launch_ct_centers(nxt.bx, nxt.by, nxt.q[C_BX], nxt.q[C_BY], g, dtp, c->stream);
    c->launches += 9;
    if (stage == 4) {
        launch_face_bc(nxt.bx, nxt.by, g, c->side, dtp, c->stream);
        c->launches++;
    }
    launch_bc_fill(nxt.q, NCELL, g, c->side, dtp, c->stream);
    c->launches++;
    if (c->comm) { if (int r = exchange_rows(c, nxt.q, NCELL)) return r; }
    CU(cudaGetLastError());
    return 0;
}

// 1-D twin (Weno5_1D.jl:152-236): per stage the energy sweep and the entropy sweep of the switched state, the two
// RK candidates, ES_Switch (:238-258).  candE/candS hold rho,Mx,My,Mz,Bz,E|S,(unused),By.
static int enqueue_stage_dual_1d(weno5_ctx *c, int stage, State &cur, State &nxt)
{
    static const double ck[4] = {0.5, 0.5, 1.0, 1.0 / 6.0};
    const weno5_params &P = c->P;
    const Geom &g = c->g;
    const double *dtp = c->ctl;
    for (int br = 0; br < 2; ++br) {
        FluxArgs a;
        memset(&a, 0, sizeof(a));
        for (int k = 0; k < 7; ++k) a.qc[k] = cur.q[k];
        a.qc[7] = cur.q[br == 0 ? C_E : C_S];
        for (int k = 0; k < 7; ++k) a.rhs[k] = c->rhs[k];
        a.dtp = dtp;
        a.c_dx = ck[stage - 1] / P.dx;
        a.stage = stage;
        a.gam = P.gamma; a.eps = P.eps; a.g2 = (P.gamma - 2.0) / (P.gamma - 1.0);
        a.is1d = 1;
        a.pitch = g.pitch;
        a.branch = br;
        a.N = g.NXI; a.NP = g.NYI; a.bs = c->fsy;
        a.seg_len = c->segx_len; a.nseg = c->nsegx;
        launch_flux_x(a, c->stream);
        double *qn7[7]; const double *base7[7]; double *acc7[7];
        double **cand = br == 0 ? c->candE : c->candS;
        const int en = br == 0 ? C_E : C_S;
        const int seven[7] = {C_RHO, C_MX, C_MY, C_MZ, C_BZ, en, C_BY};
        for (int k = 0; k < 7; ++k) {
            qn7[k] = k < 6 ? cand[k] : cand[7];
            base7[k] = c->Q0.q[seven[k]];
            acc7[k] = br == 0 ? (k < 6 ? c->acc[k] : c->acc_by1d) : (k < 6 ? c->accS[k] : c->tbx);
        }
        launch_update_1d(a, qn7, base7, acc7, c->stream);
        c->launches += 2;
    }
    SwitchArgs w;
    for (int k = 0; k < 8; ++k) { w.ce[k] = c->candE[k]; w.cs[k] = c->candS[k]; }
    w.ce[6] = w.cs[6] = c->Q0.q[C_BX];                     // Bx is constant in 1-D (Weno5_1D.jl:771)
    static const int sel[7] = {C_RHO, C_MX, C_MY, C_MZ, C_BZ, C_S, C_E};
    for (int k = 0; k < 7; ++k) w.qn[k] = nxt.q[sel[k]];
    w.qn_by = nxt.q[C_BY];
    w.flag = c->flag;
    w.NXI = g.NXI; w.NYI = g.NYI; w.pitch = g.pitch; w.gam = P.gamma; w.threshold = c->es_threshold; w.dtp = dtp;
    w.flags_only = c->dual_flags_only ? 1 : 0;
    launch_es_switch(w, c->stream);
    if (nxt.q[C_BX] != c->Q0.q[C_BX])
        CU(cudaMemcpyAsync(nxt.q[C_BX], c->Q0.q[C_BX], g.plane * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    launch_bc_fill(nxt.q, NCELL, g, c->side, dtp, c->stream);
    c->launches += 2;
    CU(cudaGetLastError());
    return 0;
}

static int enqueue_step(weno5_ctx *c)
{
    if (c->dual && c->g.is1d) {
        if (int r = enqueue_stage_dual_1d(c, 1, c->Q0, c->QA)) return r;
        if (int r = enqueue_stage_dual_1d(c, 2, c->QA, c->QB)) return r;
        if (int r = enqueue_stage_dual_1d(c, 3, c->QB, c->QA)) return r;
        if (int r = enqueue_stage_dual_1d(c, 4, c->QA, c->Q0)) return r;
        return 0;
    }
    if (c->dual) {
        if (int r = enqueue_stage_dual(c, 1, c->Q0, c->QA)) return r;
        if (int r = enqueue_stage_dual(c, 2, c->QA, c->QB)) return r;
        if (int r = enqueue_stage_dual(c, 3, c->QB, c->QA)) return r;
        if (int r = enqueue_stage_dual(c, 4, c->QA, c->Q0)) return r;
        return 0;
    }
    if (int r = enqueue_stage(c, 1, c->Q0, c->QA)) return r;
    if (int r = enqueue_stage(c, 2, c->QA, c->QB)) return r;
    if (int r = enqueue_stage(c, 3, c->QB, c->QA)) return r;
    if (int r = enqueue_stage(c, 4, c->QA, c->Q0)) return r;
    return 0;
}

// 1-D, energy branch, line fits one cluster: the whole step loop runs in one cluster-resident kernel
static bool use_line1d(const weno5_ctx *c) { return c->g.is1d && !c->dual && !c->profiling && line1d_chunk(c->g.NXI) > 0; }

extern "C" int weno5_rk4_step(weno5_ctx *c, double dt)
{
    if (!c) return fail(WENO5_ERR_ARG, "null argument");
    if (!c->uploaded) return fail(WENO5_ERR_STATE, "no state on the device yet");
    if (!(dt > 0.0) || !std::isfinite(dt)) return fail(WENO5_ERR_ARG, "dt must be positive and finite (got %g)", dt);
    DevGuard dg_; CU(dg_.set(c->device));
    c->h_ctl[8] = dt;
    CU(cudaMemcpyAsync(c->ctl, c->h_ctl + 8, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (use_line1d(c)) {
        CU(launch_line1d(c->Q0.q, c->ctl, c->g, c->side, 1, dt, c->P.dx, c->P.gamma, c->P.eps, c->P.cfl, c->stream));
        c->launches++;
    } else {
        if (int r = enqueue_step(c)) return r;
    }
    if (int r = wait_exchange(c)) return r;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    return 0;
}

// Capture one step of the device-resident loop (no host interaction inside: dt, t and the maxima live in device
// memory, the state ping-pong returns to Q0 every step) into a CUDA graph.
static int build_step_graph(weno5_ctx *c, int key)
{
    if (c->step_graph && c->graph_key == key) return 0;
    if (c->step_graph) { cudaGraphExecDestroy(c->step_graph); c->step_graph = nullptr; }
    cudaGraph_t g = nullptr;
    CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const unsigned long long l0 = c->launches;
    int r = enqueue_timestep(c, c->fuse_vmax);
    if (!r) r = enqueue_step(c);
    if (!r) { launch_advance_time(c->ctl, c->stream); c->launches++; }
    const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    c->graph_launches = c->launches - l0;
    c->launches = l0;
    if (r) { if (g) cudaGraphDestroy(g); return r; }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(WENO5_ERR_CUDA, "stream capture of one step failed: %s", cudaGetErrorString(e)); }
    const cudaError_t ei = cudaGraphInstantiate(&c->step_graph, g, 0);
    cudaGraphDestroy(g);
    if (ei != cudaSuccess) { cudaGetLastError(); c->step_graph = nullptr; return fail(WENO5_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ei)); }
    c->graph_key = key;
    return 0;
}

extern "C" int weno5_advance(weno5_ctx *c, int nsteps, double tmax, double *t, int *nsteps_done)
{
    if (!c) return fail(WENO5_ERR_ARG, "null argument");
    if (!c->uploaded) return fail(WENO5_ERR_STATE, "no state on the device yet");
    if (nsteps < 0) return fail(WENO5_ERR_ARG, "nsteps < 0");
    DevGuard dg_; CU(dg_.set(c->device));
    double *h = c->h_ctl + 8;
    h[0] = 0.0; h[1] = t ? *t : 0.0; h[2] = 0.0; h[3] = 0.0; h[4] = tmax > 0.0 ? tmax : 0.0; h[5] = 0.0; h[6] = 0.0;
    CU(cudaMemcpyAsync(c->ctl, h, 7 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (use_line1d(c) && nsteps > 0) {
        // (:51-71 of Weno5_1D.jl: vmax -> dt -> classical_RK4_step, until nstep / tmax) inside ONE kernel
        CU(launch_line1d(c->Q0.q, c->ctl, c->g, c->side, nsteps, 0.0, c->P.dx, c->P.gamma, c->P.eps, c->P.cfl, c->stream));
        c->launches++;
        nsteps = 0;
    }
    // from the second step on the CFL reduction rides on stage 4 of the step before (2-D, E-only path)
    c->fuse_vmax = !c->dual && !c->g.is1d && nsteps > 1 && !getenv("WENO5_NO_FUSED_VMAX");
    if (c->fuse_vmax) CU(cudaMemsetAsync(c->vmax_next, 0, 2 * sizeof(unsigned long long), c->stream));
    struct Unfuse { weno5_ctx *c; ~Unfuse() { c->fuse_vmax = false; } } unfuse{c};
    // steps 2.. replay a captured graph (one GPU, no per-launch profiling events)
    const bool use_graph = !c->comm && !c->profiling && nsteps > 2 && !getenv("WENO5_NO_GRAPH");
    for (int s = 0; s < nsteps; ++s) {
        if (use_graph && s > 0) {
            if (int r = build_step_graph(c, (c->fuse_vmax ? 1 : 0) | (c->dual ? 2 : 0) | (c->P.es_roundtrip ? 4 : 0) | (c->dual_flags_only ? 8 : 0))) return r;
            CU(cudaGraphLaunch(c->step_graph, c->stream));
            c->launches += c->graph_launches;
        } else {
            if (int r = enqueue_timestep(c, c->fuse_vmax && s > 0)) return r;
            if (int r = enqueue_step(c)) return r;
            launch_advance_time(c->ctl, c->stream);
            c->launches++;
        }
        if (tmax <= 0.0 && (s % 64) == 63) CU(cudaStreamSynchronize(c->stream));   // keep the launch queue bounded
        if (tmax > 0.0 && (s % 16) == 15) {               // look at the clock now and then
            CU(cudaMemcpyAsync(c->h_ctl, c->ctl, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            if (c->h_ctl[1] >= tmax || c->h_ctl[6] != 0.0) break;
        }
    }
    if (int r = wait_exchange(c)) return r;
    CU(cudaMemcpyAsync(c->h_ctl, c->ctl, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    if (t) *t = c->h_ctl[1];
    if (nsteps_done) *nsteps_done = (int)c->h_ctl[5];
    if (c->h_ctl[6] != 0.0) return fail(WENO5_ERR_NUMERIC, "non-finite wave speed / time step at t=%g", c->h_ctl[1]);
    return 0;
}

// ------------------------------------------------------------------ dual-energy mode
extern "C" int weno5_set_es_switch(weno5_ctx *c, int enable, double threshold)
{
    if (!c) return fail(WENO5_ERR_ARG, "null argument");
    if (c->step_graph) { cudaGraphExecDestroy(c->step_graph); c->step_graph = nullptr; c->graph_key = -1; }   // threshold is baked in
    if (!enable) { c->dual = false; return 0; }
    if (enable != 1 && enable != 2) return fail(WENO5_ERR_ARG, "enable must be 0 (off), 1 (live switch) or 2 (flags only)");
    if (enable == 2 && !c->g.is1d) return fail(WENO5_ERR_ARG, "flags-only mode is the 1-D reference's behaviour (Weno5_1D.jl:252-254); 2-D grids take 0 or 1");
    if (!(threshold >= 0.0)) return fail(WENO5_ERR_ARG, "threshold must be >= 0 (the reference uses 1e-5)");
    DevGuard dg_; CU(dg_.set(c->device));
    const Geom &g = c->g;
    if (!c->dual_pool) {
        const size_t n = 8 + 8 + 6 + 4;
        if (cudaMalloc(&c->dual_pool, n * g.plane * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            return fail(WENO5_ERR_CUDA, "cudaMalloc of %.2f GB for the dual-energy planes failed", n * g.plane * 8.0 / 1e9);
        }
        CU(cudaMemsetAsync(c->dual_pool, 0, n * g.plane * sizeof(double), c->stream));
        double *cur = c->dual_pool;
        for (int k = 0; k < 8; ++k) { c->candE[k] = cur; cur += g.plane; }
        for (int k = 0; k < 8; ++k) { c->candS[k] = cur; cur += g.plane; }
        for (int k = 0; k < 6; ++k) { c->accS[k] = cur; cur += g.plane; }
        c->fsyS = cur; cur += g.plane;
        c->gsxS = cur; cur += g.plane;
        c->tbx = cur; cur += g.plane;
        c->tby = cur; cur += g.plane;
    }
    c->dual = true;
    c->dual_flags_only = enable == 2;
    c->es_threshold = threshold;
    return 0;
}

// the flags of the last stage in the public array shape (Nx,Ny) / (N): 1.0 where |S(E) - S| >= threshold
// (the reference's `idx`, Weno5_1D.jl:245, as a mask)
extern "C" int weno5_es_switch_mask(weno5_ctx *c, double *mask)
{
    if (!c || !mask) return fail(WENO5_ERR_ARG, "null argument");
    if (!c->dual) return fail(WENO5_ERR_STATE, "the dual-energy switch is off");
    DevGuard dg_; CU(dg_.set(c->device));
    if (int r = copy_plane(c, c->flag, nullptr, mask)) return r;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int weno5_es_switch_count(weno5_ctx *c, unsigned long long *cells)
{
    if (!c || !cells) return fail(WENO5_ERR_ARG, "null argument");
    if (!c->dual) return fail(WENO5_ERR_STATE, "the dual-energy switch is off");
    DevGuard dg_; CU(dg_.set(c->device));
    const double *pl[1] = {c->flag};
    launch_totals(pl, 1, c->g, c->sums, c->stream);
    c->launches++;
    CU(cudaMemcpyAsync(c->h_ctl, c->sums, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *cells = (unsigned long long)(c->h_ctl[0] + 0.5);
    return 0;
}

// ------------------------------------------------------------------ diagnostics
extern "C" int weno5_divb(weno5_ctx *c, double *divb)
{
    if (!c || !divb) return fail(WENO5_ERR_ARG, "null argument");
    if (c->g.is1d) return fail(WENO5_ERR_ARG, "divB is a 2-D diagnostic");
    if (!c->uploaded) return fail(WENO5_ERR_STATE, "no state on the device yet");
    DevGuard dg_; CU(dg_.set(c->device));
    launch_divb(c->Q0.bx, c->Q0.by, c->scratch, c->g, c->P.dx, c->P.dy, c->stream);
    c->launches++;
    if (int r = copy_plane(c, c->scratch, nullptr, divb)) return r;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int weno5_totals(weno5_ctx *c, double sums[11])
{
    if (!c || !sums) return fail(WENO5_ERR_ARG, "null argument");
    if (!c->uploaded) return fail(WENO5_ERR_STATE, "no state on the device yet");
    DevGuard dg_; CU(dg_.set(c->device));
    const double *pl[11];
    for (int k = 0; k < NCELL; ++k) pl[host_comp(k)] = c->Q0.q[k];     // report in the reference's order (S, then E)
    pl[9] = c->Q0.bx; pl[10] = c->Q0.by;
    launch_totals(pl, c->g.is1d ? 9 : 11, c->g, c->sums, c->stream);
    c->launches++;
    CU(cudaMemcpyAsync(c->h_ctl, c->sums, 11 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 11; ++k) sums[k] = (c->g.is1d && k >= 9) ? 0.0 : c->h_ctl[k];
    return 0;
}

// ------------------------------------------------------------------ snapshot hook
extern "C" int weno5_arr2img(weno5_ctx *c, int component, double lo, double hi, int log_scale, int ncolours, short *rimg,
                             double range_used[2])
{
    if (!c || !rimg) return fail(WENO5_ERR_ARG, "null argument");
    if (!c->uploaded) return fail(WENO5_ERR_STATE, "no state on the device yet");
    if (ncolours < 2 || ncolours > 32767) return fail(WENO5_ERR_ARG, "ncolours must be in 2..32767 (Int16 indices)");
    const Geom &g = c->g;
    if (component < 0 || component > 11 || (g.is1d && component > 8)) return fail(WENO5_ERR_ARG, "unknown component %d", component);
    DevGuard dg_; CU(dg_.set(c->device));
    const double *field;
    if (component <= 8) {
        int k = component;                         // reference order: ..., 8th = S, 9th = E
        if (component == 7) k = C_S;
        if (component == 8) k = C_E;
        field = c->Q0.q[k];
    } else if (component == 9) field = c->Q0.bx;
    else if (component == 10) field = c->Q0.by;
    else {
        launch_divb(c->Q0.bx, c->Q0.by, c->scratch, g, c->P.dx, c->P.dy, c->stream);
        c->launches++;
        field = c->scratch;
    }
    short *img = reinterpret_cast<short *>(c->rhs[1]);               // free between steps
    launch_arr2img(field, g, lo, hi, log_scale ? 1 : 0, ncolours, c->vmax_bits_img, img, c->sums, c->stream);
    c->launches += 2;
    const size_t n = (size_t)g.Nx * g.Ny;
    CU(cudaMemcpyAsync(rimg, img, n * sizeof(short), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->h_ctl, c->sums, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    if (range_used) { range_used[0] = c->h_ctl[0]; range_used[1] = c->h_ctl[1]; }
    return 0;
}

// ------------------------------------------------------------------ reference-shaped one-shot calls
extern "C" int weno5_step_host(weno5_ctx *c, const double *q0, const double *bxb0, const double *byb0, double dt,
                               double *q4, double *bxb4, double *byb4)
{
    if (!c || !q0 || !q4) return fail(WENO5_ERR_ARG, "null argument");
    if (c->g.is1d) {
        if (int r = weno5_upload_1d(c, q0)) return r;
        if (int r = weno5_rk4_step(c, dt)) return r;
        return weno5_download_1d(c, q4);
    }
    if (int r = weno5_upload(c, q0, bxb0, byb0)) return r;
    if (int r = weno5_rk4_step(c, dt)) return r;
    return weno5_download(c, q4, bxb4, byb4);
}

extern "C" int weno5_classical_rk4_step_2d(const weno5_params *P, const double *q0, const double *bxb0,
                                           const double *byb0, double dt, double *q4, double *bxb4, double *byb4)
{
    if (!P) return fail(WENO5_ERR_ARG, "null argument");
    if (P->ny == 1) return fail(WENO5_ERR_ARG, "use weno5_classical_rk4_step_1d for ny = 1");
    weno5_ctx *c = nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    if (int r = weno5_create(P, dev, P->ny, 0, &c)) return r;
    const int r = weno5_step_host(c, q0, bxb0, byb0, dt, q4, bxb4, byb4);
    weno5_destroy(c);
    return r;
}

extern "C" int weno5_classical_rk4_step_1d(const weno5_params *P, const double *q0, double dt, double *q4)
{
    if (!P) return fail(WENO5_ERR_ARG, "null argument");
    weno5_params P1 = *P;
    P1.ny = 1;
    weno5_ctx *c = nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    if (int r = weno5_create(&P1, dev, 1, 0, &c)) return r;
    const int r = weno5_step_host(c, q0, nullptr, nullptr, dt, q4, nullptr, nullptr);
    weno5_destroy(c);
    return r;
}

// ------------------------------------------------------------------ pipelined host -> host step
// classical_RK4_step for host-resident arrays, streamed through the GPU in y-slabs.  One explicit RK4
// step has a finite domain of dependence: a row of the result depends on the 4 stages x 4 rows of the
// input around it (3 for the WENO stencil of the y faces + 1 for the constrained-transport
// centring).  A slab that carries PIPE_H >= 16 extra rows on each cut side therefore reproduces its
// owned rows bit for bit without talking to its neighbours.  Three slab contexts on three streams
// rotate, so the H2D copy of slab k+1, the RK4 step of slab k and the D2H copy of slab k-1 overlap:
// the call is bound by the PCIe transfer of the state (once in, once out) instead of
// upload + compute + download in sequence.
// Every row crosses PCIe once: the input arrays are uploaded, row range by row range, on a stream of their own into an
// image of the whole grid on the device, and a slab gathers the rows of its window (its own and the overlap) from that
// image as soon as the ranges it needs have arrived (one event per range).  Windows therefore cost no extra transfer,
// and slabs can be small: the fill and drain of the pipeline -- the upload of the first window and the download of
// the last slab, which nothing overlaps -- shrink with them.  (WENO5_PIPE_DIRECT=1: every slab uploads its whole
// window from the host itself, the first version: 6-8 % more bytes up, and slabs of 512 rows were the best.)
constexpr int PIPE_H = 16;
constexpr int PIPE_NCTX = 3;
constexpr int PIPE_NIN = 10;  // planes of the input image: 8 cell fields (no S) + bxb + byb

struct weno5_pipe {
    weno5_params P;
    int device;
    int R;                    // owned rows per slab
    int W;                    // interior rows of a slab window = R + 2 PIPE_H
    bool last_pinned = true;  // the last weno5_pipe_step got page-locked host arrays (its copies overlapped compute)
    weno5_ctx *ctx[PIPE_NCTX];
    bool direct = false;      // slabs upload their windows themselves (no input image)
    double *inbuf = nullptr;  // input image: PIPE_NIN planes of Nx x Ny doubles (public layout)
    double *outbuf = nullptr; // result image: 11 planes in the host order (q[:,:,1..9], bxb, byb)
    cudaStream_t up = nullptr, down = nullptr;
    std::vector<cudaEvent_t> ev;   // [0]: the wrap-around rows (periodic), [1 + c]: upload range c
    std::vector<cudaEvent_t> done; // [k]: slab k has put its rows into the result image
};

extern "C" int weno5_pipe_overlapped(const weno5_pipe *p) { return p && p->last_pinned ? 1 : 0; }

extern "C" int weno5_pipe_destroy(weno5_pipe *p)
{
    if (!p) return 0;
    for (int k = 0; k < PIPE_NCTX; ++k) weno5_destroy(p->ctx[k]);
    {
        DevGuard dg_;
        dg_.set(p->device);
        for (cudaEvent_t e : p->ev) cudaEventDestroy(e);
        for (cudaEvent_t e : p->done) cudaEventDestroy(e);
        if (p->up) cudaStreamDestroy(p->up);
        if (p->down) cudaStreamDestroy(p->down);
        if (p->inbuf) cudaFree(p->inbuf);
        if (p->outbuf) cudaFree(p->outbuf);
        cudaGetLastError();
    }
    delete p;
    return 0;
}

extern "C" int weno5_pipe_create(const weno5_params *P, int device, int rows_per_slab, weno5_pipe **out)
{
    if (!P || !out) return fail(WENO5_ERR_ARG, "null argument");
    if (P->ny == 1) return fail(WENO5_ERR_ARG, "the pipelined step is 2-D only");
    if (rows_per_slab < 2 * PIPE_H + 8) return fail(WENO5_ERR_ARG, "rows_per_slab must be >= %d", 2 * PIPE_H + 8);
    if (P->ny < rows_per_slab + 2 * PIPE_H)
        return fail(WENO5_ERR_ARG, "grid of %d rows is too small for slabs of %d rows (+%d overlap); use weno5_step_host", P->ny,
                    rows_per_slab, 2 * PIPE_H);
    weno5_pipe *p = new weno5_pipe();
    p->P = *P;
    p->device = device;
    p->R = rows_per_slab;
    p->W = rows_per_slab + 2 * PIPE_H;
    for (int k = 0; k < PIPE_NCTX; ++k) p->ctx[k] = nullptr;
    for (int k = 0; k < PIPE_NCTX; ++k) {
        if (int r = create_ctx(P, device, p->W, 0, SIDE_NONE, SIDE_NONE, &p->ctx[k])) {
            weno5_pipe_destroy(p);
            return r;
        }
    }
    p->direct = getenv("WENO5_PIPE_DIRECT") != nullptr;
    if (!p->direct) {
        DevGuard dg_;
        const size_t hp = (size_t)(P->nx + 2 * WENO5_NBND) * (P->ny + 2 * WENO5_NBND);
        const int nranges = (P->ny + 2 * WENO5_NBND + p->R - 1) / p->R;
        cudaError_t e = dg_.set(device);
        if (e == cudaSuccess) e = cudaMalloc(&p->inbuf, PIPE_NIN * hp * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc(&p->outbuf, 11 * hp * sizeof(double));
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->up, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->down, cudaStreamNonBlocking);
        for (int k = 0; e == cudaSuccess && k < nranges + 1; ++k) {
            cudaEvent_t x;
            e = cudaEventCreateWithFlags(&x, cudaEventDisableTiming);
            if (e == cudaSuccess) p->ev.push_back(x);
        }
        const int nslab_ = (P->ny + p->R - 1) / p->R;
        for (int k = 0; e == cudaSuccess && k < nslab_; ++k) {
            cudaEvent_t x;
            e = cudaEventCreateWithFlags(&x, cudaEventDisableTiming);
            if (e == cudaSuccess) p->done.push_back(x);
        }
        if (e != cudaSuccess) {
            weno5_pipe_destroy(p);
            return fail(WENO5_ERR_CUDA, "pipelined step: %s", cudaGetErrorString(e));
        }
    }
    *out = p;
    return 0;
}

// copy `nrows` rows starting at host public row hj <-> internal row ij of one plane
static int copy_rows(weno5_ctx *c, double *dev, const double *host_in, double *host_out, int Ny_host, int hj, int ij, int nrows)
{
    if (nrows <= 0) return 0;
    const Geom &g = c->g;
    (void)Ny_host;
    double *d = dev + 1 + (size_t)g.pitch * ij;
    const size_t w = (size_t)g.Nx * sizeof(double);
    if (host_in)
        CU(cudaMemcpy2DAsync(d, g.pitch * sizeof(double), host_in + (size_t)g.Nx * hj, w, w, nrows, cudaMemcpyHostToDevice, c->stream));
    if (host_out)
        CU(cudaMemcpy2DAsync(host_out + (size_t)g.Nx * hj, w, d, g.pitch * sizeof(double), w, nrows, cudaMemcpyDeviceToHost, c->stream));
    return 0;
}

extern "C" int weno5_pipe_step(weno5_pipe *p, const double *q0, const double *bxb0, const double *byb0, double dt, double *q4,
                               double *bxb4, double *byb4, double *dt_next)
{
    if (!p || !q0 || !bxb0 || !byb0 || !q4 || !bxb4 || !byb4) return fail(WENO5_ERR_ARG, "null argument");
    {
        // inputs are read while outputs of earlier slabs are being written: no byte of an output may lie in an input
        const weno5_params &Q = p->P;
        const size_t np_ = (size_t)(Q.nx + 2 * WENO5_NBND) * (Q.ny + 2 * WENO5_NBND) * sizeof(double);
        const char *in[3] = {(const char *)q0, (const char *)bxb0, (const char *)byb0};
        const char *ou[3] = {(const char *)q4, (const char *)bxb4, (const char *)byb4};
        const size_t sz[3] = {9 * np_, np_, np_};
        for (int i = 0; i < 3; ++i)
            for (int o = 0; o < 3; ++o)
                if (in[i] < ou[o] + sz[o] && ou[o] < in[i] + sz[i])
                    return fail(WENO5_ERR_ARG, "the pipelined step cannot run in place: an output array overlaps an input array");
        for (int a = 0; a < 3; ++a)
            for (int b = a + 1; b < 3; ++b)
                if (ou[a] < ou[b] + sz[b] && ou[b] < ou[a] + sz[a]) return fail(WENO5_ERR_ARG, "output arrays overlap each other");
        // copies from/to pageable memory are staged by the driver and do not overlap compute: note it for the caller
        cudaPointerAttributes at;
        bool pinned = true;
        for (const void *ptr : {(const void *)q0, (const void *)q4})
            if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess || at.type != cudaMemoryTypeHost) { cudaGetLastError(); pinned = false; }
        p->last_pinned = pinned;
    }
    if (!(dt > 0.0) || !std::isfinite(dt)) return fail(WENO5_ERR_ARG, "dt must be positive and finite (got %g)", dt);
    DevGuard dg_; CU(dg_.set(p->device));
    const weno5_params &P = p->P;
    const int ny = P.ny, Ny = ny + 2 * WENO5_NBND, Nx = P.nx + 2 * WENO5_NBND;
    const bool periodic = P.bc_y == WENO5_BC_PERIODIC;
    const size_t hp = (size_t)Nx * Ny;
    const int nslab = (ny + p->R - 1) / p->R;

    for (int k = 0; k < PIPE_NCTX; ++k) {
        weno5_ctx *c = p->ctx[k];
        c->h_ctl[8] = dt;
        CU(cudaMemcpyAsync(c->ctl, c->h_ctl + 8, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemsetAsync(c->vmax_bits, 0, 2 * sizeof(unsigned long long), c->stream));
    }

    // ---- input image: host planes in the order of the slab planes (8 cell fields without S, bxb, byb)
    const double *hsrc[PIPE_NIN];
    const double *isrc[PIPE_NIN];
    {
        int n = 0;
        for (int f = 0; f < NCELL; ++f)
            if (f != C_S) { hsrc[n] = q0 + host_comp(f) * hp; isrc[n] = p->inbuf + (size_t)n * hp; ++n; }
        hsrc[n] = bxb0; isrc[n] = p->inbuf + (size_t)n * hp; ++n;
        hsrc[n] = byb0; isrc[n] = p->inbuf + (size_t)n * hp; ++n;
    }
    const int nranges = (Ny + p->R - 1) / p->R;
    const int T0 = periodic ? ny + WENO5_NBND - (PIPE_H + 8) : Ny, T1 = periodic ? ny + WENO5_NBND : Ny;   // wrap-around rows
    int next_range = 0;
    // public rows [a, b) of every plane, then the event.  The planes of q are hp doubles apart on both sides, so the
    // same row range of several planes is ONE 2-D copy ("rows" = planes): host planes 0..6 (rho .. Bz) -> image planes
    // 0..6, host plane 8 (E) -> image plane 7, bxb, byb: four copies per range instead of ten
    auto upload_range = [&](int a, int b, cudaEvent_t ev) -> int {
        const size_t w = (size_t)Nx * (b - a) * sizeof(double), off = (size_t)Nx * a, pitch = hp * sizeof(double);
        const bool planes2d = pitch < ((size_t)1 << 31);                  // cudaMemcpy2D pitch limit (16384^2 planes exceed it)
        if (planes2d) CU(cudaMemcpy2DAsync(p->inbuf + off, pitch, q0 + off, pitch, w, 7, cudaMemcpyHostToDevice, p->up));
        for (int n = planes2d ? 7 : 0; n < PIPE_NIN; ++n)
            CU(cudaMemcpyAsync(p->inbuf + (size_t)n * hp + off, hsrc[n] + off, w, cudaMemcpyHostToDevice, p->up));
        CU(cudaEventRecord(ev, p->up));
        return 0;
    };
    if (!p->direct && periodic)
        if (int r = upload_range(T0, T1, p->ev[0])) return r;           // the first slab's window starts behind the last row
    // rows [a, b) of the image are needed on stream s: enqueue the uploads up to them, wait for their events
    auto need_rows = [&](int a, int b, cudaStream_t s) -> int {
        if (a >= T0 && b <= T1) { CU(cudaStreamWaitEvent(s, p->ev[0], 0)); return 0; }
        const int c1 = (b - 1) / p->R;
        for (; next_range <= c1 && next_range < nranges; ++next_range) {
            const int ra = next_range * p->R, rb = ra + p->R < Ny ? ra + p->R : Ny;
            if (int r = upload_range(ra, rb, p->ev[1 + next_range])) return r;
        }
        for (int c = a / p->R; c <= c1; ++c) CU(cudaStreamWaitEvent(s, p->ev[1 + c], 0));
        return 0;
    };

    for (int k = 0; k < nslab; ++k) {
        weno5_ctx *c = p->ctx[k % PIPE_NCTX];
        const Geom &g = c->g;
        const int r0 = k * p->R, r1 = r0 + p->R < ny ? r0 + p->R : ny;       // owned interior rows
        // window of W interior rows around the owned rows
        int w0 = r0 - PIPE_H;
        if (!periodic) { if (w0 < 0) w0 = 0; if (w0 > ny - p->W) w0 = ny - p->W; }
        const bool phys_lo = !periodic && w0 == 0, phys_hi = !periodic && w0 + p->W == ny;
        c->side[2] = phys_lo ? SIDE_OUTFLOW : SIDE_NONE;
        c->side[3] = phys_hi ? SIDE_OUTFLOW : SIDE_NONE;

        // ---- upload: internal row ii holds interior row w0 - 4 + ii, i.e. public row w0 - 1 + ii
        for (int ii = 0; ii < g.NYI;) {
            int src = w0 - 1 + ii;                        // public row
            int run;
            if (periodic) {
                int rr = ((w0 - 4 + ii) % ny + ny) % ny;   // interior row, wrapped
                src = rr + WENO5_NBND;
                run = ny - rr;                             // rows until the wrap
            } else {
                if (src < 0) { ++ii; continue; }           // beyond the public ghosts: filled by bc_fill
                if (src >= Ny) break;
                run = Ny - src;
            }
            if (run > g.NYI - ii) run = g.NYI - ii;
            // S (q0[:,:,8]) is an output of the E-branch step, never an input (ES_Switch derives it from E,
            // Weno5_2D.jl:1738): its plane is not sent over PCIe
            if (p->direct) {
                for (int f = 0; f < NCELL; ++f)
                    if (f != C_S)
                        if (int r = copy_rows(c, c->Q0.q[f], q0 + host_comp(f) * hp, nullptr, Ny, src, ii, run)) return r;
                if (int r = copy_rows(c, c->Q0.bx, bxb0, nullptr, Ny, src, ii, run)) return r;
                if (int r = copy_rows(c, c->Q0.by, byb0, nullptr, Ny, src, ii, run)) return r;
            } else {
                if (int r = need_rows(src, src + run, c->stream)) return r;
                const double *sp[PIPE_NIN];
                double *dp[PIPE_NIN];
                int n = 0;
                for (int f = 0; f < NCELL; ++f)
                    if (f != C_S) { dp[n] = c->Q0.q[f] + 1 + (size_t)g.pitch * ii; ++n; }
                dp[n++] = c->Q0.bx + 1 + (size_t)g.pitch * ii;
                dp[n++] = c->Q0.by + 1 + (size_t)g.pitch * ii;
                for (int m = 0; m < PIPE_NIN; ++m) sp[m] = isrc[m] + (size_t)Nx * src;
                launch_gather_rows(sp, dp, PIPE_NIN, g.Nx, Nx, g.pitch, run, c->stream);
                c->launches++;
            }
            ii += run;
        }
        c->uploaded = true;
        if (int r = apply_boundaries(c, c->Q0, true)) return r;

        // ---- one RK4 step of the window, wave speeds of the owned rows for the next dt
        if (int r = enqueue_step(c)) return r;
        const int io0 = r0 - w0 + G, io1 = r1 - w0 + G;       // owned rows, internal
        launch_vmax_rows(c->Q0.q, g, P.gamma, c->vmax_bits, io0, io1, 0, c->stream);
        c->launches++;

        // ---- download the owned rows (and the physical ghost rows next to them)
        int d0 = io0, d1 = io1;
        if (phys_lo && r0 == 0) d0 = 1;                        // public ghost rows 0..2 = internal 1..3
        if (phys_hi && r1 == ny) d1 = g.NYI - 1;
        const int hj = w0 - 1 + d0;                            // public row of internal row d0
        if (p->direct) {
            for (int f = 0; f < NCELL; ++f)
                if (int r = copy_rows(c, c->Q0.q[f], nullptr, q4 + host_comp(f) * hp, Ny, hj, d0, d1 - d0)) return r;
            if (int r = copy_rows(c, c->Q0.bx, nullptr, bxb4, Ny, hj, d0, d1 - d0)) return r;
            if (int r = copy_rows(c, c->Q0.by, nullptr, byb4, Ny, hj, d0, d1 - d0)) return r;
        } else {
            // rows -> result image (contiguous, host order); the download stream takes them from there, so this
            // slab's stream is free for its next window while they travel
            const double *sp[11];
            double *dp[11];
            for (int f = 0; f < NCELL; ++f) {
                sp[host_comp(f)] = c->Q0.q[f] + 1 + (size_t)g.pitch * d0;
                dp[host_comp(f)] = p->outbuf + (size_t)host_comp(f) * hp + (size_t)Nx * hj;
            }
            sp[9] = c->Q0.bx + 1 + (size_t)g.pitch * d0; dp[9] = p->outbuf + 9 * hp + (size_t)Nx * hj;
            sp[10] = c->Q0.by + 1 + (size_t)g.pitch * d0; dp[10] = p->outbuf + 10 * hp + (size_t)Nx * hj;
            launch_gather_rows(sp, dp, 11, g.Nx, g.pitch, Nx, d1 - d0, c->stream);
            c->launches++;
            CU(cudaEventRecord(p->done[k], c->stream));
            CU(cudaStreamWaitEvent(p->down, p->done[k], 0));
            const size_t nb = (size_t)Nx * (d1 - d0) * sizeof(double), off = (size_t)Nx * hj;
            // the nine planes of q4 are hp doubles apart on both sides: one 2-D copy whose "rows" are the planes
            if (hp * sizeof(double) < ((size_t)1 << 31)) {
                CU(cudaMemcpy2DAsync(q4 + off, hp * sizeof(double), p->outbuf + off, hp * sizeof(double), nb, 9, cudaMemcpyDeviceToHost, p->down));
            } else {
                for (int f = 0; f < 9; ++f)
                    CU(cudaMemcpyAsync(q4 + (size_t)f * hp + off, p->outbuf + (size_t)f * hp + off, nb, cudaMemcpyDeviceToHost, p->down));
            }
            CU(cudaMemcpyAsync(bxb4 + off, p->outbuf + 9 * hp + off, nb, cudaMemcpyDeviceToHost, p->down));
            CU(cudaMemcpyAsync(byb4 + off, p->outbuf + 10 * hp + off, nb, cudaMemcpyDeviceToHost, p->down));
        }
    }

    if (!p->direct) { CU(cudaStreamSynchronize(p->up)); CU(cudaStreamSynchronize(p->down)); }
    unsigned long long vb[2] = {0ull, 0ull};
    for (int k = 0; k < PIPE_NCTX; ++k) {
        weno5_ctx *c = p->ctx[k];
        unsigned long long *hv = reinterpret_cast<unsigned long long *>(c->h_ctl + 12);
        CU(cudaMemcpyAsync(hv, c->vmax_bits, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        if (hv[0] > vb[0]) vb[0] = hv[0];
        if (hv[1] > vb[1]) vb[1] = hv[1];
    }
    if (periodic) {
        // boundaries! of the result (Weno5_2D.jl:313, periodic addition): ghost rows are images
        double *arrs[3] = {q4, bxb4, byb4};
        const int nplanes[3] = {NCELL, 1, 1};
        for (int a = 0; a < 3; ++a)
            for (int f = 0; f < nplanes[a]; ++f) {
                double *pl = arrs[a] + (size_t)f * hp;
                memcpy(pl, pl + (size_t)Nx * ny, sizeof(double) * Nx * WENO5_NBND);
                memcpy(pl + (size_t)Nx * (ny + WENO5_NBND), pl + (size_t)Nx * WENO5_NBND, sizeof(double) * Nx * WENO5_NBND);
            }
    }
    if (dt_next) {
        double vx, vy;
        memcpy(&vx, &vb[0], 8);
        memcpy(&vy, &vb[1], 8);
        const double dtx = P.dx / vx, dty = P.dy / vy;               // timestep(), Weno5_2D.jl:1930-1940
        *dt_next = P.cfl / (1 / dtx + 1 / dty);
        if (!(*dt_next > 0.0) || !std::isfinite(*dt_next))
            return fail(WENO5_ERR_NUMERIC, "non-finite wave speed in the result (vxmax=%g vymax=%g)", vx, vy);
    }
    return 0;
}

// ------------------------------------------------------------------ host memory, instrumentation
extern "C" int weno5_host_alloc(void **ptr, unsigned long long bytes)
{
    if (!ptr) return fail(WENO5_ERR_ARG, "null argument");
    CU(cudaMallocHost(ptr, bytes));
    return 0;
}

extern "C" int weno5_host_free(void *ptr)
{
    CU(cudaFreeHost(ptr));
    return 0;
}

extern "C" unsigned long long weno5_launch_count(const weno5_ctx *c) { return c ? c->launches : 0; }

extern "C" int weno5_profile(weno5_ctx *c, int enable)
{
    if (!c) return fail(WENO5_ERR_ARG, "null argument");
    c->profiling = enable != 0;
    return 0;
}

extern "C" int weno5_profile_read(weno5_ctx *c, double *flux_ms, unsigned long long *flux_launches, int reset)
{
    if (!c) return fail(WENO5_ERR_ARG, "null argument");
    DevGuard dg_; CU(dg_.set(c->device));
    CU(cudaStreamSynchronize(c->stream));
    for (size_t k = 0; k < c->ev_used; ++k) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, c->ev[k].first, c->ev[k].second));
        c->prof_ms += ms;
    }
    c->ev_used = 0;
    if (flux_ms) *flux_ms = c->prof_ms;
    if (flux_launches) *flux_launches = c->prof_launches;
    if (reset) { c->prof_ms = 0.0; c->prof_launches = 0; }
    return 0;
}

extern "C" void *weno5_stream(weno5_ctx *c) { return c ? (void *)c->stream : nullptr; }

extern "C" int weno5_debug_fetch(weno5_ctx *c, const char *name, double *out)
{
    if (!c || !name || !out) return fail(WENO5_ERR_ARG, "null argument");
    DevGuard dg_; CU(dg_.set(c->device));
    const std::string n(name);
    double *src = nullptr;
    if (n == "fsy") src = c->fsy;
    else if (n == "gsx") src = c->gsx;
    else if (n.rfind("rhs", 0) == 0 && n.size() == 4 && n[3] >= '0' && n[3] <= '6') src = c->rhs[n[3] - '0'];
    else if (n.rfind("qa", 0) == 0 && n.size() == 3 && n[2] >= '0' && n[2] <= '8') { int hk = n[2] - '0'; src = c->QA.q[hk == 7 ? C_S : (hk == 8 ? C_E : hk)]; }
    else if (n.rfind("qb", 0) == 0 && n.size() == 3 && n[2] >= '0' && n[2] <= '8') { int hk = n[2] - '0'; src = c->QB.q[hk == 7 ? C_S : (hk == 8 ? C_E : hk)]; }
    else if (n == "qa_bx") src = c->QA.bx;
    else if (n == "qa_by") src = c->QA.by;
    else return fail(WENO5_ERR_ARG, "unknown field '%s'", name);
    if (int r = copy_plane(c, src, nullptr, out)) return r;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}
