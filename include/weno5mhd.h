/*
 * libweno5mhd -- C ABI of the B200 (sm_100a) WENO5 ideal-MHD solver.
 *
 * Drop-in boundary for the hot path of jdonnert/julia's Weno5_2D.jl / Weno5_1D.jl.  The
 * reference has no FFI: the path sits behind plain Julia functions.  Each entry point
 * below names the Julia function (file:line under /root/reference) whose work it does;
 * julia_b200/julia/Weno5GPU.jl re-exports those names over ccall.
 *
 * Conventions
 *  - plain C, no exceptions cross the boundary; every call returns 0 on success or a
 *    negative weno5_status, and weno5_last_error() gives a thread-local message;
 *  - host arrays are exactly the reference's Julia arrays (column-major Float64):
 *      q   (Nx,Ny,9)  [rho,Mx,My,Mz,Bx,By,Bz,S,E]      Weno5_2D.jl:35,2006
 *      bxb (Nx,Ny)    Bx on the x-face i+1/2             Weno5_2D.jl:37,1669
 *      byb (Nx,Ny)    By on the y-face j+1/2
 *      1-D: q (N,9)                                      Weno5_1D.jl:37,1101
 *    with Nx = nx + 2*3, Ny = ny + 2*3 (NBnd = 3, Weno5_2D.jl:9-17);
 *  - the caller owns host buffers, the library owns device memory inside the handle;
 *  - one host thread per handle; handles are independent;
 *  - there is no CPU fallback: every call fails with WENO5_ERR_CUDA if no sm_100 device
 *    is usable.
 */
#ifndef WENO5MHD_H
#define WENO5MHD_H

#ifdef __cplusplus
extern "C" {
#endif

#define WENO5_NBND 3

typedef enum {
    WENO5_OK = 0,
    WENO5_ERR_ARG = -1,      /* bad argument / unsupported configuration */
    WENO5_ERR_CUDA = -2,     /* CUDA runtime error or no usable device   */
    WENO5_ERR_NCCL = -3,     /* NCCL missing or failed                   */
    WENO5_ERR_STATE = -4,    /* call sequence error (e.g. step before upload) */
    WENO5_ERR_NUMERIC = -5   /* non-finite wave speed / time step        */
} weno5_status;

enum { WENO5_BC_OUTFLOW = 0,   /* zero-gradient, the reference's only BC (Weno5_2D.jl:1684-1732) */
       WENO5_BC_PERIODIC = 1 };

/* The reference's module-level constants (Weno5_2D.jl:7-29, Weno5_1D.jl:9-29) as a struct. */
typedef struct {
    int nx, ny;              /* interior cells of the GLOBAL grid; ny = 1 selects the 1-D solver */
    double dx, dy;           /* cell sizes (dy ignored in 1-D)           */
    double gamma;            /* gam     = 5/3                             */
    double cfl;              /* courFac = 0.8                             */
    double eps;              /* eps     = 1e-6 (WENO)                     */
    int bc_x, bc_y;          /* WENO5_BC_*                                */
    int es_roundtrip;        /* 1: E <- S2E(E2S(E)) after every stage as the 2-D reference does
                                (Weno5_2D.jl:1738,1765); 0: keep E (as the 1-D reference does,
                                Weno5_1D.jl:252-254; round-off-level difference, faster)   */
} weno5_params;

typedef struct weno5_ctx weno5_ctx;

const char *weno5_last_error(void);
int weno5_version(void);
/* number of usable sm_100 devices (0 if none) */
int weno5_device_count(void);

/* Create a solver on CUDA device `device` for the slab [j_offset, j_offset + ny_local) of the
 * global grid (single GPU: j_offset = 0, ny_local = params->ny). */
int weno5_create(const weno5_params *params, int device, int ny_local, int j_offset, weno5_ctx **out);
int weno5_destroy(weno5_ctx *ctx);

/* ---- multi-GPU: one handle per rank, y-slabs, NCCL send/recv of ghost rows per RK stage and
 * an NCCL max-allreduce for the time step.  The 128-byte id is produced on rank 0 and
 * distributed by the host program (Julia Distributed / MPI / torch.distributed). */
int weno5_comm_unique_id(unsigned char id[128]);
int weno5_comm_init(weno5_ctx *ctx, int rank, int nranks, const unsigned char id[128]);
/* How the ranks talk: 0 no communicator; 1 NCCL send/recv of packed ghost-row messages and an NCCL max-all-reduce of
 * the wave speeds (timestep, Weno5_2D.jl:1930); 2 ghost rows stored straight into the neighbours' planes over peer
 * memory (CUDA IPC, one process per GPU on one node), all-reduce still NCCL; 3 (the default whenever every rank can
 * map the others) additionally the all-reduce as a peer-memory exchange inside the dt kernel.  WENO5_HALO=nccl in the
 * environment forces 1, WENO5_ALLREDUCE=nccl forces 2. */
int weno5_comm_halo_mode(const weno5_ctx *ctx);
/* weno5_destroy of a handle whose memory other ranks map (modes 2, 3) first unmaps what it mapped itself and then
 * waits -- at most 2 s -- until the other ranks have unmapped it: destroy the handles of all ranks around the same
 * time, like the communicator they share. */

/* ---- state transfer.  Host arrays cover the LOCAL slab: (Nx, ny_local+6, 9) etc. */
/* q/bxb/byb as produced by initial_conditions() (Weno5_2D.jl:2001) and
 * interpolateB_center2face() (:1669); bxb == NULL derives the face fields on the device with
 * that same 4th-order interpolation.  Applies boundaries! (:39). */
int weno5_upload(weno5_ctx *ctx, const double *q, const double *bxb, const double *byb);
int weno5_download(weno5_ctx *ctx, double *q, double *bxb, double *byb);
/* 1-D: q is (N,9) */
int weno5_upload_1d(weno5_ctx *ctx, const double *q);
int weno5_download_1d(weno5_ctx *ctx, double *q);

/* ---- the hot path */
/* boundaries!(q, bxb, byb)                    Weno5_2D.jl:1684 ; boundaries!(q) Weno5_1D.jl:1076 */
int weno5_boundaries(weno5_ctx *ctx);
/* timestep(q) -> dt, vxmax, vymax             Weno5_2D.jl:1930,1942 ; 1-D: compute_global_vmax
 * Weno5_1D.jl:1038 and dt = courFac*dx/vmax (:55), vymax = 0.  Global over all ranks. */
int weno5_timestep(weno5_ctx *ctx, double *dt, double *vxmax, double *vymax);
/* classical_RK4_step(q,bxb,byb,dt) in place   Weno5_2D.jl:125 ; 1-D (q,dt) Weno5_1D.jl:152 */
int weno5_rk4_step(weno5_ctx *ctx, double dt);
/* the driver loop of WENO5_2D()/WENO5_1D() (Weno5_2D.jl:50-72, Weno5_1D.jl:51-71) kept on the
 * device: repeat {dt = timestep(q); classical_RK4_step} nsteps times, or until t >= tmax when
 * tmax > 0 (the last step is shortened to land on tmax).  *t is read and updated. */
int weno5_advance(weno5_ctx *ctx, int nsteps, double tmax, double *t, int *nsteps_done);
/* compute_divB(bxb,byb) -> divB (Nx,Ny_local+6)          Weno5_2D.jl:1986 */
int weno5_divb(weno5_ctx *ctx, double *divb);
/* interior sums of the 9 components of q and of bxb, byb (conservation diagnostics,
 * cf. the per-component max/mean print at Weno5_2D.jl:69-71); local to this rank */
int weno5_totals(weno5_ctx *ctx, double sums[11]);

/* ---- dual-energy mode (SURVEY "next" row f3).  The shipped reference computes both the E sweeps
 * (weno5_flux_difference_E, Weno5_2D.jl:973) and the entropy-variable S sweeps (weno5_flux_difference_S, :361)
 * every stage and then discards the S state, because ES_Switch (:1734-1768) has its criterion commented out
 * (`dS = 1`, :1741).  With enable = 1 the step runs both branches and applies the criterion the reference
 * wrote and disabled (:1740): cells whose entropy derived from the E state differs from the advected entropy by
 * >= threshold (reference value 1e-5) take the E state, all others the S state, and the constrained-transport
 * seed fluxes follow the selection (:1754-1758).  enable = 0 (the default) is the shipped 2-D behaviour.
 * Works on y-slabs too (the flags of the neighbours' edge rows travel with the halo).
 * 1-D grids (Weno5_1D.jl:152-258): enable = 2 is the 1-D reference AS SHIPPED -- both branches are swept, the state
 * stays the energy branch's (:252-254) and the flags are the output (`idx`, :245); enable = 1 drops that override
 * (entropy state by default, energy state in flagged cells).
 * weno5_es_switch_count: interior cells of this slab flagged in the most recent stage; weno5_es_switch_mask: the
 * flags themselves in the public array shape (Nx,Ny) / (N), 1.0 = flagged. */
int weno5_set_es_switch(weno5_ctx *ctx, int enable, double threshold);
int weno5_es_switch_count(weno5_ctx *ctx, unsigned long long *cells);
int weno5_es_switch_mask(weno5_ctx *ctx, double *mask);

/* ---- snapshot hook (SURVEY "next" row f4).  Arr2Img(arr, cmap; range, log) of JColorMaps.jl:92-146, which the
 * reference driver calls on q[:,:,k] and divB for its frames (Weno5_2D.jl:75-99), evaluated on the resident
 * state up to the colour-table lookup: rimg (Ny x Nx Int16, column-major -- the reference transposes) receives
 * the 1-based colour index of every pixel, so the caller's image is cmap[rimg].  component: 0..8 = the 9
 * components of q in the reference's order, 9 = bxb, 10 = byb, 11 = compute_divB.  lo == hi == 0 selects the
 * automatic range (finite pixels); log_scale = 1 is the reference's log = true.  range_used may be NULL. */
int weno5_arr2img(weno5_ctx *ctx, int component, double lo, double hi, int log_scale, int ncolours, short *rimg,
                  double range_used[2]);

/* ---- 3-D extension (SURVEY "next" row f4).  The reference has no 3-D solver; the docstring of classical_RK4_step
 * (Weno5_2D.jl:113-123) states the design: the scheme is dimensionally unsplit and "the Roe solver / Weno functions are
 * inherently the same for 1D, 2D and 3D", every direction being swept by the same functions on a rotated state.  These
 * entry points are classical_RK4_step / timestep / boundaries! / compute_divB continued that way: a third sweep in the
 * frame (z,x,y), flux differences of rho, M, E from three sweeps, constrained transport of three face fields from
 * edge-centred EMFs (corner_fluxes, :319, per coordinate plane).  A state that is uniform along one axis reproduces the
 * 2-D step.  Host arrays continue the 2-D layout:
 *      q   (Nx,Ny,Nz,9)  [rho,Mx,My,Mz,Bx,By,Bz,S,E]     bxb, byb, bzb (Nx,Ny,Nz): B on the x-, y-, z-face i+1/2, ...
 * with N* = n* + 2*3, column-major.  Single GPU. */
typedef struct {
    int nx, ny, nz;          /* interior cells                            */
    double dx, dy, dz;
    double gamma, cfl, eps;  /* as in weno5_params                        */
    int bc_x, bc_y, bc_z;    /* WENO5_BC_*                                */
    int es_roundtrip;        /* as in weno5_params                        */
} weno5_params3d;

typedef struct weno5_ctx3d weno5_ctx3d;
int weno5_create_3d(const weno5_params3d *params, int device, weno5_ctx3d **out);
int weno5_destroy_3d(weno5_ctx3d *ctx);
/* bxb == byb == bzb == NULL derives the face fields with interpolateB_center2face (:1669); applies boundaries! (:39) */
int weno5_upload_3d(weno5_ctx3d *ctx, const double *q, const double *bxb, const double *byb, const double *bzb);
int weno5_download_3d(weno5_ctx3d *ctx, double *q, double *bxb, double *byb, double *bzb);
/* timestep(q) (:1930) with three terms: dt = courFac / (vxmax/dx + vymax/dy + vzmax/dz) */
int weno5_timestep_3d(weno5_ctx3d *ctx, double *dt, double vmax[3]);
/* classical_RK4_step (:125) in place */
int weno5_rk4_step_3d(weno5_ctx3d *ctx, double dt);
/* the driver loop (:50-72) on the device, as weno5_advance */
int weno5_advance_3d(weno5_ctx3d *ctx, int nsteps, double tmax, double *t, int *nsteps_done);
/* compute_divB (:1986) with the z term -> (Nx,Ny,Nz) */
int weno5_divb_3d(weno5_ctx3d *ctx, double *divb);
/* interior sums of the 9 components of q and of bxb, byb, bzb */
int weno5_totals_3d(weno5_ctx3d *ctx, double sums[12]);
/* one-shot: host arrays in -> host arrays out */
int weno5_classical_rk4_step_3d(const weno5_params3d *params, const double *q0, const double *bxb0, const double *byb0,
                                const double *bzb0, double dt, double *q4, double *bxb4, double *byb4, double *bzb4);
unsigned long long weno5_launch_count_3d(const weno5_ctx3d *ctx);
/* accumulated device time (ms, CUDA events) of the x, y and z sweep kernels and their total launch count since the
 * last read */
int weno5_profile_3d(weno5_ctx3d *ctx, int enable);
int weno5_profile_read_3d(weno5_ctx3d *ctx, double sweep_ms[3], unsigned long long *sweep_launches);

/* ---- one-shot, reference-shaped calls: host arrays in -> host arrays out */
/* (q4,bxb4,byb4) = classical_RK4_step(q0,bxb0,byb0,dt)   Weno5_2D.jl:125-317 */
int weno5_classical_rk4_step_2d(const weno5_params *params, const double *q0, const double *bxb0,
                                const double *byb0, double dt, double *q4, double *bxb4, double *byb4);
/* q4 = classical_RK4_step(q,dt)[1]                        Weno5_1D.jl:152-236 */
int weno5_classical_rk4_step_1d(const weno5_params *params, const double *q0, double dt, double *q4);
/* same through an existing handle (no allocation): upload, step, download */
int weno5_step_host(weno5_ctx *ctx, const double *q0, const double *bxb0, const double *byb0,
                    double dt, double *q4, double *bxb4, double *byb4);

/* ---- pipelined host -> host step (2-D): the same classical_RK4_step(q0,bxb0,byb0,dt) -> (q4,bxb4,byb4)
 * (Weno5_2D.jl:125-317), streamed through the GPU in y-slabs of `rows_per_slab` rows that carry the
 * domain of dependence of one step as overlap, so that the host->device copies, the steps of the slabs and
 * the device->host copies overlap.  Every row crosses PCIe once in each direction: the inputs are uploaded row range
 * by row range into an image of the grid on the device from which the slabs gather their windows, the result rows
 * leave through a second image (device memory: 21 planes of the grid + three slab solvers; rows_per_slab = 128 is
 * the measured optimum at 4096^2).  Results are bit-identical to weno5_step_host for every slab size.  Output arrays must not alias the inputs; page-locked host
 * arrays (weno5_host_alloc) are needed for the copies to overlap.  *dt_next receives
 * timestep(q4) (Weno5_2D.jl:1930), computed on the fly from the result. */
typedef struct weno5_pipe weno5_pipe;
int weno5_pipe_create(const weno5_params *params, int device, int rows_per_slab, weno5_pipe **out);
int weno5_pipe_step(weno5_pipe *pipe, const double *q0, const double *bxb0, const double *byb0, double dt,
                    double *q4, double *bxb4, double *byb4, double *dt_next);
/* 1 if the last weno5_pipe_step was given page-locked host arrays (weno5_host_alloc, or memory registered with
 * cudaHostRegister): only then do its H2D / D2H copies overlap the compute.  Pageable arrays (a plain Julia
 * Array, numpy.zeros) give correct results through driver-staged, serialised copies -- slower than weno5_step_host. */
int weno5_pipe_overlapped(const weno5_pipe *pipe);
int weno5_pipe_destroy(weno5_pipe *pipe);

/* ---- host memory helpers (page-locked buffers make upload/download asynchronous-capable) */
int weno5_host_alloc(void **ptr, unsigned long long bytes);
int weno5_host_free(void *ptr);

/* ---- instrumentation */
/* kernels launched by this handle since creation */
unsigned long long weno5_launch_count(const weno5_ctx *ctx);
/* accumulated device time (ms, CUDA events on the handle's stream) of the flux kernels and
 * their launch count since the last reset; enables the per-kernel roofline in bench.py */
int weno5_profile(weno5_ctx *ctx, int enable);
int weno5_profile_read(weno5_ctx *ctx, double *flux_ms, unsigned long long *flux_launches, int reset);
/* debug: copy an internal device field (by name: "rhs0".."rhs6","fsy","gsx",...) in the public
 * (Nx,Ny_local+6) layout to the host */
int weno5_debug_fetch(weno5_ctx *ctx, const char *name, double *out);
/* raw CUDA stream of the handle (cudaStream_t as void*) for event timing by the caller */
void *weno5_stream(weno5_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* WENO5MHD_H */
