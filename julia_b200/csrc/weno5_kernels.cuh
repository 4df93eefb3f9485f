// Kernel-side data structures and launcher declarations of libweno5mhd (sm_100a).
//
// Device layout (DESIGN.md "data layout in HBM"): every field is its own plane of
// pitch x NYI doubles, i fastest, with G = 4 ghost layers per side (one more than the
// reference's NBnd = 3, so that every face the CT update needs has a complete stencil
// and y-slabs need a single exchange per stage).  With G = 4 the internal 0-based index
// equals the reference's 1-based public index: internal (i,j) == public (i,j).
#pragma once
#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched through the runtime)
#include <cuda_runtime.h>

namespace w5 {

constexpr int G = 4;

struct Geom {
    int NXI, NYI;        // cells incl. 4 ghosts per side (NYI = 1 in 1-D)
    int pitch;           // doubles per row (multiple of 16 -> 128-byte aligned rows)
    int Nx, Ny;          // public array sizes (nx+6, ny_local+6; Ny = 1 in 1-D)
    int is1d;
    size_t plane;        // pitch * NYI
};

// physical components of the cell state planes
enum { C_RHO = 0, C_MX, C_MY, C_MZ, C_BX, C_BY, C_BZ, C_E, C_S, NCELL = 9 };
// the six cell fields advanced by flux differences in 2-D (Bx,By come from the CT update)
// order: rho, Mx, My, Mz, Bz, E  (+ By as 7th entry of rhs in 1-D)

enum SideBC { SIDE_OUTFLOW = 0, SIDE_PERIODIC = 1, SIDE_NONE = 2 /* slab-internal: filled by exchange */ };

struct FluxArgs {
    const double *qc[8];     // current state: rho,Mx,My,Mz,Bx,By,Bz,E
    double *rhs[7];          // x sweep: out (d fhat);  y sweep: in
    double *bs;              // fsy (x sweep) / gsx (y sweep)
    const double *base[6];   // y sweep: RK base state (stages 1-3)
    double *acc[6];          // y sweep: running RK combination -q0 + q1 + 2 q2
    double *qn[6];           // y sweep: new state
    const double *dtp;       // device time step
    double c_dx, c_dy;       // c_k/dx, c_k/dy
    int stage;               // 1..4
    int N, NP, pitch;        // cells along the sweep, lines across it
    int seg_len, nseg;       // faces per segment, segments per line
    double gam, eps;
    double g2;               // (gam-2)/(gam-1), a constant of the run (face_coefficients)
    int is1d;
    int branch;              // 0: E branch (qc[7] = E), 1: S branch (qc[7] = S; weno5_flux_difference_S, Weno5_2D.jl:361)
    // x sweep on a y-slab whose ghost rows are still in flight: the strips with ghost rows wait for flags[0] (rows
    // from the lower neighbour) / flags[1] (upper) >= wait_seq inside the kernel; nullptr: nothing to wait for
    const unsigned long long *wait_flags;
    unsigned long long wait_seq;
    int wait_lo, wait_hi;
    double *wait_err;        // control block: [6] = 2 if a neighbour never arrived
    // TMA staging: plane indices into the pool (z coordinate of the 3-D tensor maps) and the maps
    // themselves (host pointers; copied into the kernel's parameter space at launch)
    int qc_plane[8], rhs_plane[7], base_plane[6], acc_plane[6];
    const CUtensorMap *tmap_x, *tmap_y;      // each points at 4 maps: boxes of 8, 6, 4, 2 planes
};

struct CtArgs {
    const double *fsy, *gsx;
    const double *bx0, *by0;     // base face fields (stages 1-3)
    const double *bxc, *byc;     // current face fields (for the RK combination)
    double *accbx, *accby;
    double *bxn, *byn;
    const double *dtp;
    double c_dx, c_dy;
    int stage;
    int Nx, Ny, pitch;
    int oxlo, oxhi, oylo, oyhi;  // 1 where the side is a physical outflow boundary
    int bx_i0, bx_i1, bx_j0, bx_j1;   // write ranges (inclusive) of bxn
    int by_i0, by_i1, by_j0, by_j1;   // and of byn
    int no_acc;                  // 1: do not touch accbx/accby (candidate CT passes of the dual-energy stage)
};

// dual-energy stage (Weno5_2D.jl:134-180 with a live ES_Switch, :1734-1768)
struct SwitchArgs {
    const double *ce[8];         // E candidate: rho,Mx,My,Mz,Bz,E,Bx,By
    const double *cs[8];         // S candidate: rho,Mx,My,Mz,Bz,S,Bx,By
    double *qn[7];               // selected state: rho,Mx,My,Mz,Bz,S,E
    double *flag;                // 1.0 where |S(E) - S| >= threshold
    int NXI, NYI, pitch;
    double gam, threshold;
    const double *dtp;
    double *qn_by;               // 1-D only: selected By (flux-advanced there); nullptr in 2-D
    int flags_only;              // 1: keep the energy state everywhere, only record the flags (Weno5_1D.jl:252-254)
};

void launch_flux_x(const FluxArgs &a, cudaStream_t s);
void launch_flux_y(const FluxArgs &a, cudaStream_t s);
void launch_update_1d(const FluxArgs &a, double *const qn7[7], const double *const base7[7],
                      double *const acc7[7], cudaStream_t s);
void launch_ct_faces(const CtArgs &a, cudaStream_t s);
void launch_ct_centers(const double *bxn, const double *byn, double *Bx, double *By,
                       const Geom &g, const double *dtp, cudaStream_t s);
void launch_face_bc(double *bx, double *by, const Geom &g, const int side[4], const double *dtp, cudaStream_t s);
void launch_bc_fill(double *const planes[], int nplanes, const Geom &g, const int side[4],
                    const double *dtp, cudaStream_t s);
void launch_halo_pack(double *const planes[], int n, const Geom &g, double *up, double *down, int has_up, int has_down,
                      int unpack, cudaStream_t s);
// pipelined host step: `rows` rows of n unpadded source planes (pitch spitch) -> padded slab planes (pitch dpitch)
void launch_gather_rows(const double *const src[], double *const dst[], int n, int Nx, int spitch, int dpitch, int rows,
                        cudaStream_t s);
// halo push over peer memory: edge rows of n planes -> the neighbours' ghost rows, then seq -> their flag words
struct PushArgs {
    const double *src[NCELL];    // this slab's planes
    int plane_idx[NCELL];        // their index in the plane pool (the same on every rank)
    int n, NYI, pitch;
    double *peer_up, *peer_down; // neighbours' pools (nullptr: no neighbour on that side)
    size_t plane_up, plane_down; // their plane sizes (slabs may differ by one row)
    int NYI_down;                // rows of the lower neighbour's planes
    unsigned long long *flag_up, *flag_down;   // upper neighbour's "from below" word, lower neighbour's "from above" word
    unsigned int *counter;       // blocks done (local)
    unsigned long long seq;
};
void launch_halo_push(const PushArgs &a, cudaStream_t s);
void launch_peer_count(unsigned long long *counter, long long delta, cudaStream_t s);
void launch_halo_wait(const unsigned long long *flags, int need_down, int need_up, unsigned long long seq, double *ctl,
                      cudaStream_t s);
void launch_es_switch(const SwitchArgs &a, cudaStream_t s);
// fsy/gsx of the S sweep <- E sweep on the faces of switched cells and of their low-side neighbours (:1754-1758)
void launch_flux_select(const double *flag, const double *fsyE, const double *gsxE, double *fsyS, double *gsxS,
                        const Geom &g, const double *dtp, cudaStream_t s);
void launch_ct_centers_entropy(const double *bxn, const double *byn, double *const q9[9], const Geom &g, double gam,
                               int roundtrip, const double *dtp, cudaStream_t s);
// stage 4: the same pass + wave-speed maxima of the new state (atomicMax into vmax_out[0..1])
void launch_ct_centers_entropy_vmax(const double *bxn, const double *byn, double *const q9[9], const Geom &g, double gam,
                                    int roundtrip, const double *dtp, int xhi_side, unsigned long long *vmax_out,
                                    cudaStream_t s);
void launch_entropy(double *const q9[9], const Geom &g, double gam, int roundtrip, int all_cells,
                    const double *dtp, cudaStream_t s);
void launch_vmax(const double *const q9[9], const Geom &g, double gam, unsigned long long *vmax_bits,
                 cudaStream_t s);
void launch_vmax_rows(const double *const q9[9], const Geom &g, double gam, unsigned long long *vmax_bits, int j0, int j1,
                      int zero_first, cudaStream_t s);
// ctl[0]=dt ctl[1]=t ctl[2]=vxmax ctl[3]=vymax ctl[4]=tmax ctl[5]=steps_done ctl[6]=bad-flag
// max-all-reduce of the two wave speeds over peer memory inside dt_kernel: slots[r] = rank r's slot array
// ([2 parities][32 ranks][4 words]: vx bits, vy bits, sequence number), own rank included
struct ShareArgs { int n, rank; unsigned long long seq; unsigned long long *slots[8]; };
constexpr int SHARE_MAX_RANKS = 8;
constexpr int SHARE_SLOT_WORDS = 2 * 32 * 4;
// clear = 1: zero the two maxima after reading them (accumulator of the fused stage-4 reduction)
void launch_dt_from_vmax(unsigned long long *vmax_bits, double *ctl, double dx, double dy,
                         double cfl, int is1d, int clear, const ShareArgs *share, cudaStream_t s);
void launch_advance_time(double *ctl, cudaStream_t s);
void launch_center2face(const double *Bx, const double *By, double *bxb, double *byb, const Geom &g,
                        cudaStream_t s);
void launch_divb(const double *bxb, const double *byb, double *divb, const Geom &g, double dx, double dy,
                 cudaStream_t s);
void launch_totals(const double *const planes[], int nplanes, const Geom &g, double *sums, cudaStream_t s);
// Arr2Img (JColorMaps.jl:92-146) up to the colour-table lookup; keys: 3 x u64 scratch, rng: 2 doubles (range used)
void launch_arr2img(const double *field, const Geom &g, double lo, double hi, int log_scale, int ncolours,
                    unsigned long long *keys, short *out, double *rng, cudaStream_t s);
// 1-D solver as one cluster-resident kernel (weno5_line1d.cu): cells per CTA, or 0 if the line does not fit;
// nsteps CFL steps of the driver loop (fixed_dt <= 0) or one classical_RK4_step with the given dt
int line1d_chunk(int N);
cudaError_t launch_line1d(double *const q9[NCELL], double *ctl, const Geom &g, const int side[4], int nsteps, double fixed_dt,
                          double dx, double gam, double eps, double cfl, cudaStream_t s);
int flux_chunk(int dir);          // faces per march step of the active launch variant
int flux_lines_per_cta(int dir);  // lines across the sweep per CTA
int flux_uses_tma();              // 1 if the active variant stages through TMA tensor maps
int flux_waits_in_kernel();       // 1 if the x sweep of the active variant can wait for the halo flags itself
int flux_resident_ctas();         // CTAs of the flux kernel resident on the GPU at a time (wave planning)
void flux_box(int dir, int box[2]);   // TMA box (cells along x, rows along y) of one march step

}  // namespace w5
