/* The closest executable stand-in for the Julia ccall sequence of INTEGRATION.md (the reference driver
 * WENO5_2D, Weno5_2D.jl:33-72): plain C, linked against libweno5mhd.so only.
 *     create -> upload(q, bxb, byb) -> { timestep -> classical_RK4_step } x nsteps -> divB -> download
 * Input  file: int32 nx, ny, nsteps; double xsize, ysize; then q[Nx*Ny*9], bxb[Nx*Ny], byb[Nx*Ny]
 * Output file: per step dt, vxmax, vymax; then q, bxb, byb, divb of the final state (all Float64, column-major,
 * the reference's layout).  tests/test_reference_pin.py feeds it the reference's own 16x4 shock-cloud state and
 * compares with the reference's output. */
#include <stdio.h>
#include <stdlib.h>
#include "weno5mhd.h"

#define CHECK(call)                                                                         \
    do {                                                                                    \
        int rc_ = (call);                                                                   \
        if (rc_ != WENO5_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, weno5_last_error()); return 2; } \
    } while (0)

int main(int argc, char **argv)
{
    if (argc != 3) { fprintf(stderr, "usage: c_driver in.bin out.bin\n"); return 1; }
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) return 1;
    int hdr[3];
    double size[2];
    if (fread(hdr, sizeof(int), 3, fi) != 3 || fread(size, sizeof(double), 2, fi) != 2) return 1;
    const int nx = hdr[0], ny = hdr[1], nsteps = hdr[2];
    const size_t Nx = (size_t)nx + 2 * WENO5_NBND, Ny = (size_t)ny + 2 * WENO5_NBND, np = Nx * Ny;
    double *q = malloc(9 * np * sizeof(double)), *bxb = malloc(np * sizeof(double)), *byb = malloc(np * sizeof(double));
    double *divb = malloc(np * sizeof(double));
    if (!q || !bxb || !byb || !divb) return 1;
    if (fread(q, sizeof(double), 9 * np, fi) != 9 * np || fread(bxb, sizeof(double), np, fi) != np ||
        fread(byb, sizeof(double), np, fi) != np) return 1;
    fclose(fi);

    weno5_params P;
    P.nx = nx; P.ny = ny; P.dx = size[0] / nx; P.dy = size[1] / ny;
    P.gamma = 5.0 / 3.0; P.cfl = 0.8; P.eps = 1e-6;                 /* Weno5_2D.jl:7-29 */
    P.bc_x = WENO5_BC_OUTFLOW; P.bc_y = WENO5_BC_OUTFLOW; P.es_roundtrip = 1;

    FILE *fo = fopen(argv[2], "wb");
    if (!fo) return 1;
    weno5_ctx *ctx = NULL;
    CHECK(weno5_create(&P, 0, ny, 0, &ctx));
    CHECK(weno5_upload(ctx, q, bxb, byb));
    for (int s = 0; s < nsteps; ++s) {
        double dt[3];
        CHECK(weno5_timestep(ctx, &dt[0], &dt[1], &dt[2]));          /* timestep(q), :53 */
        CHECK(weno5_rk4_step(ctx, dt[0]));                           /* classical_RK4_step, :58 */
        fwrite(dt, sizeof(double), 3, fo);
    }
    CHECK(weno5_divb(ctx, divb));                                    /* compute_divB, :63 */
    CHECK(weno5_download(ctx, q, bxb, byb));
    CHECK(weno5_destroy(ctx));
    fwrite(q, sizeof(double), 9 * np, fo);
    fwrite(bxb, sizeof(double), np, fo);
    fwrite(byb, sizeof(double), np, fo);
    fwrite(divb, sizeof(double), np, fo);
    fclose(fo);
    free(q); free(bxb); free(byb); free(divb);
    return 0;
}
