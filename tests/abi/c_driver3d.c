/* The 3-D extension driven from plain C (the ccall sequence of Weno5GPU.WENO5_3D, INTEGRATION.md "3-D extension"),
 * linked against libweno5mhd.so only:
 *     create_3d -> upload_3d(q, bxb, byb, bzb) -> advance_3d(nsteps) -> divb_3d -> download_3d
 * Input  file: int32 nx, ny, nz, nsteps, bc; double xsize, ysize, zsize; then q[N*9], bxb[N], byb[N], bzb[N]
 * Output file: t, steps done (as double); then q, bxb, byb, bzb, divb of the final state (Float64, column-major).
 * tests/test_gpu_3d.py feeds it a smooth 3-D state and compares with the 3-D oracle's driver loop. */
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
    if (argc != 3) { fprintf(stderr, "usage: c_driver3d in.bin out.bin\n"); return 1; }
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) return 1;
    int hdr[5];
    double size[3];
    if (fread(hdr, sizeof(int), 5, fi) != 5 || fread(size, sizeof(double), 3, fi) != 3) return 1;
    const int nx = hdr[0], ny = hdr[1], nz = hdr[2], nsteps = hdr[3], bc = hdr[4];
    const size_t n = ((size_t)nx + 2 * WENO5_NBND) * ((size_t)ny + 2 * WENO5_NBND) * ((size_t)nz + 2 * WENO5_NBND);
    double *q = malloc(9 * n * sizeof(double)), *b[3], *divb = malloc(n * sizeof(double));
    for (int m = 0; m < 3; ++m) b[m] = malloc(n * sizeof(double));
    if (!q || !b[0] || !b[1] || !b[2] || !divb) return 1;
    if (fread(q, sizeof(double), 9 * n, fi) != 9 * n) return 1;
    for (int m = 0; m < 3; ++m)
        if (fread(b[m], sizeof(double), n, fi) != n) return 1;
    fclose(fi);

    weno5_params3d P;
    P.nx = nx; P.ny = ny; P.nz = nz;
    P.dx = size[0] / nx; P.dy = size[1] / ny; P.dz = size[2] / nz;
    P.gamma = 5.0 / 3.0; P.cfl = 0.8; P.eps = 1e-6;                 /* Weno5_2D.jl:7-29 */
    P.bc_x = P.bc_y = P.bc_z = bc; P.es_roundtrip = 1;

    weno5_ctx3d *ctx = NULL;
    double t = 0.0, head[2];
    int done = 0;
    CHECK(weno5_create_3d(&P, 0, &ctx));
    CHECK(weno5_upload_3d(ctx, q, b[0], b[1], b[2]));
    CHECK(weno5_advance_3d(ctx, nsteps, 0.0, &t, &done));            /* {timestep; classical_RK4_step} x nsteps */
    CHECK(weno5_divb_3d(ctx, divb));
    CHECK(weno5_download_3d(ctx, q, b[0], b[1], b[2]));
    CHECK(weno5_destroy_3d(ctx));

    FILE *fo = fopen(argv[2], "wb");
    if (!fo) return 1;
    head[0] = t; head[1] = (double)done;
    fwrite(head, sizeof(double), 2, fo);
    fwrite(q, sizeof(double), 9 * n, fo);
    for (int m = 0; m < 3; ++m) fwrite(b[m], sizeof(double), n, fo);
    fwrite(divb, sizeof(double), n, fo);
    fclose(fo);
    free(q); free(divb);
    for (int m = 0; m < 3; ++m) free(b[m]);
    return 0;
}
