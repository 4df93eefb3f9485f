/* Compiled as C11 against include/weno5mhd.h (proves the header is valid C) and prints the layout of
 * weno5_params and weno5_params3d as JSON: tests/test_abi.py compares it with the ctypes mirror (julia_b200/lib.py) and with
 * the table committed next to the Julia struct (julia_b200/julia/abi_layout.json). */
#include <stddef.h>
#include <stdio.h>
#include "weno5mhd.h"

#define FIELD(name) printf("  \"%s\": {\"offset\": %zu, \"size\": %zu},\n", #name, offsetof(weno5_params, name), sizeof(((weno5_params *)0)->name))

int main(void)
{
    printf("{\n \"fields\": {\n");
    FIELD(nx); FIELD(ny); FIELD(dx); FIELD(dy); FIELD(gamma); FIELD(cfl); FIELD(eps); FIELD(bc_x); FIELD(bc_y);
    printf("  \"es_roundtrip\": {\"offset\": %zu, \"size\": %zu}\n },\n", offsetof(weno5_params, es_roundtrip),
           sizeof(((weno5_params *)0)->es_roundtrip));
    printf(" \"sizeof_weno5_params\": %zu,\n \"alignof_weno5_params\": %zu,\n", sizeof(weno5_params), _Alignof(weno5_params));
    printf(" \"WENO5_NBND\": %d,\n \"status\": {\"OK\": %d, \"ERR_ARG\": %d, \"ERR_CUDA\": %d, \"ERR_NCCL\": %d, \"ERR_STATE\": %d, \"ERR_NUMERIC\": %d},\n",
           WENO5_NBND, WENO5_OK, WENO5_ERR_ARG, WENO5_ERR_CUDA, WENO5_ERR_NCCL, WENO5_ERR_STATE, WENO5_ERR_NUMERIC);
    printf(" \"bc\": {\"OUTFLOW\": %d, \"PERIODIC\": %d},\n", WENO5_BC_OUTFLOW, WENO5_BC_PERIODIC);
    /* weno5_params3d (3-D extension) */
#define FIELD3(name) printf("  \"%s\": {\"offset\": %zu, \"size\": %zu},\n", #name, offsetof(weno5_params3d, name), sizeof(((weno5_params3d *)0)->name))
    printf(" \"fields3d\": {\n");
    FIELD3(nx); FIELD3(ny); FIELD3(nz); FIELD3(dx); FIELD3(dy); FIELD3(dz); FIELD3(gamma); FIELD3(cfl); FIELD3(eps);
    FIELD3(bc_x); FIELD3(bc_y); FIELD3(bc_z);
    printf("  \"es_roundtrip\": {\"offset\": %zu, \"size\": %zu}\n },\n", offsetof(weno5_params3d, es_roundtrip),
           sizeof(((weno5_params3d *)0)->es_roundtrip));
    printf(" \"sizeof_weno5_params3d\": %zu,\n \"alignof_weno5_params3d\": %zu\n}\n", sizeof(weno5_params3d), _Alignof(weno5_params3d));
    return 0;
}
