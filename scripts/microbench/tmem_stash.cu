// TMEM as a per-thread stash for a non-tensor kernel (B200, sm_100a): every thread parks 56 doubles
// (tcgen05.st.32x32b.x2, one double per instruction, strided "transposed" layout [field][slot]) and takes them
// back field by field (tcgen05.ld.32x32b.x16 = 8 doubles).  Checks the values and measures the cost per
// round trip, alone and next to a stream of dependent-free DFMAs -- 3 CTAs of 128 threads per SM, 128 columns each.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_stash tmem_stash.cu && ./tmem_stash
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_st2(unsigned taddr, double v)
{
    const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ void tm_ld16(unsigned taddr, double v[8])
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

template <int MODE>   // 0: check, 1: stash round trips only, 2: round trips + 1000 DFMAs, 3: DFMAs only
__global__ void __launch_bounds__(128, 3) stash_kernel(double *out, int iters, int *bad)
{
    __shared__ unsigned tbase;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&tbase)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned base = tbase + ((unsigned)(warp * 32) << 16);      // this warp's 32 lanes
    double acc[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    const double x = 1.0 + 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        if (MODE != 3) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int m = 0; m < 7; ++m) tm_st2(base + m * 16 + k * 2, 1000.0 * threadIdx.x + 10.0 * m + k + it + (MODE == 0 ? blockIdx.x : 0));
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        if (MODE >= 2) {
#pragma unroll 1
            for (int r = 0; r < 125; ++r)
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fma(acc[k], x, 1e-3);
        }
        if (MODE != 3) {
#pragma unroll
            for (int m = 0; m < 7; ++m) {
                double v[8];
                tm_ld16(base + m * 16, v);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (MODE == 0 && v[k] != 1000.0 * threadIdx.x + 10.0 * m + k + it + blockIdx.x) atomicAdd(bad, 1);
                    acc[k] += v[k];
                }
            }
        }
    }
    double s = 0;
    for (int k = 0; k < 8; ++k) s += acc[k];
    out[blockIdx.x * 128 + threadIdx.x] = s;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(128));
}

template <int MODE> float run(double *out, int *bad, int iters)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    stash_kernel<MODE><<<148 * 3, 128>>>(out, 2, bad);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    stash_kernel<MODE><<<148 * 3, 128>>>(out, iters, bad);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", MODE, cudaGetErrorString(e)); exit(1); }
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main()
{
    double *out; int *bad, hbad = 0;
    cudaMalloc(&out, 148 * 3 * 128 * sizeof(double));
    cudaMalloc(&bad, sizeof(int));
    cudaMemset(bad, 0, sizeof(int));
    run<0>(out, bad, 50);
    cudaMemcpy(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost);
    printf("check: %d mismatches (3 CTAs x 128 threads per SM, 56 doubles per thread)\n", hbad);
    const int iters = 2000;
    const float t1 = run<1>(out, bad, iters), t2 = run<2>(out, bad, iters), t3 = run<3>(out, bad, iters);
    const double clk = 1.965e9;
    printf("round trip alone        : %8.3f ms  = %7.1f cycles per round trip (56 st.x2 + 7 ld.x16 per thread, 12 warps/SM)\n", t1, t1 * 1e-3 * clk / iters);
    printf("1000 DFMA alone         : %8.3f ms  = %7.1f cycles per iteration\n", t3, t3 * 1e-3 * clk / iters);
    printf("round trip + 1000 DFMA  : %8.3f ms  = %7.1f cycles per iteration (overlap: %.0f %% of the round trip hidden)\n", t2,
           t2 * 1e-3 * clk / iters, 100.0 * (1.0 - (t2 - t3) / t1));
    return hbad != 0;
}
