// FP64 pipe: dependent-issue latency and the instruction-level parallelism one warp / two warps per SM
// sub-partition need to keep it busy (B200).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double *out, long long *cyc, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 1e-3 + k;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP> void run(int warps_per_smsp, double *out, long long *cyc)
{
    const int iters = 2000, threads = 128 * warps_per_smsp;          // 4 sub-partitions per SM
    chain<ILP><<<148, threads>>>(out, cyc, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double per = avg / (iters * 8.0 * ILP);                    // cycles per DFMA per warp
    printf("warps/SMSP %d  ILP %d : %.2f cycles per DFMA per warp, pipe busy %.0f %%\n", warps_per_smsp, ILP, per,
           100.0 * 2.0 * warps_per_smsp / per);
}

int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    for (int w = 1; w <= 3; ++w) {
        run<1>(w, out, cyc); run<2>(w, out, cyc); run<3>(w, out, cyc); run<4>(w, out, cyc); run<6>(w, out, cyc); run<8>(w, out, cyc);
    }
    return 0;
}
