// FP64 FMA pipe microbenchmark for the roofline denominator ("measure a DFMA
// microbenchmark").  Prints throughput for several ILP/occupancy points and the dependent-issue latency.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_tput(double *out, double a, double b, int iters)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_latency(double *out, long long *cyc, double a, double b, int iters)
{
    double x = threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b);
        x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
double run(int blocks_per_sm, int threads, double *d, int sms)
{
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_tput<ILP><<<sms * blocks_per_sm, threads>>>(d, 0.999999, 1e-9, 100);
    cudaEventRecord(e0);
    dfma_tput<ILP><<<sms * blocks_per_sm, threads>>>(d, 0.999999, 1e-9, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * ILP * (double)iters * threads * blocks_per_sm * sms;
    return flops / (ms * 1e-3) / 1e12;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double *d; cudaMalloc(&d, sizeof(double) * sms * 64 * 1024);
    long long *c; cudaMalloc(&c, 8);
    printf("device %s, %d SMs, clock %.0f MHz\n", p.name, sms, p.clockRate / 1e3);
    printf("warps/SM ILP  TFLOP/s\n");
    for (int w : {4, 8, 12, 16, 32, 64}) {
        int threads = 128, bps = w * 32 / threads;
        if (bps < 1) { bps = 1; threads = w * 32; }
        printf("%8d %3d  %7.2f\n", w, 1, run<1>(bps, threads, d, sms));
        printf("%8d %3d  %7.2f\n", w, 2, run<2>(bps, threads, d, sms));
        printf("%8d %3d  %7.2f\n", w, 4, run<4>(bps, threads, d, sms));
        printf("%8d %3d  %7.2f\n", w, 8, run<8>(bps, threads, d, sms));
    }
    dfma_latency<<<1, 32>>>(d, c, 0.999999, 1e-9, 10000);
    long long hc; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency: %.2f cycles\n", (double)hc / (10000.0 * 8));
    // rcp.approx seed accuracy
    return 0;
}
