// Does the FP64 pipe of B200 sustain one warp-wide DFMA every 2 cycles per sub-partition when all three source
// operands are DISTINCT registers (no constant / reused operand)?  fp64_peak.cu measures x = fma(x, a, b) with a, b
// fixed (ptxas marks them .reuse): 1 fresh 64-bit operand per instruction.  The flux kernel's DFMAs mostly read 2-3
// fresh operands.  Variants: fresh operands per instruction = 1, 2, 3 (and DADD with 2), ILP 16, 8 or 12 warps per SM.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fp64_operands fp64_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ILP = 16;

template <int MODE>
__global__ void k(double *out, const double *in, int iters)
{
    double x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        x[i] = in[i] + threadIdx.x * 1e-9;
        y[i] = in[ILP + i] + threadIdx.x * 1e-10;      // thread-dependent: stays in vector registers
        z[i] = in[2 * ILP + i] - threadIdx.x * 1e-10;
    }
    const double a = in[100] + threadIdx.x * 1e-12, b = in[101] + threadIdx.x * 1e-15;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 1) x[i] = fma(x[i], a, b);
            if (MODE == 2) x[i] = fma(x[i], y[i], b);
            if (MODE == 3) x[i] = fma(y[i], z[(i + 1) % ILP], x[i]);
            if (MODE == 4) x[i] = x[i] + y[i];
            if (MODE == 5) x[i] = fma(y[i], z[(i + 1) % ILP], x[(i + 3) % ILP]);   // 3 fresh, result to a 4th register
            // pairs that share the multiplier in the same operand position (the lock-step pair projection):
            // first of the pair 3 fresh operands, second 2 fresh + 1 from the operand-reuse cache
            if (MODE == 6 && (i & 1) == 0) { x[i] = fma(y[i], z[i], x[i]); x[i + 1] = fma(y[i], z[i + 1], x[i + 1]); }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> double run(int warps_per_sm, double *out, const double *in, int sms)
{
    const int iters = 4000, threads = 128, bps = warps_per_sm / 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * bps, threads>>>(out, in, 10);
    cudaEventRecord(e0);
    k<MODE><<<sms * bps, threads>>>(out, in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // warp instructions per cycle per sub-partition at 1.965 GHz (peak 0.5)
    const double insts = (double)ILP * iters * warps_per_sm / 4.0;
    return insts / (ms * 1e-3 * 1.965e9);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *out, *in, h[128];
    for (int i = 0; i < 128; ++i) h[i] = 1.0 + 1e-7 * i;
    h[100] = 0.9999999; h[101] = 1e-9;
    cudaMalloc(&out, sizeof(double) * sms * 16 * 128);
    cudaMalloc(&in, sizeof(h));
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("FP64 warp instructions per cycle per sub-partition (peak 0.5), ILP %d\n", ILP);
    printf("warps/SM  fma(x,a,b) 1 fresh   fma(x,y,b) 2 fresh   fma(y,z,x) 3 fresh   x+y 2 fresh   fma(y,z,x')->x 3 fresh   pairs sharing the multiplier\n");
    for (int w : {4, 8, 12, 16}) {
        printf("%5d     %8.3f            %8.3f             %8.3f            %8.3f       %8.3f                 %8.3f\n", w, run<1>(w, out, in, sms),
               run<2>(w, out, in, sms), run<3>(w, out, in, sms), run<4>(w, out, in, sms), run<5>(w, out, in, sms), run<6>(w, out, in, sms));
    }
    return 0;
}
