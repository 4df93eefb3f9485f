// Accuracy (in ulp, against the correctly rounded device functions) of the branch-free fast_rcp / fast_rsqrt /
// fast_sqrt of julia_b200/csrc/weno5_math.cuh.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../julia_b200/csrc
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "weno5_math.cuh"

__device__ double ulps(double got, double ref) { return fabs(got - ref) / (fabs(ref) * 2.220446049250313e-16); }

__global__ void k(double *err, int n)
{
    double e[3] = {0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double x = exp2((double)(i % 257) - 128.0) * (1.0 + (double)((i * 2654435761u) & 0xffffff) / 16777216.0);
        e[0] = fmax(e[0], ulps(w5::fast_rcp(x), __drcp_rn(x)));
        e[1] = fmax(e[1], ulps(w5::fast_rsqrt(x), 1.0 / __dsqrt_rn(x)));
        e[2] = fmax(e[2], ulps(w5::fast_sqrt(x), __dsqrt_rn(x)));
    }
    for (int j = 0; j < 3; ++j) atomicMax((unsigned long long *)&err[j], (unsigned long long)__double_as_longlong(e[j]));
}

int main()
{
    double *d, h[3];
    cudaMalloc(&d, 24);
    cudaMemset(d, 0, 24);
    k<<<296, 256>>>(d, 1 << 26);
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("max error in ulp over 6.7e7 arguments in [2^-128, 2^129): fast_rcp %.3f  fast_rsqrt %.3f  fast_sqrt %.3f\n", h[0], h[1], h[2]);
    double z;
    return 0;
}
