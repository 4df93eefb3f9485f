// accuracy of the rcp/rsqrt hardware seeds and of the Newton refinements used in weno5_math.cuh
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(double *err, int n)
{
    double e_seed = 0, e1 = 0, e2 = 0, e3 = 0, s_seed = 0, s1 = 0, s2 = 0, h1 = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double x = exp2((i % 97) - 48.0) * (1.0 + (double)((i * 2654435761u) & 0xffffff) / 16777216.0);
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        double ex = 1.0 / x;
        e_seed = fmax(e_seed, fabs(r - ex) / ex);
        double e = fma(-x, r, 1.0);
        double hh = fma(r, fma(e, e, e), r);            // one Halley-type step
        h1 = fmax(h1, fabs(hh - ex) / ex);
        r = fma(r, e, r); e1 = fmax(e1, fabs(r - ex) / ex);
        e = fma(-x, r, 1.0); r = fma(r, e, r); e2 = fmax(e2, fabs(r - ex) / ex);
        e = fma(-x, r, 1.0); r = fma(r, e, r); e3 = fmax(e3, fabs(r - ex) / ex);
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        double ey = 1.0 / sqrt(x);
        s_seed = fmax(s_seed, fabs(y - ey) / ey);
        double hx = 0.5 * x;
        double d = fma(-(hx * y), y, 0.5); y = fma(y, d, y); s1 = fmax(s1, fabs(y - ey) / ey);
        d = fma(-(hx * y), y, 0.5); y = fma(y, d, y); s2 = fmax(s2, fabs(y - ey) / ey);
    }
    double v[8] = {e_seed, e1, e2, e3, s_seed, s1, s2, h1};
    for (int j = 0; j < 8; ++j) atomicMax((unsigned long long *)&err[j], (unsigned long long)__double_as_longlong(v[j]));
}
int main()
{
    double *d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
    k<<<296, 256>>>(d, 1 << 24);
    double h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("rcp seed %.3e  +1 Newton %.3e  +2 %.3e  +3 %.3e | Halley x1 %.3e\nrsqrt seed %.3e  +1 Newton %.3e  +2 %.3e\n", h[0], h[1], h[2], h[3], h[7], h[4], h[5], h[6]);
}
