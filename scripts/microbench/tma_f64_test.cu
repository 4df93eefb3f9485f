// Stand-alone check of cp.async.bulk.tensor.3d on double planes (development aid).
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tm, double *out, int x, int y, int z, int n)
{
    extern __shared__ __align__(128) double sm[];
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(sm + n);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n * 8) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                             smem_u32(sm)),
                         "l"(&tm), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                         : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                             smem_u32(sm)),
                         "l"(&tm), "r"(x), "r"(y), "r"(smem_u32(bar))
                         : "memory");
    }
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = sm[i];
}

int main(int argc, char **argv)
{
    int dtype = argc > 1 ? atoi(argv[1]) : 8;
    int b0 = argc > 2 ? atoi(argv[2]) : 64, b1 = argc > 3 ? atoi(argv[3]) : 2;
    int rank = argc > 4 ? atoi(argv[4]) : 3;
    int x0 = argc > 5 ? atoi(argv[5]) : 5;
    const int pitch = 80, ny = 72, np = 5;
    double *d, *o;
    cudaMalloc(&d, sizeof(double) * pitch * ny * np);
    cudaMalloc(&o, sizeof(double) * 1024);
    double *h = new double[pitch * ny * np];
    for (int i = 0; i < pitch * ny * np; ++i) h[i] = i;
    cudaMemcpy(d, h, sizeof(double) * pitch * ny * np, cudaMemcpyHostToDevice);
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    CUtensorMap tm;
    cuuint64_t dims[3] = {pitch, ny, np};
    cuuint64_t strides[2] = {pitch * 8ull, (cuuint64_t)pitch * ny * 8ull};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, 1}, es[3] = {1, 1, 1};
    CUresult r = fn(&tm, (CUtensorMapDataType)dtype, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode dtype %d box %d x %d rank %d -> %d\n", dtype, b0, b1, rank, (int)r);
    if (rank == 3) k<3><<<1, 128, 128 * 8 + 64>>>(tm, o, x0, 3, 2, b0 * b1); else k<2><<<1, 128, 128 * 8 + 64>>>(tm, o, x0, 3, 0, b0 * b1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    double ho[128];
    cudaMemcpy(ho, o, sizeof(ho), cudaMemcpyDeviceToHost);
    printf("got %g %g ... %g ; expect %d %d ... %d\n", ho[0], ho[1], ho[b0], 2 * pitch * ny + 3 * pitch + 5, 2 * pitch * ny + 3 * pitch + 6,
           2 * pitch * ny + 4 * pitch + 5);
    return 0;
}
