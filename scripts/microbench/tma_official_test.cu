// CUDA programming guide pattern (cuda::barrier + cde::cp_async_bulk_tensor_2d_global_to_shared), 2-D, int and double
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int BW = 16, BH = 8;
__global__ void k(const __grid_constant__ CUtensorMap tm, double *out, int x, int y)
{
    __shared__ alignas(128) double tile[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&tile, &tm, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(tile));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i / BW][i % BW];
}
int main()
{
    const int pitch = 80, ny = 72;
    double *d, *o;
    cudaMalloc(&d, sizeof(double) * pitch * ny);
    cudaMalloc(&o, sizeof(double) * 1024);
    double *h = new double[pitch * ny];
    for (int i = 0; i < pitch * ny; ++i) h[i] = i;
    cudaMemcpy(d, h, sizeof(double) * pitch * ny, cudaMemcpyHostToDevice);
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    CUtensorMap tm;
    cuuint64_t dims[2] = {pitch, ny};
    cuuint64_t strides[1] = {pitch * 8ull};
    cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d (query %d)\n", (int)r, (int)q);
    k<<<1, 128>>>(tm, o, 16, 3);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    double ho[128];
    cudaMemcpy(ho, o, sizeof(ho), cudaMemcpyDeviceToHost);
    printf("got %g %g ... %g ; expect %d %d ... %d\n", ho[0], ho[1], ho[16], 3 * pitch + 16, 3 * pitch + 17, 4 * pitch + 16);
    return 0;
}
