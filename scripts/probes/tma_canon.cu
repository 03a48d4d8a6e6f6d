// canonical TMA example of the CUDA programming guide (libcu++ wrappers), int 2-D tile 64x64: does tensor TMA run on this box at all?
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <stdio.h>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int GW = 1024, GH = 1024, SW = 64, SH = 64;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int *out)
{
    __shared__ alignas(128) int smem_buffer[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    out[threadIdx.x] = smem_buffer[0][threadIdx.x % SW];
}
int main()
{
    cudaFree(0);
    int *d, *out;
    cudaMalloc(&d, GW * GH * 4); cudaMalloc(&out, 128 * 4);
    int *h = new int[GW * GH];
    for (int i = 0; i < GW * GH; ++i) h[i] = i;
    cudaMemcpy(d, h, GW * GH * 4, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t size[2] = {GW, GH}; cuuint64_t stride[1] = {GW * 4}; cuuint32_t box[2] = {SW, SH}; cuuint32_t es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    kernel<<<1, 128>>>(tm, 64, 128, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel -> %s\n", cudaGetErrorString(e));
    int ho[128];
    cudaMemcpy(ho, out, sizeof ho, cudaMemcpyDeviceToHost);
    printf("out[0..3] = %d %d %d %d (expect %d..)\n", ho[0], ho[1], ho[2], ho[3], 128 * GW + 64);
    return 0;
}
