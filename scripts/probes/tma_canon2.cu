// steps from the canonical libcu++ example towards the probe: u8 map, box {256,1}; then raw PTX with per-warp barriers
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
__global__ void k_a(const __grid_constant__ CUtensorMap tm, int x, int y, uint8_t *out)
{
    __shared__ alignas(128) uint8_t buf[256];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(buf, &tm, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(buf));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    out[threadIdx.x] = buf[threadIdx.x];
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k_b(const __grid_constant__ CUtensorMap tm, int x, int y, uint8_t *out, int variant)
{
    __shared__ __align__(128) uint8_t buf[4][256];
    __shared__ __align__(8) unsigned long long bar[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t mb = smem_u32(&bar[warp]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
        if (variant & 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        else asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
        if (variant & 2) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(256u) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(buf[warp])), "l"(&tm), "r"(x + warp), "r"(y), "r"(mb) : "memory");
        if (!(variant & 2)) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(256u) : "memory");
    }
    __syncwarp();
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mb), "r"(0u) : "memory");
    out[threadIdx.x] = buf[warp][lane];
}
int main(int argc, char **argv)
{
    const int which = argc > 1 ? atoi(argv[1]) : 0, variant = argc > 2 ? atoi(argv[2]) : 0;
    cudaFree(0);
    const size_t bytes = 64 << 20;
    uint8_t *d, *out;
    cudaMalloc(&d, bytes); cudaMalloc(&out, 256);
    uint8_t *h = new uint8_t[bytes];
    for (size_t i = 0; i < bytes; ++i) h[i] = (uint8_t)(i * 7 + (i >> 8));
    cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t size[2] = {1 << 20, bytes >> 20}; cuuint64_t stride[1] = {1 << 20}; cuuint32_t box[2] = {256, 1}; cuuint32_t es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    const int x = argc > 3 ? atoi(argv[3]) : 1003, y = 5;
    if (which == 0) k_a<<<1, 128>>>(tm, x, y, out); else k_b<<<1, 128>>>(tm, x, y, out, variant);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel %d variant %d -> %s\n", which, variant, cudaGetErrorString(e));
    uint8_t ho[128];
    cudaMemcpy(ho, out, sizeof ho, cudaMemcpyDeviceToHost);
    const size_t o = ((size_t)y << 20) + x;
    printf("out = %d %d %d %d, expect %d %d %d %d\n", ho[0], ho[1], ho[2], ho[3], h[o], h[o + 1], h[o + 2], h[o + 3]);
    return 0;
}
