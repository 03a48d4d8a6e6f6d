// TMA probe (not part of the product): can a u8 tensor map with byte-granular coordinates realign arbitrary
// line-sized pieces global -> shared at useful rates on B200, and does a 2-D map with overlapping rows encode?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int WARPS = 7, SEG = 1280, SLOT = 1536;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: five 1-D boxes of 256 B per segment; mode 1: one 2-D box {256, 5}; mode 2: LDG/STS funnel (reference: plain loads)
template <int MODE>
__global__ void __launch_bounds__(WARPS * 32, 4) k_probe(const __grid_constant__ CUtensorMap tm_param, const CUtensorMap *tm_glob, const uint8_t *src, uint8_t *dst, uint64_t nseg, int depth, int stop)
{
    __shared__ __align__(128) uint8_t stage[WARPS][2][SLOT];
    __shared__ __align__(8) unsigned long long bar[WARPS][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t gw = (uint64_t)blockIdx.x * WARPS + warp, nw = (uint64_t)gridDim.x * WARPS;
    const CUtensorMap *tmp = tm_glob ? tm_glob : &tm_param;
    if (lane == 0) {
        for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[warp][b])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (stop == 1) return;
    uint32_t phase[2] = {0, 0};
    auto issue = [&](uint64_t seg, int b) {
        // source: segment seg starts at a byte offset that is not 16-aligned; it lands at stage[b] + 0
        const uint64_t soff = seg * SEG + 3 + (seg * 7) % 13;
        const uint32_t mb = smem_u32(&bar[warp][b]);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((uint32_t)SEG) : "memory");
        __syncwarp();
        if (MODE == 0) {
            if (lane < SEG / 256) {
                const uint64_t o = soff + 256ull * lane;
                const int c0 = (int)(o & 0xfffff), c1 = (int)(o >> 20);       // map C: rows of 1 MiB (+255 overlap)
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(&stage[warp][b][256 * lane])), "l"(tmp), "r"(c0), "r"(c1), "r"(mb) : "memory");
            }
        } else {
            if (lane == 0) {
                const int c0 = (int)(soff & 255), c1 = (int)(soff >> 8);       // map B: rows of 256 B (+255 overlap), box {256, 5}
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(&stage[warp][b][0])), "l"(tmp), "r"(c0), "r"(c1), "r"(mb) : "memory");
            }
        }
    };
    auto wait = [&](int b) {
        const uint32_t mb = smem_u32(&bar[warp][b]);
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mb), "r"(phase[b]) : "memory");
        phase[b] ^= 1;
    };
    uint64_t seg = gw;
    if (depth == 2 && seg < nseg) issue(seg, 0);
    int b = 0;
    for (; seg < nseg; seg += nw) {
        if (depth == 2) { if (seg + nw < nseg) issue(seg + nw, b ^ 1); }
        else issue(seg, b);
        if (stop == 2) { if (depth == 2) b ^= 1; continue; }
        wait(b);
        if (stop == 3) { if (depth == 2) b ^= 1; continue; }
        // smem -> global, 16-byte aligned destination, one bulk store
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + seg * SEG), "r"(smem_u32(&stage[warp][b][0])), "r"((uint32_t)SEG) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
        if (depth == 2) b ^= 1;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// reference: the same segments copied by lanes with two aligned loads + funnel shifts per 16 bytes (what the product does today)
__global__ void __launch_bounds__(WARPS * 32, 4) k_ldg(const uint8_t *src, uint8_t *dst, uint64_t nseg)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t gw = (uint64_t)blockIdx.x * WARPS + warp, nw = (uint64_t)gridDim.x * WARPS;
    for (uint64_t seg = gw; seg < nseg; seg += nw) {
        const uint64_t soff = seg * SEG + 3 + (seg * 7) % 13;
        const uint32_t u = (uint32_t)(soff & 15), bsh = (u & 3) * 8, wsh = u >> 2;
        const uint8_t *sa = src + (soff - u);
        uint4 *d = (uint4 *)(dst + seg * SEG);
        for (int c = lane; c < SEG / 16; c += 32) {
            const uint4 q0 = *(const uint4 *)(sa + 16 * c), q1 = *(const uint4 *)(sa + 16 * c + 16);
            const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            uint4 o;
            o.x = __funnelshift_r(w[wsh], w[wsh + 1], bsh); o.y = __funnelshift_r(w[wsh + 1], w[wsh + 2], bsh);
            o.z = __funnelshift_r(w[wsh + 2], w[wsh + 3], bsh); o.w = __funnelshift_r(w[wsh + 3], w[wsh + 4], bsh);
            __stcs(d + c, o);
        }
    }
}

static int g_dtype = 0, g_promo = 0, g_rank = 2;
static bool encode(CUtensorMap *tm, void *base, uint64_t d0, uint64_t d1, uint64_t stride, uint32_t b0, uint32_t b1, const char *what)
{
    cuuint64_t dims[2] = {d0, d1};
    cuuint64_t strides[1] = {stride};
    cuuint32_t box[2] = {b0, b1};
    cuuint32_t es[2] = {1, 1};
    if (g_dtype) { dims[0] /= 4; box[0] /= 4; }
    CUresult r = cuTensorMapEncodeTiled(tm, g_dtype ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, g_rank, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, g_promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const char *s = nullptr;
    cuGetErrorString(r, &s);
    printf("encode %-44s -> %d (%s)\n", what, (int)r, s ? s : "?");
    return r == CUDA_SUCCESS;
}

int main(int argc, char **argv)
{
    const int test = argc > 1 ? atoi(argv[1]) : 0;
    CK(cudaSetDevice(0));
    CK(cudaFree(0));
    const uint64_t bytes = (argc > 2 ? (uint64_t)atoll(argv[2]) : 2048ull) << 20;
    uint8_t *src, *dst;
    CK(cudaMalloc(&src, bytes + 4096));
    CK(cudaMalloc(&dst, bytes));
    std::vector<uint8_t> h(64 << 20);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    for (uint64_t o = 0; o < bytes; o += h.size()) CK(cudaMemcpy(src + o, h.data(), h.size(), cudaMemcpyHostToDevice));
    const uint64_t nseg = (bytes - 4096) / SEG;
    CUtensorMap tmB, tmC;
    bool okB = false, okC = false;
    if (test == 1) okB = encode(&tmB, src, 511, bytes / 256, 256, 256, SEG / 256, "2-D rows 256 B + 255 overlap, box {256,5}");
    if (test == 2) okB = encode(&tmB, src, 512, bytes / 256, 256, 256, SEG / 256, "2-D rows 256 B + 256 overlap, box {256,5}");
    if (test == 3) okC = encode(&tmC, src, (1 << 20) + 255, bytes >> 20, 1 << 20, 256, 1, "2-D rows 1 MiB + 255 overlap, box {256,1}");
    if (test == 4) okC = encode(&tmC, src, (1 << 20) + 256, bytes >> 20, 1 << 20, 256, 1, "2-D rows 1 MiB + 256 overlap, box {256,1}");
    g_dtype = argc > 6 ? atoi(argv[6]) : 0; g_promo = argc > 7 ? atoi(argv[7]) : 0;
    const uint32_t bx = argc > 8 ? atoi(argv[8]) : 256;
    if (test == 6) okC = encode(&tmC, src, (1 << 20), bytes >> 20, 1 << 20, bx, 1, "2-D rows 1 MiB, box {bx,1}");
    if (test == 5) okC = encode(&tmC, src, (1 << 20), bytes >> 20, 1 << 20, 256, 1, "2-D rows 1 MiB no overlap (rows crossed: zero fill), box {256,1}");
    const int gm = argc > 5 ? atoi(argv[5]) : 0;
    CUtensorMap *d_tm;
    CK(cudaMalloc(&d_tm, 256));
    CK(cudaMemcpy(d_tm, okB ? &tmB : &tmC, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<uint8_t> back(1 << 20);
    auto check = [&](const char *name) {
        CK(cudaMemcpy(back.data(), dst + 1000ull * SEG, back.size(), cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t i = 0; i < (back.size() / SEG) * SEG; ++i) {
            const uint64_t seg = 1000 + i / SEG, soff = seg * SEG + 3 + (seg * 7) % 13 + i % SEG;
            if (back[i] != h[soff % h.size()]) ++bad;
        }
        printf("%-28s verify: %zu wrong bytes\n", name, bad);
    };
    auto run = [&](const char *name, auto launch) {
        CK(cudaMemset(dst, 0, bytes));
        launch(); CK(cudaDeviceSynchronize());
        float best = 1e9f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%-28s %.3f ms  %.0f GB/s copied (read+write %.0f GB/s)\n", name, best, nseg * SEG / best / 1e6, 2.0 * nseg * SEG / best / 1e6);
        check(name);
    };
    const int stop = argc > 4 ? atoi(argv[4]) : 4;
    const int grid = argc > 3 ? atoi(argv[3]) : 148 * 4;
    if (test == 0) run("ldg funnel (today)", [&] { k_ldg<<<grid, WARPS * 32>>>(src, dst, nseg); });
    if (okC) for (int depth = 1; depth <= 2; ++depth) {
        char nm[64]; snprintf(nm, sizeof nm, "tma 5x{256,1} depth %d", depth);
        run(nm, [&] { k_probe<0><<<grid, WARPS * 32>>>(tmC, gm ? d_tm : nullptr, src, dst, nseg, depth, stop); });
    }
    if (okB) for (int depth = 1; depth <= 2; ++depth) {
        char nm[64]; snprintf(nm, sizeof nm, "tma 1x{256,5} depth %d", depth);
        run(nm, [&] { k_probe<1><<<grid, WARPS * 32>>>(tmB, gm ? d_tm : nullptr, src, dst, nseg, depth, stop); });
    }
    return 0;
}
