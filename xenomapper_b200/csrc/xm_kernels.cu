/*
 * xm_kernels.cu -- sm_100a kernels of the read-binning path.
 *
 *   k_scan<Cfg>      one CTA per tile of the secondary stream   (xm_tile.h scan_tile)
 *   k_classify<Cfg>  one CTA per tile of the primary stream     (xm_tile.h classify_tile)
 *
 * plus the device-only pieces the tile code calls: block collectives, the two
 * decoupled look-back chains and the warp copy engine.
 *
 * Tiles are handed out by an atomic ticket, so a CTA only ever waits on tiles
 * whose CTAs have already started: the look-back spins cannot deadlock.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "xm_tile.h"
#include "xm_launch.h"

namespace xm {

/* ---- block collectives --------------------------------------------------- */
/* scratch: scr[0..31] warp totals, scr[32..63] warp exclusive prefixes */
__device__ uint32_t dev_block_scan(uint32_t v, uint32_t *scr, uint32_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();                     /* scratch may still be read by the previous collective */
    if (lane == 31) scr[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nw ? scr[lane] : 0u, s = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        scr[32 + lane] = s - w;
    }
    __syncthreads();
    total = scr[32 + nw - 1] + scr[nw - 1];
    return scr[32 + warp] + x - v;
}

__device__ uint32_t dev_block_min(uint32_t v, uint32_t *scr)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t x = __reduce_min_sync(0xffffffffu, v);
    __syncthreads();
    if (lane == 0) scr[warp] = x;
    __syncthreads();
    uint32_t w = lane < nw ? scr[lane] : 0xffffffffu;
    return __reduce_min_sync(0xffffffffu, w);
}

__device__ uint32_t dev_block_or(uint32_t v, uint32_t *scr)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t x = __reduce_or_sync(0xffffffffu, v);
    __syncthreads();
    if (lane == 0) scr[warp] = x;
    __syncthreads();
    uint32_t w = lane < nw ? scr[lane] : 0u;
    return __reduce_or_sync(0xffffffffu, w);
}

/* ---- look-back chains ------------------------------------------------------ */
__device__ __forceinline__ unsigned long long ld_volatile64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_volatile32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile32(uint32_t *p, uint32_t v)
{
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

/*
 * Chain 1: records per tile plus a stop flag (a blank line ends the stream).
 * One 64-bit word per tile carries status, flag and count, so a single store
 * publishes it.  The fold is "the oldest stop wins": counts add up to and
 * including the first tile that stops.  Called by warp 0.
 */
__device__ void dev_lookback1(unsigned long long *desc, uint32_t tile, unsigned long long agg_count, bool agg_stop,
                              unsigned long long *out)
{
    const int lane = threadIdx.x & 31;
    if (lane == 0) st_volatile64(desc + tile, C1_AGG | (agg_stop ? C1_STOP : 0ull) | agg_count);
    unsigned long long ex_count = 0;
    bool ex_stop = false;
    for (long long j = (long long)tile - 1; j >= 0; j -= 32) {
        const long long idx = j - lane;
        unsigned long long d = C1_INC;                       /* before the first tile: inclusive zero */
        if (idx >= 0) {
            do { d = ld_volatile64(desc + idx); } while ((d >> 62) == 0);
        }
        const unsigned incmask = __ballot_sync(0xffffffffu, (d >> 62) == 2);
        const int L = incmask ? __ffs((int)incmask) - 1 : 31;  /* nearest inclusive prefix */
        const bool part = lane <= L;
        const unsigned stopmask = __ballot_sync(0xffffffffu, part && (d & C1_STOP));
        const int M = stopmask ? 31 - __clz((int)stopmask) : -1;   /* oldest stopping tile in the window */
        unsigned long long c = (part && lane >= (M >= 0 ? M : 0)) ? (d & C1_COUNT) : 0ull;
        c = warp_sum64(c);
        if (M >= 0) { ex_count = c; ex_stop = true; }       /* everything nearer than a stop is dropped */
        else ex_count += c;
        if (incmask) break;
    }
    if (lane == 0) {
        out[0] = ex_count;
        out[1] = ex_stop ? 1ull : 0ull;
        st_volatile64(desc + tile, C1_INC | ((ex_stop || agg_stop) ? C1_STOP : 0ull) | (ex_stop ? ex_count : ex_count + agg_count));
    }
}

/*
 * Chain 2: bytes per output bin (six) + raw primary bytes.  Values first,
 * fence, then the flag.  Called by warp 0 with totals in tot[0..7]; returns
 * the exclusive bases in the same array.
 */
__device__ void dev_lookback2(uint32_t *flag, unsigned long long *agg, unsigned long long *inc, uint32_t tile,
                              unsigned long long *tot)
{
    const int lane = threadIdx.x & 31;
    unsigned long long mine = lane < C2_SLOTS ? tot[lane] : 0ull;
    if (lane < C2_SLOTS) agg[(size_t)tile * C2_SLOTS + lane] = mine;
    __threadfence();
    __syncwarp();
    if (lane == 0) st_volatile32(flag + tile, 1u);
    unsigned long long acc[C2_SLOTS];
#pragma unroll
    for (int b = 0; b < C2_SLOTS; ++b) acc[b] = 0;
    for (long long j = (long long)tile - 1; j >= 0; j -= 32) {
        const long long idx = j - lane;
        uint32_t f = 2u;
        if (idx >= 0) {
            do { f = ld_volatile32(flag + idx); } while (f == 0u);
        }
        __threadfence();
        const unsigned incmask = __ballot_sync(0xffffffffu, f == 2u);
        const int L = incmask ? __ffs((int)incmask) - 1 : 31;
        const bool part = lane <= L && idx >= 0;
        const unsigned long long *src = (f == 2u ? inc : agg) + (size_t)(idx >= 0 ? idx : 0) * C2_SLOTS;
#pragma unroll
        for (int b = 0; b < C2_SLOTS; ++b) {
            unsigned long long v = part ? ld_volatile64(src + b) : 0ull;
            acc[b] += warp_sum64(v);
        }
        if (incmask) break;
    }
    unsigned long long ex = 0;
#pragma unroll
    for (int b = 0; b < C2_SLOTS; ++b) if (lane == b) ex = acc[b];
    if (lane < C2_SLOTS) {
        inc[(size_t)tile * C2_SLOTS + lane] = ex + mine;
        tot[lane] = ex;
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) st_volatile32(flag + tile, 2u);
}

/* ---- warp copy engine ------------------------------------------------------- */
/*
 * Copies len bytes to an arbitrarily aligned global destination with 16-byte
 * aligned vector stores.  The source is the staged window in shared memory
 * when src_smem is set, global memory otherwise.  Source and destination
 * misalignment differ, so each destination vector is assembled from two
 * aligned source vectors with funnel shifts; the shift is the same for every
 * vector of one copy, hence warp-uniform.
 */
__device__ __forceinline__ uint4 ld_src16(const uint8_t *p, bool smem)
{
    uint4 v;
    if (smem) v = *(const uint4 *)p;
    else asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ void dev_warp_copy(uint8_t *dst, const uint8_t *src_smem, const uint8_t *src_glob, uint32_t len)
{
    const int lane = threadIdx.x & 31;
    const bool smem = src_smem != nullptr;
    const uint8_t *src = smem ? src_smem : src_glob;
    uint32_t head = (uint32_t)((16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u);
    if (head > len) head = len;
    if ((uint32_t)lane < head) dst[lane] = src[lane];
    const uint8_t *s = src + head;
    uint8_t *d = dst + head;
    const uint32_t rest = len - head;
    const uint32_t nchunk = rest >> 4;
    const uint32_t u = (uint32_t)((uintptr_t)s & 15u);       /* shared window base is 16-byte aligned */
    const uint8_t *sa = s - u;
    const uint32_t bsh = (u & 3u) * 8u;
    const uint32_t wsh = u >> 2;
    for (uint32_t c = (uint32_t)lane; c < nchunk; c += 32) {
        const uint4 q0 = ld_src16(sa + 16u * c, smem);
        uint4 o;
        if (u == 0) o = q0;
        else {
            const uint4 q1 = ld_src16(sa + 16u * c + 16u, smem);
            switch (wsh) {
            case 0: o.x = __funnelshift_r(q0.x, q0.y, bsh); o.y = __funnelshift_r(q0.y, q0.z, bsh); o.z = __funnelshift_r(q0.z, q0.w, bsh); o.w = __funnelshift_r(q0.w, q1.x, bsh); break;
            case 1: o.x = __funnelshift_r(q0.y, q0.z, bsh); o.y = __funnelshift_r(q0.z, q0.w, bsh); o.z = __funnelshift_r(q0.w, q1.x, bsh); o.w = __funnelshift_r(q1.x, q1.y, bsh); break;
            case 2: o.x = __funnelshift_r(q0.z, q0.w, bsh); o.y = __funnelshift_r(q0.w, q1.x, bsh); o.z = __funnelshift_r(q1.x, q1.y, bsh); o.w = __funnelshift_r(q1.y, q1.z, bsh); break;
            default: o.x = __funnelshift_r(q0.w, q1.x, bsh); o.y = __funnelshift_r(q1.x, q1.y, bsh); o.z = __funnelshift_r(q1.y, q1.z, bsh); o.w = __funnelshift_r(q1.z, q1.w, bsh); break;
            }
        }
        *(uint4 *)(d + 16u * c) = o;
    }
    const uint32_t tail = rest & 15u;
    if ((uint32_t)lane < tail) d[16u * nchunk + lane] = s[16u * nchunk + lane];
}

/* ---- kernels ------------------------------------------------------------------ */
template <class C>
__global__ void __launch_bounds__(C::THREADS) k_scan(const ScanArgs a)
{
    extern __shared__ uint4 xm_smem[];
    __shared__ uint32_t s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(&a.g->ticket[a.stream_id], 1u);
    __syncthreads();
    TileCtx<C> T;
    T.m = carve<C>(xm_smem);
    T.emu = nullptr;
    scan_tile<C>(T, a, s_tile);
}

template <class C>
__global__ void __launch_bounds__(C::THREADS) k_classify(const ClassifyArgs a)
{
    extern __shared__ uint4 xm_smem[];
    __shared__ uint32_t s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(&a.g->ticket[0], 1u);
    __syncthreads();
    TileCtx<C> T;
    T.m = carve<C>(xm_smem);
    T.emu = nullptr;
    classify_tile<C>(T, a, s_tile);
}

/* ---- launchers ------------------------------------------------------------------ */
template <class C>
static cudaError_t launch_scan_t(const ScanArgs &a, cudaStream_t st)
{
    const size_t smem = TileLayout<C>::total;
    cudaError_t e = cudaFuncSetAttribute(k_scan<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_scan<C><<<a.ntiles, C::THREADS, smem, st>>>(a);
    return cudaGetLastError();
}
template <class C>
static cudaError_t launch_classify_t(const ClassifyArgs &a, cudaStream_t st)
{
    const size_t smem = TileLayout<C>::total;
    cudaError_t e = cudaFuncSetAttribute(k_classify<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_classify<C><<<a.ntiles, C::THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

uint32_t tile_bytes(bool small) { return small ? CfgSmall::TILE : CfgBig::TILE; }

cudaError_t launch_scan(const ScanArgs &a, bool small, cudaStream_t st)
{
    return small ? launch_scan_t<CfgSmall>(a, st) : launch_scan_t<CfgBig>(a, st);
}
cudaError_t launch_classify(const ClassifyArgs &a, bool small, cudaStream_t st)
{
    return small ? launch_classify_t<CfgSmall>(a, st) : launch_classify_t<CfgBig>(a, st);
}

/* replicate a block of bytes `times` times back to back (bench workloads) -- plain strided copy */
__global__ void k_fill_u64(unsigned long long *p, unsigned long long v, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
cudaError_t launch_fill_u64(unsigned long long *p, unsigned long long v, size_t n, cudaStream_t st)
{
    if (!n) return cudaSuccess;
    k_fill_u64<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, v, n);
    return cudaGetLastError();
}

}  // namespace xm
