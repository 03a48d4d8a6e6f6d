/*
 * xm_kernels.cu -- sm_100a kernels of the read-binning path.
 *
 *   k_scan<Cfg>      one CTA per tile of the secondary stream   (xm_tile.h scan_tile)
 *   k_classify<Cfg>  one CTA per tile of the primary stream     (xm_tile.h classify_tile)
 *
 * plus the device-only pieces the tile code calls: window staging by TMA bulk
 * copy, block collectives, the two decoupled look-back chains and the copy
 * engines.
 *
 * Tile = blockIdx.x.  The look-back spins wait only on lower-numbered tiles,
 * which the hardware dispatches first (the ordering every single-pass
 * decoupled look-back scan relies on).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "xm_tile.h"
#include "xm_launch.h"

namespace xm {

/* ---- block collectives --------------------------------------------------- */
/* One barrier each.  wt: one word per warp; consecutive collectives alternate between two banks so that a
 * slow reader of one never meets the next one's writes. */
__device__ uint32_t dev_block_scan(uint32_t v, uint32_t *wt, uint32_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wt[warp] = x;
    __syncthreads();
    /* every warp scans the (at most 32) warp totals for itself */
    const uint32_t t = lane < nw ? wt[lane] : 0u;
    uint32_t s = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
    }
    total = __shfl_sync(0xffffffffu, s, 31);
    const uint32_t pre = __shfl_sync(0xffffffffu, s - t, warp);
    return pre + x - v;
}

__device__ uint32_t dev_block_min(uint32_t v, uint32_t *wt)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t x = __reduce_min_sync(0xffffffffu, v);
    if (lane == 0) wt[warp] = x;
    __syncthreads();
    return __reduce_min_sync(0xffffffffu, lane < nw ? wt[lane] : 0xffffffffu);
}

/* Exclusive offsets of each thread's bytes inside its bin, over the whole CTA: warp scans per bin, per-warp
 * totals in wt[warp][8] (slot 6 = raw input bytes), one barrier, then the prefix over the warps before. */
__device__ uint32_t dev_scan_bins(uint32_t bin, uint32_t bytes, uint32_t raw, uint32_t *wt)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t mine = 0;
#pragma unroll
    for (uint32_t b = 0; b < 6; ++b) {
        const uint32_t v = bin == b ? bytes : 0u;
        uint32_t x = 0;
        if (__any_sync(0xffffffffu, v != 0u)) {
            x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (bin == b) mine = x - v;
        }
        if (lane == 31) wt[warp * 8 + b] = x;
    }
    const uint32_t r = __reduce_add_sync(0xffffffffu, raw);
    if (lane == 31) { wt[warp * 8 + 6] = r; wt[warp * 8 + 7] = 0; }
    __syncthreads();
    if (bin < 6u)
        for (int w = 0; w < warp; ++w) mine += wt[w * 8 + bin];
    return mine;
}
/* the CTA's totals per slot, 64-bit; called by warp 0 after dev_scan_bins */
__device__ void dev_bin_totals(const uint32_t *wt, unsigned long long *tot)
{
    const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (lane < C2_SLOTS) {
        unsigned long long t = 0;
        for (int w = 0; w < nw; ++w) t += wt[w * 8 + lane];
        tot[lane] = t;
    }
    __syncwarp();
}

/* warp-aggregated histogram increment: one shared-memory atomic per distinct key in the warp */
__device__ void dev_hist_add(uint32_t *hist, uint32_t key)
{
    if (__all_sync(0xffffffffu, key >= 36u)) return;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (key < 36u && (int)(threadIdx.x & 31) == __ffs((int)peers) - 1) atomicAdd(&hist[key], (uint32_t)__popc(peers));
}

/* ---- window staging: one TMA bulk copy per tile -------------------------------- */
__device__ void dev_stage_window(uint8_t *win, const uint8_t *src, uint32_t bytes, unsigned long long *mbar)
{
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(win);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(0u) : "memory");
    }
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
 * publishes it and a single load reads it.  The fold is "the oldest stop
 * wins": counts add up to and including the first tile that stops.
 * A tile publishes its own count as early as it knows it (dev_publish1) and
 * resolves its exclusive prefix later (dev_resolve1, warp 0): 256 predecessor
 * descriptors are fetched per round trip, so a whole wave of concurrently
 * running tiles is crossed in two or three steps.
 */
__device__ void dev_publish1(unsigned long long *desc, uint32_t tile, unsigned long long agg_count, bool agg_stop)
{
    st_volatile64(desc + tile, C1_AGG | (agg_stop ? C1_STOP : 0ull) | agg_count);
}

__device__ void dev_resolve1(unsigned long long *desc, uint32_t tile, unsigned long long agg_count, bool agg_stop,
                             unsigned long long *out)
{
    const int lane = threadIdx.x & 31;
    unsigned long long ex_count = 0;
    bool ex_stop = false, done = false;
    for (long long j = (long long)tile - 1; j >= 0 && !done; j -= 256) {
        unsigned long long d[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const long long idx = j - 32 * q - lane;
            d[q] = idx >= 0 ? ld_volatile64(desc + idx) : C1_INC;      /* before the first tile: inclusive zero */
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (done) break;
            const long long idx = j - 32 * q - lane;
            if (idx >= 0)
                while ((d[q] >> 62) == 0) { d[q] = ld_volatile64(desc + idx); }
            const unsigned incmask = __ballot_sync(0xffffffffu, (d[q] >> 62) == 2);
            const int L = incmask ? __ffs((int)incmask) - 1 : 31;  /* nearest inclusive prefix */
            const bool part = lane <= L;
            const unsigned stopmask = __ballot_sync(0xffffffffu, part && (d[q] & C1_STOP));
            const int M = stopmask ? 31 - __clz((int)stopmask) : -1;   /* oldest stopping tile in the window */
            unsigned long long c = (part && lane >= (M >= 0 ? M : 0)) ? (d[q] & C1_COUNT) : 0ull;
            c = warp_sum64(c);
            if (M >= 0) { ex_count = c; ex_stop = true; }       /* everything nearer than a stop is dropped */
            else ex_count += c;
            if (incmask) done = true;
        }
    }
    if (lane == 0) {
        out[0] = ex_count;
        out[1] = ex_stop ? 1ull : 0ull;
        st_volatile64(desc + tile, C1_INC | ((ex_stop || agg_stop) ? C1_STOP : 0ull) | (ex_stop ? ex_count : ex_count + agg_count));
    }
}

/*
 * Chain 2: bytes per output bin.  Six independent chains share one 64-byte
 * row per tile; every word carries its own status, so rows need no fence and
 * no separate flag.  Called by warp 0.
 */
__device__ void dev_publish2(unsigned long long *chain, uint32_t tile, const unsigned long long *tot)
{
    const int lane = threadIdx.x & 31;
    if (lane < 6) st_volatile64(chain + (size_t)tile * C2_SLOTS + lane, C2_AGG | tot[lane]);
}

__device__ void dev_resolve2(unsigned long long *chain, uint32_t tile, unsigned long long *tot, unsigned long long *dbg)
{
#ifdef XM_PHASE_TIMING
    unsigned long long n_win = 0, n_spin = 0;   /* windows walked, polls that found nothing */
#else
    (void)dbg;
#endif
    const int lane = threadIdx.x & 31;
    unsigned long long acc[6];
#pragma unroll
    for (int b = 0; b < 6; ++b) acc[b] = 0;
    unsigned pending = 0x3fu;                              /* bins still looking for an inclusive prefix */
    for (long long j = (long long)tile - 1; j >= 0 && pending; j -= 32) {
#ifdef XM_PHASE_TIMING
        ++n_win;
#endif
        unsigned long long d[1][6];
#pragma unroll
        for (int q = 0; q < 1; ++q) {
            const long long idx = j - 32 * q - lane;
#pragma unroll
            for (int b = 0; b < 6; ++b) d[q][b] = idx >= 0 ? ld_volatile64(chain + (size_t)idx * C2_SLOTS + b) : C2_INC;
        }
#pragma unroll
        for (int q = 0; q < 1; ++q) {
            if (!pending) break;
            const long long idx = j - 32 * q - lane;
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                if (!((pending >> b) & 1u)) continue;
                if (idx >= 0)
                    while ((d[q][b] >> 62) == 0) {
                        d[q][b] = ld_volatile64(chain + (size_t)idx * C2_SLOTS + b);
#ifdef XM_PHASE_TIMING
                        ++n_spin;
#endif
                    }
                const unsigned incmask = __ballot_sync(0xffffffffu, (d[q][b] >> 62) == 2);
                const int L = incmask ? __ffs((int)incmask) - 1 : 31;
                acc[b] += warp_sum64(lane <= L ? (d[q][b] & C2_VAL) : 0ull);
                if (incmask) pending &= ~(1u << b);
            }
        }
    }
    unsigned long long ex = 0;
#pragma unroll
    for (int b = 0; b < 6; ++b) if (lane == b) ex = acc[b];
    if (lane < 6) {
        st_volatile64(chain + (size_t)tile * C2_SLOTS + lane, C2_INC | (ex + tot[lane]));
        tot[lane] = ex;
    }
#ifdef XM_PHASE_TIMING
    if (dbg) {
        if (lane == 0) atomicAdd(dbg, n_win);
        n_spin = warp_sum64(n_spin);
        if (lane == 0) atomicAdd(dbg + 1, n_spin);
    }
#endif
    __syncwarp();
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

#ifndef XM_COPY_ROWS
#define XM_COPY_ROWS 4      /* 512-byte rows a lane keeps in flight per trip of the copy */
#endif
#ifndef XM_WHATIF
#define XM_WHATIF 0      /* timing experiments that break the output: 1 copy without source loads, 2 no copy, 3 no aux parse, 4 no QNAME hash */
#endif
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

/*
 * Copies len bytes from the staged window to an arbitrarily aligned global destination: unaligned head and
 * tail bytes by single lanes, the 16-byte aligned body as vector stores assembled from two aligned shared-memory
 * vectors with funnel shifts (the shift is the same for the whole piece).  Called by a full warp.
 */
__device__ void dev_copy_piece(uint8_t *dst, const uint8_t *win, uint32_t src_off, uint32_t len)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t head = (16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u;
    if (head > len) head = len;
    const uint32_t rest = len - head, nchunk = rest >> 4, tail = rest & 15u;
    const uint32_t so = src_off + head;
    const uint32_t u = so & 15u;
    const uint8_t *sa = win + (so - u);
    uint8_t *body = dst + head;
    if (lane < head) dst[lane] = win[src_off + lane];
    if (lane < tail) body[16u * nchunk + lane] = win[so + 16u * nchunk + lane];
    const uint32_t bsh = (u & 3u) * 8u;
    if (u == 0) {
        for (uint32_t c = lane; c < nchunk; c += 32u)
            __stcs((uint4 *)(body + 16u * c), *(const uint4 *)(sa + 16u * c));
        return;
    }
    for (uint32_t c = lane; c < nchunk; c += 32u) {
        const uint4 q0 = *(const uint4 *)(sa + 16u * c);
        const uint4 q1 = *(const uint4 *)(sa + 16u * c + 16u);
        uint4 o;
        switch (u >> 2) {
        case 0: o.x = __funnelshift_r(q0.x, q0.y, bsh); o.y = __funnelshift_r(q0.y, q0.z, bsh); o.z = __funnelshift_r(q0.z, q0.w, bsh); o.w = __funnelshift_r(q0.w, q1.x, bsh); break;
        case 1: o.x = __funnelshift_r(q0.y, q0.z, bsh); o.y = __funnelshift_r(q0.z, q0.w, bsh); o.z = __funnelshift_r(q0.w, q1.x, bsh); o.w = __funnelshift_r(q1.x, q1.y, bsh); break;
        case 2: o.x = __funnelshift_r(q0.z, q0.w, bsh); o.y = __funnelshift_r(q0.w, q1.x, bsh); o.z = __funnelshift_r(q1.x, q1.y, bsh); o.w = __funnelshift_r(q1.y, q1.z, bsh); break;
        default: o.x = __funnelshift_r(q0.w, q1.x, bsh); o.y = __funnelshift_r(q1.x, q1.y, bsh); o.z = __funnelshift_r(q1.y, q1.z, bsh); o.w = __funnelshift_r(q1.z, q1.w, bsh); break;
        }
        __stcs((uint4 *)(body + 16u * c), o);
    }
}

/*
 * Copies len bytes between arbitrarily aligned global addresses, called by a full warp: unaligned head and tail
 * bytes by single lanes, the 16-byte aligned body of the destination as vector stores assembled from two aligned
 * source vectors with funnel shifts (the shift is the same for the whole piece).  The source was read by this
 * kernel a few hundred tiles ago and normally still sits in L2; it is readable up to the next multiple of 16.
 */
/* the aligned body of a copy whose source sits WSH words and bsh bits past a 16-byte boundary: four rows of 32
 * chunks per trip, all eight loads of a lane requested before the first is used */
#if XM_WHATIF == 5
__device__ __forceinline__ void xm_st_fake(uint4 *p, uint4 v) { if ((v.x ^ v.y ^ v.z ^ v.w) == 0x9e3779b9u) *p = v; }
#define XM_ST(p, v) xm_st_fake(p, v)
#elif XM_WHATIF == 6
#define XM_ST(p, v) (*(p) = (v))
#else
#define XM_ST(p, v) __stcs(p, v)
#endif
#ifndef XM_COPY_SHFL_ROWS
#define XM_COPY_SHFL_ROWS 0      /* > 0: the experiment below, that many rows per trip (measured slower: 7.6 / 8.0 / 10.5 ms against 6.9 ms at 4 / 6 / 8 rows, profiles/r02_kernel_experiments.md) */
#endif
#if XM_COPY_SHFL_ROWS > 0
/* Experiment (off by default): every destination chunk needs source vectors V[c] and V[c + 1]; lane L of row r holds
 * V[c], and V[c + 1] sits in lane L + 1 of the same row -- for lane 31 in lane 0 of the next row.  Each lane loads one
 * vector per chunk and takes the words it needs of the next one by a rotating shuffle to which lane 0 contributes its
 * next row's vector: half the load instructions and L1 wavefronts, more rows in flight per trip -- and slower, the
 * shuffles sit on the dependent path between the load and the store. */
template <int WSH>
__device__ __forceinline__ void copy_body(uint8_t *body, const uint8_t *sa, uint32_t nchunk, uint32_t bsh, uint32_t lane)
{
    constexpr int R = XM_COPY_SHFL_ROWS;
    auto ld = [](const uint8_t *q) { return ld_src16(q, false); };
    const int nxt = (int)((lane + 1u) & 31u);
#pragma unroll 1
    for (uint32_t cb = 0; cb < nchunk; cb += 32u * R) {
        uint4 a[R + 1];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t c = cb + 32u * r + lane;
            a[r] = ld(sa + 16u * (c < nchunk ? c : nchunk));            /* V[nchunk] is the last vector any chunk needs */
        }
        {
            const uint32_t c = cb + 32u * R;
            a[R] = ld(sa + 16u * (c < nchunk ? c : nchunk));            /* one address for the whole warp */
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t c = cb + 32u * r + lane;
            if (cb + 32u * r >= nchunk) break;                            /* uniform: the trip's last rows may be empty */
            const uint4 q0 = a[r];
            uint4 q1;
            q1.x = __shfl_sync(0xffffffffu, lane == 0 ? a[r + 1].x : q0.x, nxt);
            if (WSH >= 1) q1.y = __shfl_sync(0xffffffffu, lane == 0 ? a[r + 1].y : q0.y, nxt);
            if (WSH >= 2) q1.z = __shfl_sync(0xffffffffu, lane == 0 ? a[r + 1].z : q0.z, nxt);
            if (WSH >= 3) q1.w = __shfl_sync(0xffffffffu, lane == 0 ? a[r + 1].w : q0.w, nxt);
            uint4 o;
            if (WSH == 0) { o.x = __funnelshift_r(q0.x, q0.y, bsh); o.y = __funnelshift_r(q0.y, q0.z, bsh); o.z = __funnelshift_r(q0.z, q0.w, bsh); o.w = __funnelshift_r(q0.w, q1.x, bsh); }
            else if (WSH == 1) { o.x = __funnelshift_r(q0.y, q0.z, bsh); o.y = __funnelshift_r(q0.z, q0.w, bsh); o.z = __funnelshift_r(q0.w, q1.x, bsh); o.w = __funnelshift_r(q1.x, q1.y, bsh); }
            else if (WSH == 2) { o.x = __funnelshift_r(q0.z, q0.w, bsh); o.y = __funnelshift_r(q0.w, q1.x, bsh); o.z = __funnelshift_r(q1.x, q1.y, bsh); o.w = __funnelshift_r(q1.y, q1.z, bsh); }
            else { o.x = __funnelshift_r(q0.w, q1.x, bsh); o.y = __funnelshift_r(q1.x, q1.y, bsh); o.z = __funnelshift_r(q1.y, q1.z, bsh); o.w = __funnelshift_r(q1.z, q1.w, bsh); }
            if (c < nchunk) XM_ST((uint4 *)(body + 16u * c), o);
        }
    }
}
#else
template <int WSH>
__device__ __forceinline__ void copy_body(uint8_t *body, const uint8_t *sa, uint32_t nchunk, uint32_t bsh, uint32_t lane)
{
#if XM_WHATIF == 1 || XM_WHATIF == 8
    auto ld = [](const uint8_t *q) { return make_uint4((uint32_t)(uintptr_t)q, 1u, 2u, 3u); };
#else
    auto ld = [](const uint8_t *q) { return ld_src16(q, false); };
#endif
    auto shift = [bsh](const uint4 &q0, const uint4 &q1) {
        uint4 o;
        if (WSH == 0) { o.x = __funnelshift_r(q0.x, q0.y, bsh); o.y = __funnelshift_r(q0.y, q0.z, bsh); o.z = __funnelshift_r(q0.z, q0.w, bsh); o.w = __funnelshift_r(q0.w, q1.x, bsh); }
        else if (WSH == 1) { o.x = __funnelshift_r(q0.y, q0.z, bsh); o.y = __funnelshift_r(q0.z, q0.w, bsh); o.z = __funnelshift_r(q0.w, q1.x, bsh); o.w = __funnelshift_r(q1.x, q1.y, bsh); }
        else if (WSH == 2) { o.x = __funnelshift_r(q0.z, q0.w, bsh); o.y = __funnelshift_r(q0.w, q1.x, bsh); o.z = __funnelshift_r(q1.x, q1.y, bsh); o.w = __funnelshift_r(q1.y, q1.z, bsh); }
        else { o.x = __funnelshift_r(q0.w, q1.x, bsh); o.y = __funnelshift_r(q1.x, q1.y, bsh); o.z = __funnelshift_r(q1.y, q1.z, bsh); o.w = __funnelshift_r(q1.z, q1.w, bsh); }
        return o;
    };
    uint32_t c = lane;
    /* XM_COPY_ROWS rows per trip, the last trip with fewer: every load of a trip is requested before the first is used */
    for (; c < nchunk; c += 32u * XM_COPY_ROWS) {
        uint4 a0[XM_COPY_ROWS], a1[XM_COPY_ROWS];
#pragma unroll
        for (int r = 0; r < XM_COPY_ROWS; ++r) {
            const uint32_t cr = c + 32u * r < nchunk ? c + 32u * r : c;       /* rows past the end re-read this lane's first chunk */
            a0[r] = ld(sa + 16u * cr); a1[r] = ld(sa + 16u * cr + 16u);
        }
#pragma unroll
        for (int r = 0; r < XM_COPY_ROWS; ++r)
            if (c + 32u * r < nchunk) XM_ST((uint4 *)(body + 16u * (c + 32u * r)), shift(a0[r], a1[r]));
    }
}

#endif

__device__ void dev_copy_global(uint8_t *dst, const uint8_t *src, uint32_t len)
{
#if XM_WHATIF == 2
    return;
#endif
    const uint32_t lane = threadIdx.x & 31;
    uint32_t head = (16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u;
    if (head > len) head = len;
    const uint32_t rest = len - head, nchunk = rest >> 4, tail = rest & 15u;
    const uint8_t *so = src + head;
    const uint32_t u = (uint32_t)((uintptr_t)so & 15u);
    const uint8_t *sa = so - u;
    uint8_t *body = dst + head;
    /* the unaligned head bytes by lanes 0-15, the tail bytes by lanes 16-31: one load instruction, requested before
     * the body's loads, and one store instruction behind the body -- the copy waits for L2 once, not twice */
    const uint32_t hi = lane & 15u;
    const bool is_tail = lane >= 16u;
    const bool has_byte = hi < (is_tail ? tail : head);
    const size_t ho = is_tail ? (size_t)head + 16u * (size_t)nchunk + hi : (size_t)hi;
    uint8_t hb = 0;
    if (has_byte) hb = src[ho];
    const uint32_t bsh = (u & 3u) * 8u;
    if (u == 0) {
#pragma unroll 4
        for (uint32_t c = lane; c < nchunk; c += 32u)
            __stcs((uint4 *)(body + 16u * c), ld_src16(sa + 16u * c, false));
    } else {
        switch (u >> 2) {
        case 0: copy_body<0>(body, sa, nchunk, bsh, lane); break;
        case 1: copy_body<1>(body, sa, nchunk, bsh, lane); break;
        case 2: copy_body<2>(body, sa, nchunk, bsh, lane); break;
        default: copy_body<3>(body, sa, nchunk, bsh, lane); break;
        }
    }
    if (has_byte) dst[ho] = hb;
}
}  // namespace xm
#include "xm_scan2.cuh"
#include "xm_emit.cuh"
namespace xm {

/* ---- kernels ------------------------------------------------------------------ */
template <class C>
__global__ void __launch_bounds__(C::THREADS, (C::THREADS == XM_BIG_THREADS ? XM_BIG_OCC : 4)) k_scan(const ScanArgs a)
{
    extern __shared__ uint4 xm_smem[];
    TileCtx<C> T;
    T.m = carve<C>(xm_smem);
    T.emu = nullptr;
    scan_tile<C>(T, a, blockIdx.x);
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, (C::THREADS == XM_BIG_THREADS ? XM_BIG_OCC : 4)) k_classify(const ClassifyArgs a)
{
    extern __shared__ uint4 xm_smem[];
    TileCtx<C> T;
    T.m = carve<C>(xm_smem);
    T.emu = nullptr;
    classify_tile<C>(T, a, blockIdx.x);
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

__global__ void k_fill_u64(unsigned long long *p, unsigned long long v, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
/* p[i] += d (mod 2^64): received rows are moved to where their text landed (xm_shard.h) */
__global__ void k_add_u64(unsigned long long *p, unsigned long long d, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] += d;
}
cudaError_t launch_add_u64(unsigned long long *p, unsigned long long d, size_t n, cudaStream_t st)
{
    if (!n) return cudaSuccess;
    k_add_u64<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, d, n);
    return cudaGetLastError();
}
cudaError_t launch_fill_u64(unsigned long long *p, unsigned long long v, size_t n, cudaStream_t st)
{
    if (!n) return cudaSuccess;
    k_fill_u64<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, v, n);
    return cudaGetLastError();
}

}  // namespace xm
