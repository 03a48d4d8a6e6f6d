/*
 * xm_scan2.cuh -- the secondary-stream scan without CTA-wide phases (included by xm_kernels.cu).
 *
 * k_scan (xm_tile.h) stages a window per CTA and walks it in barrier-separated
 * phases; most warps spend most of their time at those barriers.  Here every
 * warp walks its own span of the stream: it reads the span from global memory
 * with two coalesced 16-byte loads per lane and step (requested one step ahead), builds the same two byte
 * class masks, keeps them (and nothing else) in a small private piece of shared
 * memory, lists the line starts, and parses one line per lane with the same
 * fast_head / fast_tail as k_scan, reading the few bytes a line's parse looks at
 * straight from L1/L2.  The warps of a CTA meet twice per tile, to add up their
 * record counts for look-back chain 1 and to pick up the tile's record base.
 *
 * It is the clean-input path only: a span that holds anything the mask-driven
 * parse does not handle (a terminator candidate that is not '\n', two touching
 * separators, a line longer than the window extension, an unterminated last
 * line, more lines than the queue holds) raises Globals::fallback and the host
 * runs k_scan instead, which is exact for every input.  Same outputs as k_scan:
 * SCompact rows, n_stream / end_off.
 */
#pragma once

namespace xm {

template <int WARPS_, int SPAN_, int BACK_, int EXT_>
struct Scan2Cfg {
    static constexpr int WARPS = WARPS_, SPAN = SPAN_, BACK = BACK_, EXT = EXT_;
    static constexpr int TILE = WARPS * SPAN;             /* bytes of the stream one CTA owns */
    static constexpr int WINB = BACK + SPAN + EXT;          /* bytes a warp may look at */
    static constexpr int NWW = WINB / 32 + 1;               /* mask words per warp */
    static constexpr int LQ = 160;                          /* line starts a warp can list */
    static_assert(SPAN % 1024 == 0 && BACK % 1024 == 0 && EXT % 1024 == 0 && WINB < 65535, "geometry");
};
#ifndef XM_SCAN2_WARPS
#define XM_SCAN2_WARPS 8
#endif
#ifndef XM_SCAN2_SPAN
#define XM_SCAN2_SPAN 10240
#endif
#ifndef XM_SCAN2_OCC
#define XM_SCAN2_OCC 4
#endif
using Scan2Big = Scan2Cfg<XM_SCAN2_WARPS, XM_SCAN2_SPAN, 1024, 2048>;

template <class C>
__global__ void __launch_bounds__(C::WARPS * 32, XM_SCAN2_OCC) k_scan2(const ScanArgs a)
{
    __shared__ uint32_t s_tbm[C::WARPS][C::NWW], s_nlm[C::WARPS][C::NWW];
    __shared__ uint16_t s_trk[C::WARPS][C::NWW];
    __shared__ uint16_t s_start[C::WARPS][C::LQ];
    __shared__ uint32_t s_cnt[C::WARPS];
    __shared__ unsigned long long s_base[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x;
    const uint64_t span_lo = ((uint64_t)tile * C::WARPS + (uint64_t)warp) * (uint64_t)C::SPAN;
    const bool live = span_lo < a.S.len;
    const bool skip = a.skip != 0;
    bool bad = false;                                 /* this span needs the exact kernel */

    uint32_t *tbm = s_tbm[warp], *nlm = s_nlm[warp];
    uint16_t *trk = s_trk[warp], *starts = s_start[warp];
    /* window: [win0, win0 + wbytes); the span's first byte sits at hoff.  The bytes before it are looked at for the
     * line that precedes the span (its last newline decides whether the span opens a line; skipping needs its QNAME) */
    const uint64_t win0 = live ? (span_lo >= (uint64_t)C::BACK ? span_lo - C::BACK : 0) : 0;
    const uint32_t hoff = (uint32_t)(span_lo - win0);
    uint64_t wend = span_lo + C::SPAN + C::EXT;
    if (wend > a.S.len) wend = a.S.len;
    const uint32_t wbytes = live ? (uint32_t)(wend - win0) : 0u;
    const uint32_t own_hi = hoff + C::SPAN < wbytes ? hoff + (uint32_t)C::SPAN : wbytes;     /* owned starts lie in [hoff, own_hi) */
    const uint8_t *win = a.S.p + win0;          /* read through L1/L2: no staged copy, so a SM holds 32 of these warps */

    /* ---- masks and line starts ------------------------------------------------------------------ */
    uint32_t nst = 0;                                 /* line starts listed so far (uniform) */
    uint32_t tab_run = 0;
    bool adj = false;
    if (live) {
        if (win0 == 0) { if (lane == 0) starts[0] = 0; nst = 1; }      /* the stream's first byte opens a line */
        const uint32_t w_first = skip ? 0u : (hoff >= 32u ? hoff / 32u - 1u : 0u);
        const uint32_t nwords = (wbytes + 31u) >> 5;
        uint32_t carryW = 0;                          /* was the byte before this word a separator? (lane 0's view) */
        bool done = false;
        const uint32_t lim16 = (wbytes + 15u) & ~15u;      /* the buffer is readable up to the next multiple of 16 */
        const uint4 filler = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
        uint4 va = filler, vb = filler, na = filler, nb = filler;
        if (w_first + (uint32_t)lane < nwords) {
            const uint32_t off = (w_first + (uint32_t)lane) * 32u;
            va = ld_src16(win + off, false);
            if (off + 16u < lim16) vb = ld_src16(win + off + 16u, false);
        }
        for (uint32_t wb = w_first; wb < nwords && !done; wb += 32u) {
            const uint32_t w = wb + (uint32_t)lane;
            /* the next step's bytes are requested before this step's are looked at */
            if (w + 32u < nwords) {
                const uint32_t off = (w + 32u) * 32u;
                na = ld_src16(win + off, false);
                nb = off + 16u < lim16 ? ld_src16(win + off + 16u, false) : filler;
            }
            uint32_t W = 0, Tm = 0;
            if (w < nwords) {
                const uint32_t off = w * 32u;
                uint32_t Wa, Ta, Wb, Tb;
                masks16(va, Wa, Ta);
                masks16(vb, Wb, Tb);
                W = Wa | (Wb << 16);
                Tm = Ta | (Tb << 16);
                const uint32_t valid = wbytes - off;
                if (valid < 32u) { const uint32_t k = (1u << valid) - 1u; W &= k; Tm &= k; }
            }
            const uint32_t N = W & ~Tm;
            uint32_t prevW = __shfl_up_sync(0xffffffffu, W >> 31, 1);
            if (lane == 0) prevW = carryW;
            carryW = __shfl_sync(0xffffffffu, W >> 31, 31);
            if (W & ((W << 1) | prevW)) adj = true;
            /* tabs and line ends before each lane's word: one packed warp scan */
            const uint32_t x = (uint32_t)__popc(Tm) | ((uint32_t)__popc(N) << 16);
            uint32_t inc = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
            const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
            const uint32_t exc = inc - x;
            if (w < (uint32_t)C::NWW) { tbm[w] = Tm; nlm[w] = N; trk[w] = (uint16_t)(tab_run + (exc & 0xffffu)); }
            uint32_t idx = nst + (exc >> 16);
            for (uint32_t m = N; m; m &= m - 1) {
                const uint32_t p = w * 32u + (uint32_t)__ffs((int)m);          /* the byte after the terminator */
                if (idx < (uint32_t)C::LQ) starts[idx] = (uint16_t)p;
                ++idx;
            }
            tab_run += tot & 0xffffu;
            nst += tot >> 16;
            va = na; vb = nb;
            /* the last owned line is closed once a terminator at or beyond the span's last byte has been seen */
            const uint32_t lastp = own_hi - 1u;
            const bool closed = N && ((w * 32u + 31u - (uint32_t)__clz((int)N)) >= lastp);
            done = __any_sync(0xffffffffu, closed);
        }
        adj = __any_sync(0xffffffffu, adj);
        if (nst > (uint32_t)C::LQ || adj) bad = true;
    }
    __syncwarp();

    /* ---- the lines this span owns: starts in [hoff, own_hi) -------------------------------------------- */
    uint32_t j0 = 0, j1 = 0;
    if (live && !bad) {
        for (uint32_t k = (uint32_t)lane; k < ((nst + 31u) & ~31u); k += 32u) {
            const uint32_t s = k < nst ? (uint32_t)starts[k] : 0xffffffffu;
            j0 += (uint32_t)__popc(__ballot_sync(0xffffffffu, s < hoff));
            j1 += (uint32_t)__popc(__ballot_sync(0xffffffffu, s < own_hi));
        }
        /* every owned line needs its terminator listed (the next start); at the end of the stream the last line
         * must be terminated, and nothing may follow the last listed start except the end of the data */
        if (j1 > j0) {
            if (j1 >= nst) bad = true;
        }
        if (wend == a.S.len && !bad) {
            const uint32_t last = nst ? (uint32_t)starts[nst - 1] : 0xffffffffu;
            if (last != wbytes && own_hi == wbytes) bad = true;       /* unterminated last line */
        }
        if (skip && j1 > j0 && j0 == 0 && win0 + starts[0] != 0) bad = true;      /* the line before the span is not in the window */
    }
    const uint32_t nown = (live && !bad) ? j1 - j0 : 0u;

    /* ---- parse: one line per lane ---------------------------------------------------------------- */
    /* skipping walks also parse the head of the line before the first owned one (lane 0 of the first batch) */
    const WinMasks M_{win, tbm, nlm, trk, wbytes, false};
    const Reader rd_{win, a.S.p, win0, wbytes, a.S.len};
    const bool ctx = skip && nown > 0 && (win0 + starts[j0]) != 0;
    const uint32_t first = ctx ? j0 - 1u : j0;
    const uint32_t nparse = nown + (ctx ? 1u : 0u);
    uint32_t count = 0;                               /* records this span yields */
    /* results are kept for two batches (64 lines, i.e. lines of 160 bytes or more); shorter lines take the exact kernel */
    constexpr int MAXB = 2;
    uint4 Rrec[MAXB];
    uint32_t Rs[MAXB], Rmeta[MAXB];
    uint32_t rank[MAXB];
    if (nparse > 32u * MAXB) bad = true;
    uint32_t prev_qlen = 0, prev_h1 = 0, prev_h2 = 0, prev_qs = 0;
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
        rank[b] = NOT_YIELDED;
        if (bad || (uint32_t)(32 * b) >= nparse) continue;
        const uint32_t k = (uint32_t)(32 * b + lane);
        const bool mine = k < nparse;
        bool ok = true;
        LineRec L;
        L.qlen = 0; L.h1 = 0; L.h2 = 0; L.qs = 0; L.flags = 0; L.as = SCORE_ABSENT; L.xs = SCORE_ABSENT; L.s = 0; L.outlen = 0; L.rawbytes = 0;
        if (mine) {
            const int s = (int)starts[first + k], e = (int)starts[first + k + 1] - 1;
            FastCtx fc;
            ok = fast_head(M_, s, e, L, fc);
            if (ok && !(ctx && k == 0)) fast_tail(M_, s, e, a.score_src, fc, L);
        }
        if (__any_sync(0xffffffffu, !ok)) { bad = true; continue; }
        /* run heads (xm.py:110-114): a line is yielded when its QNAME differs from the line before */
        uint32_t pq = __shfl_up_sync(0xffffffffu, L.qlen, 1), p1 = __shfl_up_sync(0xffffffffu, L.h1, 1),
                 p2 = __shfl_up_sync(0xffffffffu, L.h2, 1), ps = __shfl_up_sync(0xffffffffu, L.qs, 1);
        if (lane == 0) { pq = prev_qlen; p1 = prev_h1; p2 = prev_h2; ps = prev_qs; }
        prev_qlen = __shfl_sync(0xffffffffu, L.qlen, 31); prev_h1 = __shfl_sync(0xffffffffu, L.h1, 31);
        prev_h2 = __shfl_sync(0xffffffffu, L.h2, 31); prev_qs = __shfl_sync(0xffffffffu, L.qs, 31);
        bool yield = mine && !(ctx && k == 0);
        if (yield && skip && (k > 0)) {
            if (pq == L.qlen && p1 == L.h1 && p2 == L.h2 && names_equal(rd_, win0 + ps, pq, win0 + L.qs, L.qlen)) yield = false;
        }
        const uint32_t ym = __ballot_sync(0xffffffffu, yield);
        if (yield) rank[b] = count + (uint32_t)__popc(ym & ((1u << lane) - 1u));
        count += (uint32_t)__popc(ym);
        Rrec[b] = make_uint4((uint32_t)L.as, (uint32_t)L.xs, L.h1, L.h2);
        Rs[b] = L.s;
        Rmeta[b] = (L.outlen & META_LEN_MASK) | ((L.flags & 0x3fu) << META_LEN_BITS);
    }
    if (bad) { count = 0; if (lane == 0) a.g->pad = 1u; }      /* Globals::pad doubles as the fallback flag */

    /* ---- the tile's record base: look-back chain 1 over CTAs ------------------------------------------ */
    if (lane == 0) s_cnt[warp] = count;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < C::WARPS; ++w) { const uint32_t c = s_cnt[w]; if (w < warp) wbase += c; total += c; }
    if (warp == 0) {
        if (lane == 0) dev_publish1(a.chain1, tile, total, false);
        dev_resolve1(a.chain1, tile, total, false, s_base);
    }
    __syncthreads();
    const unsigned long long base = s_base[0] + wbase;

    /* ---- the compact rows ------------------------------------------------------------------------------ */
    if (!bad) {
#pragma unroll
        for (int b = 0; b < MAXB; ++b) {
            if (rank[b] == NOT_YIELDED) continue;
            const unsigned long long gi = base + rank[b];
            if (gi < a.sc_cap) {
                a.sc.start[gi] = win0 + Rs[b];
                a.sc.rec[gi] = Rrec[b];
                a.sc.meta[gi] = Rmeta[b];
            }
        }
    }
    if (threadIdx.x == 0 && tile + 1 == a.ntiles) {
        const unsigned long long n = s_base[0] + total;
        a.g->n_stream[a.stream_id] = n;
        a.g->end_off[a.stream_id] = a.S.len;
        if (a.sc.start && n <= a.sc_cap) a.sc.start[n] = a.S.len;
    }
}

template <class C>
static cudaError_t launch_scan2_t(ScanArgs a, cudaStream_t st)
{
    const uint64_t nt = (a.S.len + C::TILE - 1) / C::TILE;
    a.ntiles = (uint32_t)nt;
    k_scan2<C><<<(unsigned)nt, C::WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_scan2(const ScanArgs &a, cudaStream_t st) { return launch_scan2_t<Scan2Big>(a, st); }
uint32_t scan2_tile_bytes() { return Scan2Big::TILE; }

}  // namespace xm
