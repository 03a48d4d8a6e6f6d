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
#include <type_traits>

namespace xm {

template <int WARPS_, int SPAN_, int BACK_, int EXT_>
struct Scan2Cfg {
    static constexpr int WARPS = WARPS_, SPAN = SPAN_, BACK = BACK_, EXT = EXT_;
    static constexpr int TILE = WARPS * SPAN;             /* bytes of the stream one CTA owns */
    static constexpr int WINB = BACK + SPAN + EXT;          /* bytes a warp may look at */
    static constexpr int NWW = WINB / 32 + 2;               /* mask words per warp (an even count and one spare) */
    static constexpr int LQ = 160;                          /* line starts a warp can list */
    static_assert(SPAN % 1024 == 0 && BACK % 1024 == 0 && EXT % 1024 == 0 && WINB < 65535, "geometry");
};
#ifndef XM_SCAN2_WARPS
#define XM_SCAN2_WARPS 8
#endif
#ifndef XM_SCAN2_SPAN
#define XM_SCAN2_SPAN 12288
#endif
#ifndef XM_SCAN2_OCC
#define XM_SCAN2_OCC 4
#endif
#ifndef XM_SCAN2_SPAN_S
#define XM_SCAN2_SPAN_S 11264
#endif
/* k_classify2 needs 68 registers: seven warps per CTA let four CTAs (28 warps) share a SM's register file; told to
 * fit 64 the compiler produces markedly slower code (profiles/r01_geometry_sweep.md) */
#ifndef XM_CLS2_WARPS
#define XM_CLS2_WARPS 7
#endif
using Scan2Big = Scan2Cfg<XM_CLS2_WARPS, XM_SCAN2_SPAN, 1024, 2048>;        /* k_classify2: ~28 primary lines of 440 bytes per span */
using Scan2Sec = Scan2Cfg<XM_SCAN2_WARPS, XM_SCAN2_SPAN_S, 1024, 2048>;      /* k_scan2: ~30 secondary lines of 378 bytes (one parse batch) */
/* short reads: a span holds at most 64 lines (two parse batches), so lines of less than ~190 bytes overflow the spans above and
 * send the whole call to the exact kernels.  Streams whose lines are short on average take spans of half the size: 64 lines of
 * 80-96 bytes and more still fit. */
using Scan2BigShort = Scan2Cfg<XM_CLS2_WARPS, 6144, 1024, 2048>;
using Scan2SecShort = Scan2Cfg<XM_SCAN2_WARPS, 5120, 1024, 2048>;

/* what a warp knows about its span once the masks are built and the line starts are listed */
struct SpanInfo {
    uint64_t win0, wend;       /* the window [win0, wend) in the stream */
    uint32_t hoff, wbytes;     /* the span's first byte inside the window; bytes in the window */
    uint32_t j0, nown;         /* index (in starts[]) of the first owned line; owned lines */
    bool live, bad;            /* the span lies inside the stream; it needs the exact kernel */
};

/* masks (tbm, nlm, trk: this warp's private shared memory), line starts and the owned range of one span.
 * need_prev: the line before the first owned one will be looked at (run heads, pair units).
 * (Round 2 tried a three-pass version -- bare mask steps, 512-byte tail steps, one index pass per span -- with
 * 15 % fewer instructions: 10 % slower in k_scan2, 1.5 % faster in k_classify2; these kernels are paced by the
 * dependent latencies a warp strings together, not by its instruction count.  profiles/r02_kernel_experiments.md) */
#ifdef XM_ASCII_MASKS
__device__ __forceinline__ void masks16_ascii(const uint4 v, uint32_t &W, uint32_t &T, uint32_t &acc)
{
    auto wf = [](uint32_t w) { return ~(w + 0x5f5f5f5fu) & 0x80808080u; };               /* byte + 0x5f: bit 7 set iff byte >= 0x21 */
    acc |= v.x | v.y;
    acc |= v.z | v.w;
    W = pack16(wf(v.x), wf(v.y), wf(v.z), wf(v.w));
    T = pack16(tab_mask_loose(v.x), tab_mask_loose(v.y), tab_mask_loose(v.z), tab_mask_loose(v.w));
}
#endif
template <class C, bool SPLIT>
__device__ __forceinline__ SpanInfo span_front(const StreamBuf &B, uint64_t span_lo, bool need_prev, uint32_t *tbm, uint32_t *nlm, uint16_t *trk, uint16_t *starts)
{
    const int lane = threadIdx.x & 31;
    SpanInfo si;
    si.live = span_lo < B.len;
    si.bad = false;
    /* window: [win0, win0 + wbytes); the span's first byte sits at hoff.  The bytes before it are looked at for the
     * line that precedes the span (its last newline decides whether the span opens a line) */
    si.win0 = si.live ? (span_lo >= (uint64_t)C::BACK ? span_lo - C::BACK : 0) : 0;
    si.hoff = (uint32_t)(span_lo - si.win0);
    si.wend = span_lo + C::SPAN + C::EXT;
    if (si.wend > B.len) si.wend = B.len;
    si.wbytes = si.live ? (uint32_t)(si.wend - si.win0) : 0u;
    si.j0 = 0; si.nown = 0;
    const uint64_t win0 = si.win0, wend = si.wend;
    const uint32_t hoff = si.hoff, wbytes = si.wbytes;
    const bool live = si.live;
    const bool skip = need_prev;
    bool bad = false;
    const uint32_t own_hi = hoff + C::SPAN < wbytes ? hoff + (uint32_t)C::SPAN : wbytes;     /* owned starts lie in [hoff, own_hi) */
    const uint8_t *win = B.p + win0;          /* read through L1/L2: no staged copy, so a SM holds 32 of these warps */

    /* ---- masks and line starts ------------------------------------------------------------------ */
    uint32_t nst = 0;                                 /* line starts listed so far (uniform) */
    uint32_t tab_run = 0;
    uint32_t j0 = 0, j1 = 0;
    /* the first word looked at: the one before the span (its last byte decides whether the span opens a line); when
     * the line before the span matters, SHORT_BACK bytes before it -- and the whole window in a second attempt if
     * that line turns out to start earlier (word 0 straight away at the head of the stream) */
    constexpr uint32_t SHORT_BACK = 640;
    uint32_t w_first = skip ? ((win0 != 0 && hoff > SHORT_BACK) ? ((hoff - SHORT_BACK) >> 5) & ~1u : 0u)
                            : ((hoff >= 32u ? hoff / 32u - 1u : 0u) & ~1u);
#pragma unroll 1
    for (;;) {
    nst = 0; tab_run = 0; j0 = 0; j1 = 0; bad = false;
    bool adj = false;
#ifdef XM_ASCII_MASKS
    uint32_t hi_acc = 0;
#endif
    if (live) {
        if (win0 == 0) { if (lane == 0) starts[0] = 0; nst = 1; }      /* the stream's first byte opens a line */
        /* two consecutive 32-byte words per lane and step (2 KiB per warp), so the scan, the shuffles and the loop
         * control below are paid once per 64 bytes */
        const uint32_t nwords = (wbytes + 31u) >> 5;
        uint32_t carryW = 0;                          /* was the byte before this lane's first word a separator? (lane 0's view) */
        bool done = false;
        const uint32_t lim16 = (wbytes + 15u) & ~15u;      /* the buffer is readable up to the next multiple of 16 */
        const uint4 filler = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
        const uint32_t lastp = own_hi - 1u;
        /* one step: 64 mask words, two per lane.  INTERIOR steps lie wholly inside the window and before the span's
         * last byte: no bounds selects, no partial words, and the last owned line cannot close there.  SPLIT compiles
         * the step twice; it pays in k_scan2 (+4 %) and costs in k_classify2 (-1.6 %, a larger kernel already). */
        auto step = [&](auto interior_c, const uint32_t wb) {
            constexpr bool INTERIOR = decltype(interior_c)::value;
            const uint32_t w = wb + 2u * (uint32_t)lane;           /* this lane's words: w, w + 1 */
#ifndef XM_NO_MASK_PF
            if (w + 64u < nwords) asm volatile("prefetch.global.L1 [%0];" ::"l"(win + (size_t)(w + 64u) * 32u));      /* the next step's bytes */
#endif
            uint32_t W0 = 0, T0 = 0, W1 = 0, T1 = 0;
            if (INTERIOR || w < nwords) {
                const uint32_t off = w * 32u;
                const uint4 v0 = ld_src16(win + off, false);
                const uint4 v1 = (INTERIOR || off + 16u < lim16) ? ld_src16(win + off + 16u, false) : filler;
                const uint4 v2 = (INTERIOR || off + 32u < lim16) ? ld_src16(win + off + 32u, false) : filler;
                const uint4 v3 = (INTERIOR || off + 48u < lim16) ? ld_src16(win + off + 48u, false) : filler;
                uint32_t Wa, Ta, Wb, Tb;
#ifdef XM_ASCII_MASKS
                /* W by one add and one logic operation per word: exact while every byte is below 0x80; the words are ORed
                 * into hi_acc and a span that saw a byte >= 0x80 gives up before its masks are used */
                masks16_ascii(v0, Wa, Ta, hi_acc);
                masks16_ascii(v1, Wb, Tb, hi_acc);
                W0 = Wa | (Wb << 16); T0 = Ta | (Tb << 16);
                masks16_ascii(v2, Wa, Ta, hi_acc);
                masks16_ascii(v3, Wb, Tb, hi_acc);
#else
                masks16_span(v0, Wa, Ta);
                masks16_span(v1, Wb, Tb);
                W0 = Wa | (Wb << 16); T0 = Ta | (Tb << 16);
                masks16_span(v2, Wa, Ta);
                masks16_span(v3, Wb, Tb);
#endif
                W1 = Wa | (Wb << 16); T1 = Ta | (Tb << 16);
                if (!INTERIOR && off + 64u > wbytes) {         /* the window's last words: bytes past its end do not count */
                    const uint32_t valid = wbytes - off;       /* 1..63 */
                    const uint32_t k0 = valid >= 32u ? 0xffffffffu : (1u << valid) - 1u;
                    const uint32_t k1 = valid > 32u ? (1u << (valid - 32u)) - 1u : 0u;
                    W0 &= k0; T0 &= k0; W1 &= k1; T1 &= k1;
                }
            }
            const uint32_t N0 = W0 & ~T0, N1 = W1 & ~T1;
            uint32_t prevW = __shfl_up_sync(0xffffffffu, W1 >> 31, 1);
            if (lane == 0) prevW = carryW;
            carryW = __shfl_sync(0xffffffffu, W1 >> 31, 31);
            if ((W0 & ((W0 << 1) | prevW)) | (W1 & ((W1 << 1) | (W0 >> 31)))) adj = true;
            /* tabs and line ends before each lane's words: one packed warp scan */
            const uint32_t x0 = (uint32_t)__popc(T0) | ((uint32_t)__popc(N0) << 16);
            const uint32_t x = x0 + ((uint32_t)__popc(T1) | ((uint32_t)__popc(N1) << 16));
            uint32_t inc = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
            const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
            const uint32_t exc = inc - x;
            if (INTERIOR || w + 1u < (uint32_t)C::NWW) {
                tbm[w] = T0; nlm[w] = N0; trk[w] = (uint16_t)(tab_run + (exc & 0xffffu));
                tbm[w + 1] = T1; nlm[w + 1] = N1; trk[w + 1] = (uint16_t)(tab_run + ((exc + x0) & 0xffffu));
            }
            uint32_t idx = nst + (exc >> 16);
            for (uint32_t m = N0; m; m &= m - 1) {
                if (idx < (uint32_t)C::LQ) starts[idx] = (uint16_t)(w * 32u + (uint32_t)__ffs((int)m));     /* the byte after the terminator */
                ++idx;
            }
            for (uint32_t m = N1; m; m &= m - 1) {
                if (idx < (uint32_t)C::LQ) starts[idx] = (uint16_t)(w * 32u + 32u + (uint32_t)__ffs((int)m));
                ++idx;
            }
            tab_run += tot & 0xffffu;
            nst += tot >> 16;
            if (!INTERIOR) {
                /* the last owned line is closed once a terminator at or beyond the span's last byte has been seen */
                const bool closed = (N1 && (w * 32u + 63u - (uint32_t)__clz((int)N1)) >= lastp) ||
                                    (N0 && (w * 32u + 31u - (uint32_t)__clz((int)N0)) >= lastp);
                done = __any_sync(0xffffffffu, closed);
            }
        };
        for (uint32_t wb = w_first; wb < nwords && !done; wb += 64u) {
            const uint32_t step_end = (wb + 64u) * 32u;            /* the byte after this step's last */
            if (SPLIT && step_end <= wbytes && step_end <= lastp) step(std::true_type{}, wb);
            else step(std::false_type{}, wb);
        }
#ifdef XM_ASCII_MASKS
        adj = __any_sync(0xffffffffu, adj || (hi_acc & 0x80808080u));
#else
        adj = __any_sync(0xffffffffu, adj);
#endif
        if (nst > (uint32_t)C::LQ || adj) bad = true;
    }
    __syncwarp();

    /* ---- the lines this span owns: starts in [hoff, own_hi) -------------------------------------------- */
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
        if (wend == B.len && !bad) {
            const uint32_t last = nst ? (uint32_t)starts[nst - 1] : 0xffffffffu;
            if (last != wbytes && own_hi == wbytes) bad = true;       /* unterminated last line */
        }
        if (skip && j1 > j0 && j0 == 0 && win0 + starts[0] != 0) {
            if (w_first > 0 && !bad) { w_first = 0; __syncwarp(); continue; }     /* look at the whole window */
            bad = true;                                                           /* the line before the span is not in the window */
        }
    }
    break;
    }
    si.bad = bad;
    si.j0 = j0;
    si.nown = (live && !bad) ? j1 - j0 : 0u;
    return si;
}


/* WANT_SAME / COUNT_ONLY: ScanArgs::want_same / count_only as compile-time switches (the plain scan keeps its code) */
template <class C, bool WANT_SAME, bool COUNT_ONLY>
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
    const bool skip = a.skip != 0;
    const bool need_prev = skip || WANT_SAME;             /* the line before the span's first is looked at (run heads, META_SAME) */
    uint32_t *tbm = s_tbm[warp], *nlm = s_nlm[warp];
    uint16_t *trk = s_trk[warp], *starts = s_start[warp];
    const SpanInfo si = span_front<C, true>(a.S, span_lo, need_prev, tbm, nlm, trk, starts);
    bool bad = si.bad;                                /* this span needs the exact kernel */
    const uint64_t win0 = si.win0;
    const uint32_t wbytes = si.wbytes, j0 = si.j0, nown = si.nown;
    const uint8_t *win = a.S.p + win0;

    /* ---- parse: one line per lane ---------------------------------------------------------------- */
    /* skipping walks also parse the head of the line before the first owned one (lane 0 of the first batch) */
    const WinMasks M_{win, tbm, nlm, trk, wbytes, false};
    const Reader rd_{win, a.S.p, win0, wbytes, a.S.len};
    const bool ctx = need_prev && nown > 0 && (win0 + starts[j0]) != 0;
    const uint32_t first = ctx ? j0 - 1u : j0;
    const uint32_t nparse = nown + (ctx ? 1u : 0u);
    uint32_t count = 0;                               /* records this span yields */
    uint32_t wbase = 0, total = 0;
    if (!skip) {
        /* every owned line is a record: the tile's count goes out before a single line is parsed */
        if (lane == 0) s_cnt[warp] = nown;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < C::WARPS; ++w) { const uint32_t c = s_cnt[w]; if (w < warp) wbase += c; total += c; }
        if (threadIdx.x == 0) dev_publish1(a.chain1, tile, total, false);
    }
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
            ok = fast_head(M_, s, e, L, fc);                  /* the aux tokens come after the record count is out */
        }
        if (__any_sync(0xffffffffu, !ok)) { bad = true; continue; }
        /* run heads (xm.py:110-114): a line is yielded when its QNAME differs from the line before */
        uint32_t pq = __shfl_up_sync(0xffffffffu, L.qlen, 1), p1 = __shfl_up_sync(0xffffffffu, L.h1, 1),
                 p2 = __shfl_up_sync(0xffffffffu, L.h2, 1), ps = __shfl_up_sync(0xffffffffu, L.qs, 1);
        if (lane == 0) { pq = prev_qlen; p1 = prev_h1; p2 = prev_h2; ps = prev_qs; }
        prev_qlen = __shfl_sync(0xffffffffu, L.qlen, 31); prev_h1 = __shfl_sync(0xffffffffu, L.h1, 31);
        prev_h2 = __shfl_sync(0xffffffffu, L.h2, 31); prev_qs = __shfl_sync(0xffffffffu, L.qs, 31);
        bool yield = mine && !(ctx && k == 0);
        bool same = false;
        if (yield && need_prev && (k > 0)) {
            if (pq == L.qlen && p1 == L.h1 && p2 == L.h2 && names_equal(rd_, win0 + ps, pq, win0 + L.qs, L.qlen)) same = true;
        }
        if (skip && same) yield = false;
        const uint32_t ym = __ballot_sync(0xffffffffu, yield);
        if (yield) rank[b] = count + (uint32_t)__popc(ym & ((1u << lane) - 1u));
        count += (uint32_t)__popc(ym);
        Rrec[b] = make_uint4((uint32_t)L.as, (uint32_t)L.xs, L.h1, L.h2);
        Rs[b] = L.s;
        Rmeta[b] = (L.outlen & META_LEN_MASK) | ((L.flags & 0x3fu) << META_LEN_BITS) | ((WANT_SAME && same) ? META_SAME : 0u);
    }
    if (bad) { count = 0; if (lane == 0) a.g->pad = 1u; }      /* Globals::pad doubles as the fallback flag */

    /* ---- the tile's record base: look-back chain 1 over CTAs ------------------------------------------ */
    if (skip) {
        if (lane == 0) s_cnt[warp] = count;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < C::WARPS; ++w) { const uint32_t c = s_cnt[w]; if (w < warp) wbase += c; total += c; }
        /* the tile's count goes out now; its base is looked up after the aux tokens are parsed, when the tiles
         * before this one have long published theirs */
        if (threadIdx.x == 0) dev_publish1(a.chain1, tile, total, false);
    }
    if (!bad && !COUNT_ONLY) {
#pragma unroll
        for (int b = 0; b < MAXB; ++b) {
            const uint32_t k = (uint32_t)(32 * b + lane);
            if (k >= nparse || (ctx && k == 0)) continue;
            const int s = (int)starts[first + k], e = (int)starts[first + k + 1] - 1;
            FastCtx fc;
            fc.r0 = tabs_before(M_, s);
            fc.ntab = tabs_before(M_, e) - fc.r0;
            LineRec X;
            X.flags = 0; X.as = SCORE_ABSENT; X.xs = SCORE_ABSENT;
            fast_tail(M_, s, e, a.score_src, fc, X);
            Rrec[b].x = (uint32_t)X.as; Rrec[b].y = (uint32_t)X.xs;
            Rmeta[b] |= (X.flags & 0x3fu) << META_LEN_BITS;
        }
    }
    if (warp == 0) dev_resolve1(a.chain1, tile, total, false, s_base);
    __syncthreads();
    const unsigned long long base = s_base[0] + wbase;

    /* ---- the compact rows ------------------------------------------------------------------------------ */
    if (!bad) {
#pragma unroll
        for (int b = 0; b < MAXB; ++b) {
            if (rank[b] == NOT_YIELDED) continue;
            const unsigned long long gi = base + rank[b];
            if (gi < a.sc_cap) {
                a.sc.start[gi] = a.start_bias + win0 + Rs[b];
                if (!COUNT_ONLY) { a.sc.rec[gi] = Rrec[b]; a.sc.meta[gi] = Rmeta[b]; }
            }
        }
    }
    if (threadIdx.x == 0 && tile + 1 == a.ntiles) {
        const unsigned long long n = s_base[0] + total;
        a.g->n_stream[a.stream_id] = n;
        a.g->end_off[a.stream_id] = a.S.len;
        if (a.sc.start && n <= a.sc_cap) a.sc.start[n] = a.start_bias + a.S.len;
    }
}

/* =========================================================================
 * k_classify2: the primary-stream walk in the same barrier-free style.
 * Each warp parses the lines of its span, joins them with the secondary
 * stream's compact rows by record index, decides the categories, sizes its
 * part of the six bins with warp scans and copies its lines global -> global
 * (the span was read from HBM microseconds earlier and is still in L2).  The
 * warps of a CTA meet four times per tile: record counts, record base (chain
 * 1), byte totals per bin, byte bases (chain 2).  Clean inputs without errors
 * only; everything else raises the fallback flag and k_classify runs.
 * ========================================================================= */
#ifndef XM_CLS2_OCC
#define XM_CLS2_OCC 4
#endif
/* the len + 1 bytes at p and at s are equal (a QNAME and the separator behind it); both readable to the next
 * multiple of 16 past their buffers' ends */
__device__ __forceinline__ bool qnames_equal_bytes(const uint8_t *p, const uint8_t *s, uint32_t len)
{
    const uint32_t *pw = (const uint32_t *)((uintptr_t)p & ~(uintptr_t)3), *sw = (const uint32_t *)((uintptr_t)s & ~(uintptr_t)3);
    const uint32_t shp = (uint32_t)((uintptr_t)p & 3u) * 8u, shs = (uint32_t)((uintptr_t)s & 3u) * 8u;
    const uint32_t n = len + 1u;
    uint32_t plo = __ldg(pw), slo = __ldg(sw);
#pragma unroll 1
    for (uint32_t done = 0, j = 1; done < n; done += 4u, ++j) {
        const uint32_t phi = __ldg(pw + j), shi = __ldg(sw + j);
        uint32_t x = __funnelshift_r(plo, phi, shp) ^ __funnelshift_r(slo, shi, shs);
        if (n - done < 4u) x &= (1u << (8u * (n - done))) - 1u;
        if (x) return false;
        plo = phi; slo = shi;
    }
    return true;
}
constexpr int CLS2_LINES = 64;            /* lines (with the context line) a span may hold: two batches */

template <class C>
struct Cls2Smem {
    uint32_t tbm[C::WARPS][C::NWW], nlm[C::WARPS][C::NWW];
    uint16_t trk[C::WARPS][C::NWW];
    uint16_t starts[C::WARPS][C::LQ];
    /* per parsed line */
    int32_t as[C::WARPS][CLS2_LINES], xs[C::WARPS][CLS2_LINES];
    uint32_t h1[C::WARPS][CLS2_LINES], h2[C::WARPS][CLS2_LINES];
    uint32_t so[C::WARPS][CLS2_LINES];          /* start (16) | emitted length (16) */
    uint32_t q[C::WARPS][CLS2_LINES];           /* QNAME start (16) | QNAME length (16) */
    uint16_t rank[C::WARPS][CLS2_LINES];        /* rank among the span's yielded records, 0xffff none */
    uint32_t cnt[C::WARPS];
    uint32_t wtot[C::WARPS][8];                 /* bytes per bin (6), raw bytes (slot 6) of each warp */
    unsigned long long tot[8], keep[8];
    unsigned long long wbase[C::WARPS][6];       /* where each warp's bytes start in the six bins */
    unsigned long long base1[2];
};

template <class C>
__global__ void __launch_bounds__(C::WARPS * 32, XM_CLS2_OCC) k_classify2(const ClassifyArgs a)
{
    extern __shared__ uint4 xm_smem_c2[];
    Cls2Smem<C> &S = *reinterpret_cast<Cls2Smem<C> *>(xm_smem_c2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x;
    const uint64_t span_lo = ((uint64_t)tile * C::WARPS + (uint64_t)warp) * (uint64_t)C::SPAN;
    const bool skip = a.skip != 0, paired = a.mode != MODE_SE;
    const bool need_prev = skip || paired;
    uint32_t *tbm = S.tbm[warp], *nlm = S.nlm[warp];
    uint16_t *trk = S.trk[warp], *starts = S.starts[warp];
    /* copy items (primary part and secondary part of what line k emits) overlay the warp's tab mask, which is dead
     * once the lines are parsed: dst | src (16) + len (16) | dst | len, then bin and line count bytes */
    static_assert(C::NWW * 4 >= CLS2_LINES * 18, "the copy items overlay the tab mask");
    uint32_t *it_dst = tbm, *it_sl = tbm + CLS2_LINES, *is_dst = tbm + 2 * CLS2_LINES, *is_len = tbm + 3 * CLS2_LINES;
    uint8_t *it_bin = (uint8_t *)(tbm + 4 * CLS2_LINES), *is_nl = it_bin + CLS2_LINES;
    const SpanInfo si = span_front<C, false>(a.P, span_lo, need_prev, tbm, nlm, trk, starts);
    bool bad = si.bad;
    const uint64_t win0 = si.win0;
    const uint32_t wbytes = si.wbytes, j0 = si.j0, nown = si.nown;
    const uint8_t *win = a.P.p + win0;
    const WinMasks M_{win, tbm, nlm, trk, wbytes, false};
    const Reader rd_{win, a.P.p, win0, wbytes, a.P.len};

    /* ---- parse: one line per lane; line 0 is the line before the span when the walk looks at predecessors ---- */
    const bool ctx = need_prev && nown > 0 && (win0 + starts[j0]) != 0;
    const uint32_t first = ctx ? j0 - 1u : j0;
    const uint32_t nparse = nown + (ctx ? 1u : 0u);
    if (nparse > (uint32_t)CLS2_LINES) bad = true;
    uint32_t count = 0;
    uint32_t wbase = 0, total = 0;
    if (!skip) {
        /* every owned line is a record: the tile's count goes out before a single line is parsed */
        if (lane == 0) S.cnt[warp] = nown;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < C::WARPS; ++w) { const uint32_t c = S.cnt[w]; if (w < warp) wbase += c; total += c; }
        if (threadIdx.x == 0) dev_publish1(a.chain1, tile, total, false);
    }
    {
        uint32_t prev_qlen = 0, prev_h1 = 0, prev_h2 = 0, prev_qs = 0;
        for (uint32_t kb = 0; kb < nparse && !bad; kb += 32u) {
            const uint32_t k = kb + (uint32_t)lane;
            const bool mine = k < nparse;
            bool ok = true;
            LineRec L;
            L.qlen = 0; L.h1 = 0; L.h2 = 0; L.qs = 0; L.flags = 0; L.as = SCORE_ABSENT; L.xs = SCORE_ABSENT; L.s = 0; L.outlen = 0; L.rawbytes = 0;
            if (mine) {
                const int s = (int)starts[first + k], e = (int)starts[first + k + 1] - 1;
                FastCtx fc;
                ok = fast_head(M_, s, e, L, fc);              /* the aux tokens come after the record count is out */
            }
            if (__any_sync(0xffffffffu, !ok)) { bad = true; break; }
            uint32_t pq = __shfl_up_sync(0xffffffffu, L.qlen, 1), p1 = __shfl_up_sync(0xffffffffu, L.h1, 1),
                     p2 = __shfl_up_sync(0xffffffffu, L.h2, 1), ps = __shfl_up_sync(0xffffffffu, L.qs, 1);
            if (lane == 0) { pq = prev_qlen; p1 = prev_h1; p2 = prev_h2; ps = prev_qs; }
            prev_qlen = __shfl_sync(0xffffffffu, L.qlen, 31); prev_h1 = __shfl_sync(0xffffffffu, L.h1, 31);
            prev_h2 = __shfl_sync(0xffffffffu, L.h2, 31); prev_qs = __shfl_sync(0xffffffffu, L.qs, 31);
            bool yield = mine && !(ctx && k == 0);
            /* does this line carry the QNAME of the line before it?  (run heads xm.py:110-114, pair units xm.py:402) */
            bool same = false;
            if (mine && k > 0 && pq == L.qlen && p1 == L.h1 && p2 == L.h2) same = names_equal(rd_, win0 + ps, pq, win0 + L.qs, L.qlen);
            if (skip && same) yield = false;
            const uint32_t ym = __ballot_sync(0xffffffffu, yield);
            if (mine) {
                S.as[warp][k] = L.as; S.xs[warp][k] = L.xs; S.h1[warp][k] = L.h1; S.h2[warp][k] = L.h2;
                S.so[warp][k] = L.s | (L.outlen << 16);
                S.q[warp][k] = (same ? 1u : 0u) | (L.qlen << 1);   /* all later phases need of the QNAME: repeats the line before, length */
                S.rank[warp][k] = yield ? (uint16_t)(count + (uint32_t)__popc(ym & ((1u << lane) - 1u))) : (uint16_t)0xffff;
            }
            count += (uint32_t)__popc(ym);
        }
    }
    if (bad) { count = 0; if (lane == 0) a.g->pad = 1u; }

    /* ---- record base: chain 1 ------------------------------------------------------------------- */
    if (skip) {
        if (lane == 0) S.cnt[warp] = count;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < C::WARPS; ++w) { const uint32_t c = S.cnt[w]; if (w < warp) wbase += c; total += c; }
        if (threadIdx.x == 0) dev_publish1(a.chain1, tile, total, false);
    }
    /* the aux tokens (scores), while the tiles before this one publish their counts */
    for (uint32_t kb = 0; kb < nparse && !bad; kb += 32u) {
        const uint32_t k = kb + (uint32_t)lane;
        bool ok = true;
        if (k < nparse && (paired || !(ctx && k == 0))) {
            const int s = (int)starts[first + k], e = (int)starts[first + k + 1] - 1;
            FastCtx fc;
            fc.r0 = tabs_before(M_, s);
            fc.ntab = tabs_before(M_, e) - fc.r0;
            LineRec X;
            X.flags = 0; X.as = SCORE_ABSENT; X.xs = SCORE_ABSENT;
            fast_tail(M_, s, e, a.score_src, fc, X);
            S.as[warp][k] = X.as; S.xs[warp][k] = X.xs;
            if (X.flags) ok = false;                          /* score errors are reported by the exact kernel */
        }
        if (__any_sync(0xffffffffu, !ok)) { bad = true; if (lane == 0) a.g->pad = 1u; }
    }
    if (warp == 0) dev_resolve1(a.chain1, tile, total, false, S.base1);
    __syncthreads();
    const unsigned long long base = S.base1[0] + wbase;        /* record index of this span's first yielded record */
    unsigned long long ncap = a.g->n_stream[1];
    if (a.limit < ncap) ncap = a.limit;
    if (a.sc_cap < ncap) ncap = a.sc_cap;      /* more secondary records than compact rows: nothing past them is read; the host grows the arrays and walks again */

    /* ---- join, decide, size ------------------------------------------------------------------------ */
    uint32_t wtot_l = 0;                     /* lane b < 7 carries the warp's running total of slot b (six bins, raw bytes) */
    uint32_t todo0 = 0, todo1 = 0;           /* per batch of 32 lines: which lines start a copy */
    {
        int pst_carry = 0;                       /* state, emitted lengths and validity of the line before lane 0's */
        uint32_t pout_carry = 0, pslen_carry = 0, pso_carry = 0;
        bool pvalid_carry = false;
        for (uint32_t kb = 0; kb < nparse && !bad; kb += 32u) {
            const uint32_t k = kb + (uint32_t)lane;
            const bool mine = k < nparse;
            const uint32_t rk = mine ? (uint32_t)S.rank[warp][k] : 0xffffu;
            /* record index: yielded lines by their rank; the context line is the record before the span's first */
            bool valid = false;
            unsigned long long gi = 0;
            if (mine) {
                if (rk != 0xffffu) { gi = base + rk; valid = gi < ncap; }
                else if (ctx && k == 0 && !skip && base > 0) { gi = base - 1; valid = gi < ncap; }
            }
            int st = UA;
            uint32_t slen = 0, so = 0;
            bool mismatch = false;
            if (valid) {
                const uint4 sr = a.sc.rec[gi];
                const uint32_t smeta = a.sc.meta[gi];
                if ((smeta & META_FLAGS) || sr.z != S.h1[warp][k] || sr.w != S.h2[warp][k]) mismatch = true;     /* dirty / failing secondary line, QNAME assert: exact kernel */
                /* equal 64-bit hashes are taken for equal names; XM_DEBUG_EXACT_NAMES compares the bytes,
                 * separator included, as the exact kernel always does: one scattered read of the secondary line per record */
                else if ((a.debug & DBG_EXACT_NAMES) && !qnames_equal_bytes(win + (S.so[warp][k] & 0xffffu), a.S.p + a.sc.start[gi], S.q[warp][k] >> 1)) mismatch = true;
                slen = smeta & META_LEN_MASK;
                st = mapping_state(S.as[warp][k], S.xs[warp][k], (int32_t)sr.x, (int32_t)sr.y, a.thr);
                so = S.so[warp][k];
                if (rk != 0xffffu && a.p_start && gi < a.p_start_cap) a.p_start[gi] = win0 + (so & 0xffffu);
            }
            if (__any_sync(0xffffffffu, mismatch)) { bad = true; break; }
            const uint32_t outlen = so >> 16;
            /* the line before, for pair units */
            int pst = __shfl_up_sync(0xffffffffu, st, 1);
            uint32_t pout = __shfl_up_sync(0xffffffffu, outlen, 1), pslen = __shfl_up_sync(0xffffffffu, slen, 1), pso = __shfl_up_sync(0xffffffffu, so, 1);
            bool pvalid = __shfl_up_sync(0xffffffffu, (int)valid, 1) != 0;
            if (lane == 0) { pst = pst_carry; pout = pout_carry; pslen = pslen_carry; pso = pso_carry; pvalid = pvalid_carry; }
            pst_carry = __shfl_sync(0xffffffffu, st, 31); pout_carry = __shfl_sync(0xffffffffu, outlen, 31);
            pslen_carry = __shfl_sync(0xffffffffu, slen, 31); pso_carry = __shfl_sync(0xffffffffu, so, 31);
            pvalid_carry = __shfl_sync(0xffffffffu, (int)valid, 31) != 0;

            uint32_t key = 36, bin = NO_BIN, plen = 0, sbytes = 0, src = 0, raw = 0;
            int nl = 1;
            if (valid && rk != 0xffffu) {
                raw = outlen;
                if (!paired) {
                    if (!(a.halo && gi == 0)) {
                        key = (uint32_t)st; bin = (uint32_t)st;
                        const bool pside = st == PS || st == PM || st == UA || st == UR, sside = st == SS || st == SM || st == UR;
                        plen = pside ? outlen : 0u; sbytes = sside ? slen : 0u; src = so & 0xffffu;
                    }
                } else if (!skip && (S.q[warp][k] & 1u) && pvalid && gi > 0) {
                    key = (uint32_t)(pst * 6 + st);
                    bin = (uint32_t)(a.mode == MODE_PE_CONSERVATIVE ? pair_bin_conservative(pst, st) : pair_bin_liberal(pst, st));
                    const bool pside = bin == PS || bin == PM || bin == UA || bin == UR, sside = bin == SS || bin == SM || bin == UR;
                    plen = pside ? pout + outlen : 0u; sbytes = sside ? pslen + slen : 0u; src = pso & 0xffffu; nl = 2;
                }
                if (bin != NO_BIN && !((a.enabled >> bin) & 1u)) { plen = 0; sbytes = 0; }
            }
            {
                /* one global atomic per distinct category in the warp */
                const unsigned peers = __match_any_sync(0xffffffffu, key);
                if (key < 36u && lane == __ffs((int)peers) - 1) atomicAdd(&a.g->counts[key], (unsigned long long)__popc(peers));
            }
            /* offsets inside the warp's part of each bin */
            const uint32_t bytes = plen + sbytes;
            uint32_t off = 0;
#pragma unroll
            for (uint32_t b = 0; b < 6; ++b) {
                const uint32_t v = bin == b ? bytes : 0u;
                if (__any_sync(0xffffffffu, v != 0u)) {
                    uint32_t x = v;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
                    const uint32_t run = __shfl_sync(0xffffffffu, wtot_l, (int)b);
                    if (bin == b) off = run + x - v;
                    const uint32_t add = __shfl_sync(0xffffffffu, x, 31);
                    if (lane == (int)b) wtot_l += add;
                }
            }
            { const uint32_t add = __reduce_add_sync(0xffffffffu, raw); if (lane == 6) wtot_l += add; }
            /* copy items; primary parts that continue each other in the input and in the bin are merged into runs.
             * Lanes without a primary part (first mates of a pair, secondary-only bins) do not break a run. */
            const uint32_t pend_src = src + plen, pend_dst = off + plen;
            const uint32_t items = __ballot_sync(0xffffffffu, plen != 0u);
            const uint32_t below = items & ((1u << lane) - 1u);
            const int pl = below ? 31 - __clz((int)below) : 0;                          /* the item before this lane's */
            const uint32_t q_src = __shfl_sync(0xffffffffu, pend_src, pl), q_dst = __shfl_sync(0xffffffffu, pend_dst, pl),
                           q_bin = __shfl_sync(0xffffffffu, bin, pl);
            const bool head = plen && !(below && q_bin == bin && q_src == src && q_dst == off);
            const uint32_t heads = __ballot_sync(0xffffffffu, head);
            const uint32_t above = heads & (lane == 31 ? 0u : (0xffffffffu << (lane + 1)));
            const uint32_t upto = above ? ((1u << (__ffs((int)above) - 1)) - 1u) : 0xffffffffu;   /* lanes before the next run */
            const uint32_t mineq = items & upto;
            const int last = mineq ? 31 - __clz((int)mineq) : lane;                     /* last item of this lane's run */
            const uint32_t run_end = __shfl_sync(0xffffffffu, pend_src, last);
            if (mine) {
                it_dst[k] = off; it_bin[k] = (uint8_t)bin;
                it_sl[k] = head ? (src | ((run_end - src) << 16)) : 0u;
                is_dst[k] = off + plen; is_len[k] = sbytes; is_nl[k] = (uint8_t)nl;
            }
            /* the lines that start a copy (a run of primary parts, or a secondary part): the copy loop visits only these */
            const uint32_t t = __ballot_sync(0xffffffffu, mine && (head || sbytes != 0u));
            if (kb == 0) todo0 = t; else todo1 = t;
        }
    }
    if (bad) {
        if (lane == 0) a.g->pad = 1u;
        wtot_l = 0;
    }

    /* ---- byte bases: chain 2 ------------------------------------------------------------------------- */
    if (lane < 8) S.wtot[warp][lane] = lane < 7 ? wtot_l : 0u;
    __syncthreads();
    if (warp == 0) {
        if (lane < 8) {
            unsigned long long t = 0;
            for (int w = 0; w < C::WARPS; ++w) t += S.wtot[w][lane];
            S.tot[lane] = t; S.keep[lane] = t;
        }
        __syncwarp();
        dev_publish2(a.chain2, tile, S.tot);
        dev_resolve2(a.chain2, tile, S.tot, nullptr);
    }
    __syncthreads();

    /* ---- copy -------------------------------------------------------------------------------------------- */
    if (!bad) {
        if (lane < 6) {
            unsigned long long t = S.tot[lane];
            for (int w = 0; w < warp; ++w) t += S.wtot[w][lane];
            S.wbase[warp][lane] = t;
        }
        __syncwarp();
        static_assert(CLS2_LINES == 64, "two batches of copy starts");
#pragma unroll 1
        for (uint32_t m = todo0, kb = 0; kb < 64u; m = todo1, kb += 32u)
#pragma unroll 1
        for (; m; m &= m - 1) {
            const uint32_t k = kb + (uint32_t)__ffs((int)m) - 1u;
            const uint32_t sl = it_sl[k], slen = is_len[k];
            const uint32_t bin = it_bin[k];
            const unsigned long long bb = S.wbase[warp][bin < 6u ? bin : 0u];
            if (sl) {
                const uint32_t len = sl >> 16;
                const unsigned long long doff = bb + it_dst[k];
                if (doff + len <= a.out_cap[bin]) dev_copy_global(a.out[bin] + doff, win + (sl & 0xffffu), len);
            }
            if (slen) {
                const unsigned long long doff = bb + is_dst[k];
                const unsigned long long gi = base + S.rank[warp][k];
                const uint64_t ss = a.sc.start[is_nl[k] == 2 ? gi - 1 : gi];
                if (doff + slen <= a.out_cap[bin]) dev_copy_global(a.out[bin] + doff, a.S.p + ss, slen);
            }
        }
    }
    if (threadIdx.x == 0) {
        if (S.keep[6]) atomicAdd(&a.g->bytes_in[0], S.keep[6]);
        if (tile + 1 == a.ntiles) {
            a.g->n_stream[0] = S.base1[0] + total;
            a.g->end_off[0] = a.P.len;
            for (int b = 0; b < 6; ++b) a.g->out_len[b] = S.tot[b] + S.keep[b];
        }
    }
}

template <class C>
static cudaError_t launch_classify2_t(ClassifyArgs a, cudaStream_t st)
{
    const uint64_t nt = (a.P.len + C::TILE - 1) / C::TILE;
    a.ntiles = (uint32_t)nt;
    const size_t smem = sizeof(Cls2Smem<C>);
    cudaError_t e = cudaFuncSetAttribute(k_classify2<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_classify2<C><<<(unsigned)nt, C::WARPS * 32, smem, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_classify2(const ClassifyArgs &a, cudaStream_t st)
{
    return a.short_lines ? launch_classify2_t<Scan2BigShort>(a, st) : launch_classify2_t<Scan2Big>(a, st);
}

template <class C>
static cudaError_t launch_scan2_t(ScanArgs a, cudaStream_t st)
{
    const uint64_t nt = (a.S.len + C::TILE - 1) / C::TILE;
    a.ntiles = (uint32_t)nt;
    if (a.count_only) k_scan2<C, false, true><<<(unsigned)nt, C::WARPS * 32, 0, st>>>(a);
    else if (a.want_same) k_scan2<C, true, false><<<(unsigned)nt, C::WARPS * 32, 0, st>>>(a);
    else k_scan2<C, false, false><<<(unsigned)nt, C::WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_scan2(const ScanArgs &a, cudaStream_t st)
{
    return a.short_lines ? launch_scan2_t<Scan2SecShort>(a, st) : launch_scan2_t<Scan2Sec>(a, st);
}
uint32_t scan2_tile_bytes() { return Scan2Sec::TILE; }
/* the smallest tile any of the span kernels' geometries has: look-back arrays are sized for it */
uint32_t span_tile_bytes_min()
{
    uint32_t t = Scan2Sec::TILE;
    if (Scan2Big::TILE < t) t = Scan2Big::TILE;
    if (Scan2BigShort::TILE < t) t = Scan2BigShort::TILE;
    if (Scan2SecShort::TILE < t) t = Scan2SecShort::TILE;
    return t;
}

}  // namespace xm
