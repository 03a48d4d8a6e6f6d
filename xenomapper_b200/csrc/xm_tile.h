/*
 * xm_tile.h -- what one CTA does with one tile of a SAM stream.
 *
 *   scan_tile      secondary stream: stage window, byte-class masks, line
 *                  index, per-line parse, record ranks by decoupled
 *                  look-back, compact per-record arrays (SCompact).
 *   classify_tile  primary stream: the same front end, then joins each record
 *                  with its secondary-stream entry by record index, decides
 *                  the category (single reads or adjacent-QNAME pairs),
 *                  histograms it, scans the six bins' byte counts (second
 *                  look-back chain) and copies the emitted lines from the
 *                  staged window (or, for secondary-stream lines, from global
 *                  memory) to their final place in the six outputs.
 *
 * A line belongs to the tile that holds its first byte.  The window staged in
 * shared memory extends HALO bytes to both sides so that the last owned line
 * and the line before the first owned one are normally inside it; lines that
 * are not are handled through global memory by the exact byte-wise path.
 *
 * The code is written as barrier-separated phases so that one source builds
 * both the CUDA kernels (each phase runs once per thread) and the CPU
 * emulation used by the tests (each phase loops over the emulated threads).
 */
#pragma once
#include "xm_parse.h"

#if !defined(__CUDACC__)
#include <string.h>
#endif

namespace xm {

/* ---- shared-memory layout --------------------------------------------- */
template <class C>
struct TileMem {
    uint8_t *win;            /* WIN + 32 staged bytes */
    uint32_t *wsm, *nlm;     /* NW words each */
    uint32_t *stm;           /* TILE/32 words: bit set where an owned line starts */
    uint16_t *grp;           /* NG+1 exclusive line counts per 4-word group */
    uint4 *shl;              /* LCAP+1 per-line records shared with the next line's owner; [0] is the halo line */
    uint32_t *it_dst, *it_src, *it_meta;   /* ITEMS each; alias the masks (dead by then) */
    uint32_t *scr;           /* 96 words of scratch for block collectives and the halo line */
    unsigned long long *scr64;   /* 16 */
    uint32_t *hist;          /* 36 */
};

/* scratch slots */
enum { SCR_HALO_QLEN = 64, SCR_HALO_H1, SCR_HALO_H2, SCR_HALO_FLAGS, SCR_HALO_OUTLEN, SCR_HALO_AS, SCR_HALO_XS, SCR_HALO_VALID };
enum { S64_TOT = 0 /* 0..7 */, S64_HALO_START = 8, S64_HALO_QS = 9, S64_BASE = 10, S64_PSTOP = 11, S64_BLANK_OFF = 12, S64_RAW = 13 };

template <class C>
struct TileLayout {
    static constexpr size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
    static constexpr size_t masks_bytes = align16((size_t)C::NW * 4) * 2 + align16((size_t)C::TILE / 8);
    static constexpr size_t items_bytes = align16((size_t)C::ITEMS * 4) * 3;
    static constexpr size_t o_union = align16((size_t)C::WIN + 32);
    static constexpr size_t union_bytes = masks_bytes > items_bytes ? masks_bytes : items_bytes;
    static constexpr size_t o_nlm = o_union + align16((size_t)C::NW * 4);
    static constexpr size_t o_stm = o_union + align16((size_t)C::NW * 4) * 2;
    static constexpr size_t o_it_src = o_union + align16((size_t)C::ITEMS * 4);
    static constexpr size_t o_it_meta = o_union + align16((size_t)C::ITEMS * 4) * 2;
    static constexpr size_t o_grp = o_union + union_bytes;
    static constexpr size_t o_shl = o_grp + align16((size_t)(C::NG + 1) * 2);
    static constexpr size_t o_scr = o_shl + (size_t)(C::LCAP + 1) * 16;
    static constexpr size_t o_scr64 = o_scr + 96 * 4;
    static constexpr size_t o_hist = o_scr64 + 16 * 8;
    static constexpr size_t total = o_hist + align16(36 * 4);
};

template <class C>
XM_HD TileMem<C> carve(void *base)
{
    using Lo = TileLayout<C>;
    uint8_t *b = (uint8_t *)base;
    TileMem<C> m;
    m.win = b;
    m.wsm = (uint32_t *)(b + Lo::o_union);
    m.nlm = (uint32_t *)(b + Lo::o_nlm);
    m.stm = (uint32_t *)(b + Lo::o_stm);
    m.it_dst = (uint32_t *)(b + Lo::o_union);
    m.it_src = (uint32_t *)(b + Lo::o_it_src);
    m.it_meta = (uint32_t *)(b + Lo::o_it_meta);
    m.grp = (uint16_t *)(b + Lo::o_grp);
    m.shl = (uint4 *)(b + Lo::o_shl);
    m.scr = (uint32_t *)(b + Lo::o_scr);
    m.scr64 = (unsigned long long *)(b + Lo::o_scr64);
    m.hist = (uint32_t *)(b + Lo::o_hist);
    return m;
}

/* ---- per-thread state that lives across phases -------------------------- */
template <class C>
struct ThreadState {
    LineRec L[C::R];
    uint32_t rank[C::R];     /* position among the records this tile yields; ~0u = not yielded */
    uint32_t ymeta[C::R];    /* bin (3 bits, NO_BIN = nothing emitted) | Y_* bits */
    uint32_t ybytes[C::R];
    uint32_t yoff[C::R];
    uint32_t sin, sout;      /* block-collective operand / result */
};
enum : uint32_t { Y_ASSERT = 0x100, Y_PREV_DIRTY = 0x200 };
constexpr uint32_t NOT_YIELDED = 0xffffffffu;

template <class C>
struct TileCtx {
    TileMem<C> m;
    ThreadState<C> *emu;     /* CPU emulation only: the THREADS thread states */
};

/* copy item: it_meta = len (26 bits) | two_lines << 26 | bin << 27 | kind << 30 */
enum { IT_EMPTY = 0, IT_P_COPY = 1, IT_P_NORM = 2, IT_S = 3 };

/* word shared with the next line's owner: outlen (24) | state (3) << 24 | evalerr (3) << 27 | dirty << 30 | evalstream << 31 */
XM_HD uint32_t shl_pack(uint32_t outlen, int state, int evalerr, int evalstream, bool dirty)
{
    return (outlen & META_LEN_MASK) | ((uint32_t)state << 24) | ((uint32_t)evalerr << 27) | ((uint32_t)dirty << 30) | ((uint32_t)evalstream << 31);
}

/* ---- execution-model macros ---------------------------------------------- */
#if XM_DEVICE_PASS
#define XM_THREADS_BEGIN { const int tid = (int)threadIdx.x; ThreadState<C> &th = th_; (void)tid; (void)th;
#define XM_THREADS_END }
#define XM_BARRIER() __syncthreads()
#define XM_BLOCK_SCAN(total) do { th_.sout = dev_block_scan(th_.sin, T.m.scr, total); } while (0)
#define XM_BLOCK_MIN(result) do { result = dev_block_min(th_.sin, T.m.scr); } while (0)
#define XM_BLOCK_OR(result) do { result = dev_block_or(th_.sin, T.m.scr); } while (0)
#else
#define XM_THREADS_BEGIN for (int tid = 0; tid < C::THREADS; ++tid) { ThreadState<C> &th = T.emu[tid]; (void)th;
#define XM_THREADS_END }
#define XM_BARRIER() ((void)0)
#define XM_BLOCK_SCAN(total) do { uint32_t run_ = 0; for (int t_ = 0; t_ < C::THREADS; ++t_) { T.emu[t_].sout = run_; run_ += T.emu[t_].sin; } total = run_; } while (0)
#define XM_BLOCK_MIN(result) do { uint32_t m_ = 0xffffffffu; for (int t_ = 0; t_ < C::THREADS; ++t_) m_ = T.emu[t_].sin < m_ ? T.emu[t_].sin : m_; result = m_; } while (0)
#define XM_BLOCK_OR(result) do { uint32_t m_ = 0; for (int t_ = 0; t_ < C::THREADS; ++t_) m_ |= T.emu[t_].sin; result = m_; } while (0)
#endif

#if defined(__CUDACC__)
/* device collectives and the warp copy engine: xm_kernels.cu */
__device__ uint32_t dev_block_scan(uint32_t v, uint32_t *scr, uint32_t &total);
__device__ uint32_t dev_block_min(uint32_t v, uint32_t *scr);
__device__ uint32_t dev_block_or(uint32_t v, uint32_t *scr);
__device__ void dev_lookback1(unsigned long long *desc, uint32_t tile, unsigned long long agg_count, bool agg_stop,
                              unsigned long long *out /* [0] exclusive count, [1] exclusive stop */);
__device__ void dev_lookback2(uint32_t *flag, unsigned long long *agg, unsigned long long *inc, uint32_t tile,
                              unsigned long long *tot_to_base /* in: totals[C2_SLOTS]; out: exclusive bases */);
__device__ void dev_warp_copy(uint8_t *dst, const uint8_t *src_smem, const uint8_t *src_glob, uint32_t len);
#endif

XM_HD uint4 load16(const uint8_t *p)
{
    uint4 v;
#if XM_DEVICE_PASS
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
#else
    memcpy(&v, p, 16);
#endif
    return v;
}
XM_HD void atomic_add64(unsigned long long *p, unsigned long long v)
{
#if XM_DEVICE_PASS
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
XM_HD void hist_add(uint32_t *hist, int k)
{
#if XM_DEVICE_PASS
    atomicAdd(&hist[k], 1u);
#else
    hist[k]++;
#endif
}

/* ---- window geometry of a tile ------------------------------------------- */
struct Geo {
    uint64_t tstart, g0;
    uint32_t hoff, wbytes, nbits;
    int virt;
    bool at_eof;
};
template <class C>
XM_HD Geo tile_geo(const StreamBuf &B, uint32_t tile)
{
    Geo g;
    g.tstart = (uint64_t)tile * C::TILE;
    g.g0 = g.tstart >= (uint64_t)C::HALO ? g.tstart - C::HALO : 0;
    g.hoff = (uint32_t)(g.tstart - g.g0);
    uint64_t wend = g.g0 + C::WIN;
    if (wend >= B.len) { wend = B.len; g.at_eof = true; } else g.at_eof = false;
    g.wbytes = (uint32_t)(wend - g.g0);
    /* an unterminated last line gets a virtual newline one past the data */
    g.virt = (g.at_eof && B.len > 0 && B.p[B.len - 1] != '\n') ? (int)g.wbytes : -1;
    g.nbits = g.wbytes + (g.virt >= 0 ? 1u : 0u);
    return g;
}

/* i-th owned line start (window-relative) from the start mask and its group counts */
template <class C>
XM_HD int select_start(const TileMem<C> &m, uint32_t hoff, uint32_t i)
{
    int lo = 0, hi = C::NG;            /* largest group t with grp[t] <= i */
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (m.grp[mid] <= i) lo = mid; else hi = mid;
    }
    uint32_t r = i - m.grp[lo];
    int w = lo * 4;
    uint32_t bits = m.stm[w];
    for (;;) {
        uint32_t c = (uint32_t)popc32(bits);
        if (r < c) break;
        r -= c;
        bits = m.stm[++w];
    }
    for (; r; --r) bits &= bits - 1;
    return (int)hoff + (w << 5) + ffs32(bits) - 1;
}

/* global offset where the line ending just before window position s0 starts */
XM_HD uint64_t prev_line_start(const uint32_t *nlm, const uint8_t *glob, uint64_t g0, int s0)
{
    const int q = s0 - 1;              /* the newline that ends the previous line; look for the one before it */
    if (q > 0) {
        int w = (q - 1) >> 5;
        uint32_t m = nlm[w] & (0xffffffffu >> (31 - ((q - 1) & 31)));
        for (;;) {
            if (m) return g0 + (uint64_t)((w << 5) + 32 - clz32(m));
            if (--w < 0) break;
            m = nlm[w];
        }
    }
    uint64_t p = g0;                   /* not inside the window: walk back through global memory */
    while (p > 0 && glob[p - 1] != '\n') --p;
    return p;
}

template <class C>
struct Front {
    Geo geo;
    uint32_t nlines;       /* owned lines considered (0 on overflow) */
    uint32_t n_eff;        /* owned lines before the first blank one */
    bool stop;             /* a blank line ends the stream inside this tile */
    bool overflow;
};

/* ---- front end shared by both kernels: stage, mask, index, parse --------- */
template <class C>
XM_HD void front_end(TileCtx<C> &T, ThreadState<C> &th_, const StreamBuf &B, int score_src, uint32_t debug,
                     bool need_prev, Front<C> &fr)
{
    (void)th_;
    const Geo geo = fr.geo;
    /* stage the window and build the two byte-class masks, 16 bytes per step */
    XM_THREADS_BEGIN
        for (int c = tid; c < C::NSLOT; c += C::THREADS) {
            const uint32_t off = 16u * (uint32_t)c;
            uint32_t mc = 0, mn = 0;
            if (off < geo.wbytes) {
                const uint4 v = load16(B.p + geo.g0 + off);
                *(uint4 *)(T.m.win + off) = v;
                const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
                uint32_t zn[4], anyn = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    mc |= pack4(ctrl_mask(w4[q])) << (4 * q);
                    zn[q] = eq_mask(w4[q], 0x0a0a0a0au);
                    anyn |= zn[q];
                }
                if (anyn) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) mn |= pack4(zn[q]) << (4 * q);
                }
                const uint32_t valid = geo.wbytes - off;
                if (valid < 16u) { mc &= (1u << valid) - 1u; mn &= (1u << valid) - 1u; }
            }
            if (geo.virt >= 0 && (uint32_t)c == (geo.wbytes >> 4)) {
                mc |= 1u << (geo.wbytes & 15u);
                mn |= 1u << (geo.wbytes & 15u);
            }
            ((uint16_t *)T.m.wsm)[c] = (uint16_t)mc;
            ((uint16_t *)T.m.nlm)[c] = (uint16_t)mn;
        }
    XM_THREADS_END
    XM_BARRIER();
    /* owned line starts: the byte after every newline, inside the tile and inside the stream */
    XM_THREADS_BEGIN
        uint32_t cnt = 0;
        if (tid < C::NG) {
            uint32_t limit = geo.hoff + (uint32_t)C::TILE;
            if (limit > geo.wbytes) limit = geo.wbytes;
            for (int j = 0; j < 4; ++j) {
                const int W = (int)(geo.hoff >> 5) + tid * 4 + j;
                uint32_t sw = T.m.nlm[W] << 1;
                if (W > 0) sw |= T.m.nlm[W - 1] >> 31;
                else if (geo.g0 == 0) sw |= 1u;                 /* the stream's first byte starts a line */
                const uint32_t base = (uint32_t)W << 5;
                if (base >= limit) sw = 0;
                else if (limit - base < 32u) sw &= (1u << (limit - base)) - 1u;
                T.m.stm[tid * 4 + j] = sw;
                cnt += (uint32_t)popc32(sw);
            }
        }
        th.sin = cnt;
    XM_THREADS_END
    {
        uint32_t tot_;
        XM_BLOCK_SCAN(tot_);
        fr.nlines = tot_;
    }
    XM_THREADS_BEGIN
        if (tid < C::NG) T.m.grp[tid] = (uint16_t)th.sout;
        if (tid == 0) T.m.grp[C::NG] = (uint16_t)fr.nlines;
    XM_THREADS_END
    XM_BARRIER();
    fr.overflow = fr.nlines > (uint32_t)C::LCAP;
    if (fr.overflow) fr.nlines = 0;
    /* parse the owned lines; thread 0 also parses the line before the first one when the walk needs it */
    XM_THREADS_BEGIN
        const WinMasks M_{T.m.win, T.m.wsm, T.m.nlm, (int)geo.nbits, geo.virt};
        const Reader rd_{T.m.win, B.p, geo.g0, geo.wbytes, B.len};
        uint32_t fb = 0xffffffffu;
        for (int j = 0; j < C::R; ++j) {
            const uint32_t i = (uint32_t)(j * C::THREADS + tid);
            th.rank[j] = NOT_YIELDED;
            if (i < fr.nlines) {
                const int s = select_start<C>(T.m, geo.hoff, i);
                if ((debug & DBG_FORCE_GENERIC) || !fast_parse(M_, s, score_src, th.L[j]))
                    generic_parse(rd_, geo.g0 + (uint64_t)s, score_src, th.L[j]);
                if ((th.L[j].flags & F_BLANK) && i < fb) fb = i;
            }
        }
        th.sin = fb;
        if (tid == 0) {
            T.m.scr[SCR_HALO_VALID] = 0;
            if (need_prev && fr.nlines > 0 && geo.g0 + th.L[0].s > 0) {
                const uint64_t ps = prev_line_start(T.m.nlm, B.p, geo.g0, (int)th.L[0].s);
                LineRec H;
                Reader rh_ = rd_;
                const bool far = ps < geo.g0;
                if (far) { rh_.wbytes = 0; rh_.g0 = ps; }       /* read it all from global memory */
                if (far || (debug & DBG_FORCE_GENERIC) || !fast_parse(M_, (int)(ps - geo.g0), score_src, H))
                    generic_parse(rh_, ps, score_src, H);
                T.m.scr64[S64_HALO_START] = ps;
                T.m.scr64[S64_HALO_QS] = rh_.g0 + H.qs;
                T.m.scr[SCR_HALO_QLEN] = H.qlen; T.m.scr[SCR_HALO_H1] = H.h1; T.m.scr[SCR_HALO_H2] = H.h2;
                T.m.scr[SCR_HALO_FLAGS] = H.flags; T.m.scr[SCR_HALO_OUTLEN] = H.outlen;
                T.m.scr[SCR_HALO_AS] = (uint32_t)H.as; T.m.scr[SCR_HALO_XS] = (uint32_t)H.xs;
                T.m.scr[SCR_HALO_VALID] = 1;
            }
        }
    XM_THREADS_END
    {
        uint32_t fb_;
        XM_BLOCK_MIN(fb_);
        fr.stop = fb_ < fr.nlines;
        fr.n_eff = fr.stop ? fb_ : fr.nlines;
    }
    /* where the stream stops, if it does so here (needed after the masks are gone) */
    XM_THREADS_BEGIN
        if (tid == 0) T.m.scr64[S64_BLANK_OFF] = fr.stop ? geo.g0 + (uint64_t)select_start<C>(T.m, geo.hoff, fr.n_eff) : B.len;
    XM_THREADS_END
}

/* Ranks among the records the tile yields: every line before the stop, or only
 * the first line of each run of equal QNAMEs (getReadPairs skip mode,
 * xm.py:110-114).  Expects shl[i+1] = (qs, qlen, h1, h2) of line i.  Returns the count. */
template <class C>
XM_HD uint32_t rank_lines(TileCtx<C> &T, ThreadState<C> &th_, const StreamBuf &B, const Front<C> &fr, bool skip)
{
    (void)th_;
    if (!skip) {
        XM_THREADS_BEGIN
            for (int j = 0; j < C::R; ++j) {
                const uint32_t i = (uint32_t)(j * C::THREADS + tid);
                if (i < fr.n_eff) th.rank[j] = i;
            }
        XM_THREADS_END
        return fr.n_eff;
    }
    uint32_t carry = 0;
    for (int j = 0; j < C::R; ++j) {
        if ((uint32_t)(j * C::THREADS) >= fr.n_eff) break;
        XM_THREADS_BEGIN
            const uint32_t i = (uint32_t)(j * C::THREADS + tid);
            uint32_t head = 0;
            if (i < fr.n_eff) {
                const Reader rd_{T.m.win, B.p, fr.geo.g0, fr.geo.wbytes, B.len};
                const LineRec &L = th.L[j];
                head = 1;
                if (i == 0) {
                    if (T.m.scr[SCR_HALO_VALID] && T.m.scr[SCR_HALO_QLEN] == L.qlen && T.m.scr[SCR_HALO_H1] == L.h1 && T.m.scr[SCR_HALO_H2] == L.h2)
                        head = !names_equal(rd_, T.m.scr64[S64_HALO_QS], L.qlen, fr.geo.g0 + L.qs, L.qlen);
                } else {
                    const uint4 pv = T.m.shl[i];
                    if (pv.y == L.qlen && pv.z == L.h1 && pv.w == L.h2)
                        head = !names_equal(rd_, fr.geo.g0 + pv.x, pv.y, fr.geo.g0 + L.qs, L.qlen);
                }
            }
            th.sin = head;
        XM_THREADS_END
        uint32_t tot;
        XM_BLOCK_SCAN(tot);
        XM_THREADS_BEGIN
            const uint32_t i = (uint32_t)(j * C::THREADS + tid);
            if (i < fr.n_eff && th.sin) th.rank[j] = carry + th.sout;
        XM_THREADS_END
        carry += tot;
    }
    return carry;
}

/* ---- emulated look-back for the CPU build: tiles run in order ------------- */
#if !XM_DEVICE_PASS
inline void emu_lookback1(unsigned long long *desc, uint32_t tile, unsigned long long agg_count, bool agg_stop, unsigned long long *out)
{
    unsigned long long ec = 0;
    bool es = false;
    if (tile > 0) { ec = desc[tile - 1] & C1_COUNT; es = (desc[tile - 1] & C1_STOP) != 0; }
    out[0] = ec; out[1] = es;
    desc[tile] = C1_INC | ((es || agg_stop) ? C1_STOP : 0) | (es ? ec : ec + agg_count);
}
inline void emu_lookback2(uint32_t *flag, unsigned long long *agg, unsigned long long *inc, uint32_t tile, unsigned long long *tot)
{
    (void)agg;
    for (int b = 0; b < C2_SLOTS; ++b) {
        unsigned long long ex = tile ? inc[(size_t)(tile - 1) * C2_SLOTS + b] : 0;
        inc[(size_t)tile * C2_SLOTS + b] = ex + tot[b];
        tot[b] = ex;
    }
    flag[tile] = 2;
}
#endif

/* =========================================================================
 * Secondary-stream scan.
 * ========================================================================= */
template <class C>
XM_HD void scan_tile(TileCtx<C> &T, const ScanArgs &a, uint32_t tile)
{
#if XM_DEVICE_PASS
    ThreadState<C> th_;
#else
    ThreadState<C> &th_ = T.emu[0];
#endif
    Front<C> fr;
    fr.geo = tile_geo<C>(a.S, tile);
    front_end<C>(T, th_, a.S, a.score_src, a.debug, a.skip != 0, fr);

    if (a.skip) {
        XM_THREADS_BEGIN
            for (int j = 0; j < C::R; ++j) {
                const uint32_t i = (uint32_t)(j * C::THREADS + tid);
                if (i < fr.n_eff) T.m.shl[i + 1] = make_uint4(th.L[j].qs, th.L[j].qlen, th.L[j].h1, th.L[j].h2);
            }
        XM_THREADS_END
        XM_BARRIER();
    }
    const uint32_t count = rank_lines<C>(T, th_, a.S, fr, a.skip != 0);

    /* record base of this tile: first look-back chain (count + stop flag) */
    unsigned long long base, pstop;
#if XM_DEVICE_PASS
    if (threadIdx.x < 32) dev_lookback1(a.chain1, tile, count, fr.stop, T.m.scr64 + S64_BASE);
    __syncthreads();
    base = T.m.scr64[S64_BASE]; pstop = T.m.scr64[S64_PSTOP];
#else
    { unsigned long long o[2]; emu_lookback1(a.chain1, tile, count, fr.stop, o); base = o[0]; pstop = o[1]; }
#endif

    XM_THREADS_BEGIN
        if (!pstop) {
            for (int j = 0; j < C::R; ++j) {
                if (th.rank[j] == NOT_YIELDED) continue;
                const unsigned long long gi = base + th.rank[j];
                if (gi >= a.sc_cap) continue;
                const LineRec &L = th.L[j];
                a.sc.start[gi] = fr.geo.g0 + L.s;
                a.sc.as[gi] = L.as; a.sc.xs[gi] = L.xs;
                a.sc.h1[gi] = L.h1; a.sc.h2[gi] = L.h2;
                a.sc.meta[gi] = (L.outlen & META_LEN_MASK) | ((L.flags & 0x3fu) << META_LEN_BITS);
            }
            if (tid == 0 && (fr.stop || tile + 1 == a.ntiles)) {
                const unsigned long long n = base + count;
                a.g->n_stream[a.stream_id] = n;
                a.g->end_off[a.stream_id] = T.m.scr64[S64_BLANK_OFF];
                if (a.sc.start && n <= a.sc_cap) a.sc.start[n] = T.m.scr64[S64_BLANK_OFF];
            }
        }
        if (tid == 0 && fr.overflow) a.g->overflow = 1;
    XM_THREADS_END
}

/* =========================================================================
 * Primary-stream classify + emit.
 * ========================================================================= */
XM_HD void report_error(Globals *g, unsigned long long rec, int code, int stream, int prev = 0)
{
    const unsigned long long w = (rec << 8) | ((unsigned long long)code << 2) | ((unsigned long long)prev << 1) | (unsigned long long)stream;
#if XM_DEVICE_PASS
    atomicMin(&g->err, w);
#else
    if (w < g->err) g->err = w;
#endif
}

/* first tag-evaluation failure of a record, in the reference's call order AS1, XS1, AS2, XS2 (xm.py:323-326) */
XM_HD void eval_error(uint32_t pflags, uint32_t sflags, int &code, int &stream)
{
    code = 0; stream = 0;
    if (pflags & F_AS_DUP) code = EC_DUP; else if (pflags & F_AS_NUM) code = EC_NUM_AS;
    else if (pflags & F_XS_DUP) code = EC_DUP; else if (pflags & F_XS_NUM) code = EC_NUM_XS;
    if (code) return;
    if (sflags & F_AS_DUP) code = EC_DUP; else if (sflags & F_AS_NUM) code = EC_NUM_AS;
    else if (sflags & F_XS_DUP) code = EC_DUP; else if (sflags & F_XS_NUM) code = EC_NUM_XS;
    if (code) stream = 1;
}

/* byte-wise writer for lines whose output differs from their raw bytes:
 * tokens joined by single tabs + '\n' (xm.py:334).  gs may point at the
 * line's first byte or its first token. */
XM_HD void write_normalised(const Reader &rd, uint64_t gs, int nlines, uint8_t *dst)
{
    uint64_t p = gs;
    for (int l = 0; l < nlines; ++l) {
        bool in_tok = false, first = true;
        for (;; ++p) {
            bool end = p >= rd.len;
            const uint32_t c = end ? (uint32_t)'\n' : rd.at(p);
            if (c == '\n') end = true;
            const bool ws = end || c == '\r' || (c < 0x80 && is_ascii_space((uint8_t)c));
            if (ws) {
                in_tok = false;
                if (end) break;
            } else {
                if (!in_tok) { if (!first) *dst++ = '\t'; first = false; in_tok = true; }
                *dst++ = (uint8_t)c;
            }
        }
        *dst++ = '\n';
        ++p;
    }
}

template <class C>
XM_HD void classify_tile(TileCtx<C> &T, const ClassifyArgs &a, uint32_t tile)
{
#if XM_DEVICE_PASS
    ThreadState<C> th_;
#else
    ThreadState<C> &th_ = T.emu[0];
#endif
    const bool paired = a.mode != MODE_SE;
    Front<C> fr;
    fr.geo = tile_geo<C>(a.P, tile);
    front_end<C>(T, th_, a.P, a.score_src, a.debug, paired || a.skip, fr);

    XM_THREADS_BEGIN
        if (tid < 36) T.m.hist[tid] = 0;
        if (tid == 0) T.m.scr64[S64_RAW] = 0;
        for (int j = 0; j < C::R; ++j) {
            const uint32_t i = (uint32_t)(j * C::THREADS + tid);
            if (i < fr.n_eff) T.m.shl[i + 1] = make_uint4(th.L[j].qs, th.L[j].qlen, th.L[j].h1, th.L[j].h2);
        }
    XM_THREADS_END
    XM_BARRIER();
    const uint32_t count = rank_lines<C>(T, th_, a.P, fr, a.skip != 0);

    unsigned long long base, pstop;
#if XM_DEVICE_PASS
    if (threadIdx.x < 32) dev_lookback1(a.chain1, tile, count, fr.stop, T.m.scr64 + S64_BASE);
    __syncthreads();
    base = T.m.scr64[S64_BASE]; pstop = T.m.scr64[S64_PSTOP];
#else
    { unsigned long long o[2]; emu_lookback1(a.chain1, tile, count, fr.stop, o); base = o[0]; pstop = o[1]; }
#endif
    /* records at or beyond ncap are not yielded: the other stream ended first, or an error re-run cut here */
    unsigned long long ncap = a.g->n_stream[1];
    if (a.limit < ncap) ncap = a.limit;
    if (pstop) ncap = 0;

    /* per-record decision.  Paired walks publish it for the next line's owner. */
    XM_THREADS_BEGIN
        for (int j = 0; j < C::R; ++j) {
            th.ymeta[j] = NO_BIN; th.ybytes[j] = 0; th.yoff[j] = 0;
            if (th.rank[j] == NOT_YIELDED) continue;
            const unsigned long long gi = base + th.rank[j];
            const LineRec &L = th.L[j];
            if (gi == a.limit) a.g->limit_off = fr.geo.g0 + L.s;
            if (gi >= ncap) { th.rank[j] = NOT_YIELDED; continue; }
            const uint32_t smeta = a.sc.meta[gi];
            const uint32_t sflags = smeta >> META_LEN_BITS;
            int ec = 0;
            /* text the device does not tokenise like the reference comes first: nothing derived from it can be trusted */
            if ((L.flags | sflags) & F_TEXT) ec = EC_TEXT;
            else if (a.sc.h1[gi] != L.h1 || a.sc.h2[gi] != L.h2) ec = EC_ASSERT;     /* xm.py:106 */
            if (ec) report_error(a.g, gi, ec, (ec == EC_TEXT && !(L.flags & F_TEXT)) ? 1 : 0);
            int ev, evs;
            eval_error(L.flags, sflags, ev, evs);
            const int st = mapping_state(L.as, L.xs, a.sc.as[gi], a.sc.xs[gi], a.thr);
            atomic_add64(&T.m.scr64[S64_RAW], (unsigned long long)L.rawbytes);
            if (!paired) {
                if (ec) continue;
                if (ev) { report_error(a.g, gi, ev, evs); continue; }
                hist_add(T.m.hist, st);
                const uint32_t slen = smeta & META_LEN_MASK;
                uint32_t bytes = (st == PS || st == PM || st == UA) ? L.outlen : (st == UR ? L.outlen + slen : slen);
                if (!((a.enabled >> st) & 1u)) bytes = 0;
                th.ymeta[j] = (uint32_t)st;
                th.ybytes[j] = bytes;
            } else {
                th.yoff[j] = shl_pack(L.outlen, st, ev, evs, (L.flags & F_DIRTY) != 0);   /* parked until published */
                if (ec) th.ymeta[j] |= Y_ASSERT;
            }
        }
    XM_THREADS_END

    if (paired) {
        XM_THREADS_BEGIN
            for (int j = 0; j < C::R; ++j) {
                const uint32_t i = (uint32_t)(j * C::THREADS + tid);
                if (i < fr.n_eff) {
                    const bool y = th.rank[j] != NOT_YIELDED;
                    T.m.shl[i + 1] = make_uint4(th.L[j].qs, y ? th.L[j].qlen : 0xffffffffu, th.L[j].h1, y ? th.yoff[j] : 0u);
                }
            }
            if (tid == 0) {
                /* the line before the first owned one is record base-1 */
                uint4 h = make_uint4(0, 0xffffffffu, 0, 0);
                if (!a.skip && T.m.scr[SCR_HALO_VALID] && base > 0 && base - 1 < ncap) {
                    const unsigned long long gi = base - 1;
                    int ev, evs;
                    eval_error(T.m.scr[SCR_HALO_FLAGS], a.sc.meta[gi] >> META_LEN_BITS, ev, evs);
                    const int st = mapping_state((int32_t)T.m.scr[SCR_HALO_AS], (int32_t)T.m.scr[SCR_HALO_XS], a.sc.as[gi], a.sc.xs[gi], a.thr);
                    h = make_uint4(0, T.m.scr[SCR_HALO_QLEN], T.m.scr[SCR_HALO_H1],
                                   shl_pack(T.m.scr[SCR_HALO_OUTLEN], st, ev, evs, (T.m.scr[SCR_HALO_FLAGS] & F_DIRTY) != 0));
                }
                T.m.shl[0] = h;
            }
        XM_THREADS_END
        XM_BARRIER();
        XM_THREADS_BEGIN
            for (int j = 0; j < C::R; ++j) {
                const uint32_t i = (uint32_t)(j * C::THREADS + tid);
                const bool own_ec = (th.ymeta[j] & Y_ASSERT) != 0;
                th.yoff[j] = 0;
                th.ymeta[j] = NO_BIN;
                /* with skip_repeated the yielded records are run heads: adjacent ones never share a QNAME */
                if (a.skip || th.rank[j] == NOT_YIELDED) continue;
                const unsigned long long gi = base + th.rank[j];
                if (gi == 0) continue;
                const LineRec &L = th.L[j];
                const uint4 pv = T.m.shl[i];
                /* a unit fires when the previous yielded primary QNAME equals this one (xm.py:402) */
                if (pv.y != L.qlen || pv.z != L.h1) continue;
                const Reader rd_{T.m.win, a.P.p, fr.geo.g0, fr.geo.wbytes, a.P.len};
                const uint64_t pq = (i == 0) ? T.m.scr64[S64_HALO_QS] : fr.geo.g0 + pv.x;
                if (!names_equal(rd_, pq, pv.y, fr.geo.g0 + L.qs, L.qlen)) continue;
                if (own_ec) continue;                       /* the assert fires before the unit is looked at */
                const int pst = (int)((pv.w >> 24) & 7u), pev = (int)((pv.w >> 27) & 7u), pevs = (int)(pv.w >> 31);
                const uint32_t smeta = a.sc.meta[gi], pmeta = a.sc.meta[gi - 1];
                int ev, evs;
                eval_error(L.flags, smeta >> META_LEN_BITS, ev, evs);
                if (pev) { report_error(a.g, gi, pev, pevs, 1); continue; }     /* xm.py:408-411 come first */
                if (ev) { report_error(a.g, gi, ev, evs); continue; }
                const int st = mapping_state(L.as, L.xs, a.sc.as[gi], a.sc.xs[gi], a.thr);
                hist_add(T.m.hist, pst * 6 + st);
                const int bin = a.mode == MODE_PE_CONSERVATIVE ? pair_bin_conservative(pst, st) : pair_bin_liberal(pst, st);
                const uint32_t plen = (pv.w & META_LEN_MASK) + L.outlen;
                const uint32_t slen = (pmeta & META_LEN_MASK) + (smeta & META_LEN_MASK);
                uint32_t bytes = (bin == PS || bin == PM || bin == UA) ? plen : (bin == UR ? plen + slen : slen);
                if (!((a.enabled >> bin) & 1u)) bytes = 0;
                th.ymeta[j] = (uint32_t)bin | (((pv.w >> 30) & 1u) ? Y_PREV_DIRTY : 0u);
                th.ybytes[j] = bytes;
            }
        XM_THREADS_END
    }

    /* byte offsets inside each bin: block scans over the bins that occur, then the second look-back chain */
    unsigned long long tot[C2_SLOTS];
    for (int b = 0; b < C2_SLOTS; ++b) tot[b] = 0;
    {
        uint32_t present;
        XM_THREADS_BEGIN
            uint32_t pm = 0;
            for (int j = 0; j < C::R; ++j)
                if ((th.ymeta[j] & 7u) != NO_BIN && th.ybytes[j]) pm |= 1u << (th.ymeta[j] & 7u);
            th.sin = pm;
        XM_THREADS_END
        XM_BLOCK_OR(present);
        for (int j = 0; j < C::R; ++j) {
            if ((uint32_t)(j * C::THREADS) >= fr.n_eff) break;
            for (int b = 0; b < 6; ++b) {
                if (!((present >> b) & 1u)) continue;
                XM_THREADS_BEGIN
                    th.sin = ((th.ymeta[j] & 7u) == (uint32_t)b) ? th.ybytes[j] : 0u;
                XM_THREADS_END
                uint32_t t2;
                XM_BLOCK_SCAN(t2);
                XM_THREADS_BEGIN
                    if ((th.ymeta[j] & 7u) == (uint32_t)b) th.yoff[j] = (uint32_t)tot[b] + th.sout;
                XM_THREADS_END
                tot[b] += t2;
            }
        }
    }
    XM_BARRIER();
    tot[6] = T.m.scr64[S64_RAW];
    unsigned long long tsum[C2_SLOTS];
    for (int b = 0; b < C2_SLOTS; ++b) tsum[b] = tot[b];
#if XM_DEVICE_PASS
    __syncthreads();
    if (threadIdx.x < 32) {
        if (threadIdx.x < C2_SLOTS) T.m.scr64[S64_TOT + threadIdx.x] = tot[threadIdx.x];
        __syncwarp();
        dev_lookback2(a.c2_flag, a.c2_agg, a.c2_inc, tile, T.m.scr64 + S64_TOT);
    }
    __syncthreads();
    for (int b = 0; b < C2_SLOTS; ++b) tot[b] = T.m.scr64[S64_TOT + b];
#else
    emu_lookback2(a.c2_flag, a.c2_agg, a.c2_inc, tile, tot);
#endif

    /* copy items: at most two per owning line (its primary-stream part, its secondary-stream part).
     * They overlay the masks, which nobody reads any more. */
    const uint32_t nown = fr.n_eff;
    XM_THREADS_BEGIN
        for (int j = 0; j < C::R; ++j) {
            const uint32_t i = (uint32_t)(j * C::THREADS + tid);
            if (i >= nown) continue;
            uint32_t m0 = 0, m1 = 0, d0 = 0, d1 = 0, s0 = 0, s1 = 0;
            const uint32_t bin = th.ymeta[j] & 7u;
            if (th.rank[j] != NOT_YIELDED && bin != NO_BIN && th.ybytes[j]) {
                const LineRec &L = th.L[j];
                const bool pside = bin == PS || bin == PM || bin == UA || bin == UR;
                const bool sside = bin == SS || bin == SM || bin == UR;
                uint32_t plen = 0;
                if (pside) {
                    uint32_t prevlen = 0;
                    bool dirty = (L.flags & F_DIRTY) != 0;
                    if (paired) { prevlen = T.m.shl[i].w & META_LEN_MASK; dirty = dirty || (th.ymeta[j] & Y_PREV_DIRTY); }
                    plen = prevlen + L.outlen;
                    d0 = th.yoff[j];
                    if (!dirty) {
                        s0 = L.s - prevlen;                 /* clean neighbours are contiguous in the input */
                        m0 = plen | (bin << 27) | ((uint32_t)IT_P_COPY << 30);
                    } else {
                        if (!paired) s0 = L.qs;
                        else s0 = (i == 0) ? (uint32_t)(T.m.scr64[S64_HALO_QS] - fr.geo.g0) : T.m.shl[i].x;
                        m0 = plen | (paired ? (1u << 26) : 0u) | (bin << 27) | ((uint32_t)IT_P_NORM << 30);
                    }
                }
                if (sside) {
                    d1 = th.yoff[j] + plen;
                    s1 = th.rank[j];
                    m1 = (th.ybytes[j] - plen) | (paired ? (1u << 26) : 0u) | (bin << 27) | ((uint32_t)IT_S << 30);
                }
            }
            T.m.it_dst[2 * i] = d0; T.m.it_src[2 * i] = s0; T.m.it_meta[2 * i] = m0;
            T.m.it_dst[2 * i + 1] = d1; T.m.it_src[2 * i + 1] = s1; T.m.it_meta[2 * i + 1] = m1;
        }
    XM_THREADS_END
    XM_BARRIER();

    /* copy: one warp per item */
    {
#if XM_DEVICE_PASS
        const int lane = (int)(threadIdx.x & 31);
        for (uint32_t k = threadIdx.x >> 5; k < 2u * nown; k += (uint32_t)(C::THREADS / 32)) {
#else
        const int lane = 0;
        for (uint32_t k = 0; k < 2u * nown; ++k) {
#endif
            const uint32_t meta = T.m.it_meta[k];
            const uint32_t kind = meta >> 30;
            if (kind == IT_EMPTY) continue;
            const uint32_t len = meta & ((1u << 26) - 1u);
            const int bin = (int)((meta >> 27) & 7u);
            const int nl = ((meta >> 26) & 1u) ? 2 : 1;
            const unsigned long long doff = tot[bin] + T.m.it_dst[k];
            if (doff + len > a.out_cap[bin]) continue;      /* needed sizes are still reported; the host rejects the call */
            uint8_t *dst = a.out[bin] + doff;
            if (kind == IT_P_COPY) {
                const long long so = (long long)(int32_t)T.m.it_src[k];
#if XM_DEVICE_PASS
                const bool inwin = so >= 0 && (unsigned long long)so + len <= fr.geo.wbytes;
                dev_warp_copy(dst, inwin ? T.m.win + so : nullptr, a.P.p + fr.geo.g0 + so, len);
#else
                memcpy(dst, a.P.p + fr.geo.g0 + so, len);
#endif
            } else if (kind == IT_P_NORM) {
                if (lane == 0) {
                    const Reader rd_{T.m.win, a.P.p, fr.geo.g0, fr.geo.wbytes, a.P.len};
                    write_normalised(rd_, fr.geo.g0 + (long long)(int32_t)T.m.it_src[k], nl, dst);
                }
            } else {
                const unsigned long long gi = base + T.m.it_src[k];
                const unsigned long long g_first = nl == 2 ? gi - 1 : gi;
                const uint32_t fl = (a.sc.meta[gi] | a.sc.meta[g_first]) >> META_LEN_BITS;
                const uint64_t ss = a.sc.start[g_first];
                if (!(fl & F_DIRTY)) {
#if XM_DEVICE_PASS
                    dev_warp_copy(dst, nullptr, a.S.p + ss, len);
#else
                    memcpy(dst, a.S.p + ss, len);
#endif
                } else if (lane == 0) {
                    const Reader rs_{nullptr, a.S.p, 0, 0, a.S.len};
                    write_normalised(rs_, ss, nl, dst);
                }
            }
        }
    }

    /* category histogram and stream totals */
    XM_THREADS_BEGIN
        if (tid < 36 && T.m.hist[tid]) atomic_add64(&a.g->counts[tid], (unsigned long long)T.m.hist[tid]);
        if (tid == 0) {
            if (tsum[6]) atomic_add64(&a.g->bytes_in[0], tsum[6]);
            const bool last = tile + 1 == a.ntiles;
            if (!pstop && (fr.stop || last)) {
                a.g->n_stream[0] = base + count;
                a.g->end_off[0] = T.m.scr64[S64_BLANK_OFF];
            }
            if (last) for (int b = 0; b < 6; ++b) a.g->out_len[b] = tot[b] + tsum[b];
            if (fr.overflow) a.g->overflow = 1;
        }
    XM_THREADS_END
}

}  // namespace xm
