/*
 * xm_inflate.h -- BGZF blocks inflated on the GPU (BAM input, BASELINE configs[4]).
 *
 * The reference reads BAM through an external `samtools view` (xm.py:48-64), which inflates
 * the file's BGZF blocks one after the other on a host core.  A BGZF block is an independent
 * raw-DEFLATE stream (RFC 1951) of at most 64 KiB of data with its inflated size and CRC-32
 * in the trailer (SAM specification, section 4.1): the blocks of a file are decoded here all
 * at once, ONE WARP PER BLOCK.
 *
 *   - the warp's lanes run the symbol decode in lock step (same bit buffer, same table
 *     look-ups: broadcast loads, no divergence); lane 0 stores the literals, all lanes copy a
 *     match together (byte i of the match by lane i % 32, periodic source for distances
 *     shorter than the match);
 *   - per warp in shared memory: a 10-bit look-up table for the literal/length code and an
 *     8-bit one for the distance code (symbol + code length per entry); codes longer than
 *     that are found by comparing the next 15 bits with one limit per code length;
 *   - CRC-32 of the inflated block: each lane folds a 1/32 slice byte by byte, the slices are
 *     joined with the "append n zero bytes" operator of the CRC's linear feedback register.
 *
 * The decoder is plain host/device code (XM_HD): tests/test_inflate.py runs it on the CPU
 * against zlib on every block type; on the device it is driven by k_bgzf_inflate below.
 * Out-of-range reads cannot happen on corrupt input: every store is bounded by the block's
 * ISIZE, every match source by the bytes already written, every fetch by the block's BSIZE.
 */
#pragma once
#include <stdint.h>

#include "xm_common.h"

namespace xm {

constexpr int INF_LIT_BITS = 10, INF_DIST_BITS = 8;
enum { INF_OK = 0, INF_E_HEADER = 1, INF_E_CODES = 2, INF_E_SYMBOL = 3, INF_E_DISTANCE = 4, INF_E_OVERRUN = 5, INF_E_INPUT = 6,
       INF_E_SHORT = 7, INF_E_CRC = 8 };

/* decode tables of one warp (shared memory on the device) */
struct InflateTables {
    uint16_t lit_lut[1 << INF_LIT_BITS];        /* symbol | code length << 9; 0: the code is longer than the table's index */
    uint16_t dist_lut[1 << INF_DIST_BITS];
    uint16_t lit_limit[16], dist_limit[16];     /* per code length l: (first code + codes of that length) << (15 - l); while building: the counts */
    int16_t lit_delta[16], dist_delta[16];      /* index in sym[] of the first symbol of length l, minus the first code of length l */
    uint16_t lit_sym[288], dist_sym[32];        /* symbols in canonical order */
    uint8_t lens[320];                          /* code lengths as read */
    uint32_t fixed;                             /* the tables hold the fixed code */
};
/* base value << 8 | extra bits of the 29 length symbols ([0..28]) and the 30 distance symbols ([32..61]); one table per CTA */
constexpr int INF_BASE_WORDS = 64;

#if XM_DEVICE_PASS
#define XM_INF_LANE ((int)(threadIdx.x & 31))
#define XM_INF_LANES 32
#define XM_INF_SYNC() __syncwarp()
#else
#define XM_INF_LANE 0
#define XM_INF_LANES 1
#define XM_INF_SYNC() ((void)0)
#endif

XM_HD uint32_t inf_bitrev(uint32_t v, int n)
{
#if XM_DEVICE_PASS
    return __brev(v) >> (32 - n);
#else
    uint32_t r = 0;
    for (int k = 0; k < n; ++k) r |= ((v >> k) & 1u) << (n - 1 - k);
    return r;
#endif
}

/* lengths 3..258: eight symbols without extra bits, then four per extra-bit count (RFC 1951, 3.2.5); s = symbol - 257 */
XM_HD uint32_t inf_len_base(int s)
{
    if (s < 8) return 3u + (uint32_t)s;
    if (s == 28) return 258u;
    const uint32_t x = (uint32_t)(s - 4) >> 2;
    return 3u + ((4u + ((uint32_t)s & 3u)) << x);
}
XM_HD int inf_len_extra(int s) { return (s < 8 || s == 28) ? 0 : (s - 4) >> 2; }
/* distances 1..32768: four symbols without extra bits, then two per extra-bit count */
XM_HD uint32_t inf_dist_base(int s)
{
    if (s < 4) return 1u + (uint32_t)s;
    const uint32_t x = (uint32_t)(s - 2) >> 1;
    return 1u + ((2u + ((uint32_t)s & 1u)) << x);
}
XM_HD int inf_dist_extra(int s) { return s < 4 ? 0 : (s - 2) >> 1; }
XM_HD uint32_t inf_base_word(int k)
{
    if (k < 29) return (inf_len_base(k) << 8) | (uint32_t)inf_len_extra(k);
    if (k >= 32 && k < 62) return (inf_dist_base(k - 32) << 8) | (uint32_t)inf_dist_extra(k - 32);
    return 0;
}

/* canonical code of `n` lengths -> limit[], delta[], sym[] and the look-up table of `bits` index bits.
 * Returns false for an over-subscribed set, or an incomplete one unless it is a lone one-bit code (zlib's rule). */
XM_HD bool inf_build(const uint8_t *lens, int n, uint16_t *limit, int16_t *delta, uint16_t *sym, uint16_t *lut, int bits)
{
    const int lane = XM_INF_LANE;
    uint16_t offs[16], count[16];
    uint16_t *cnt = limit;                       /* the counts are gathered in limit[] */
    if (lane == 0) for (int l = 0; l < 16; ++l) cnt[l] = 0;
    XM_INF_SYNC();
    if (lane == 0) for (int s = 0; s < n; ++s) cnt[lens[s]]++;
    XM_INF_SYNC();
    int left = 1, used = 0;
    for (int l = 1; l < 16; ++l) {
        count[l] = cnt[l];
        left <<= 1;
        left -= (int)count[l];
        if (left < 0) return false;
        used += count[l];
    }
    if (used == 0) return false;
    if (left > 0 && !(used == 1 && count[1] == 1)) return false;
    XM_INF_SYNC();                               /* every lane has read the counts: limit[] may be overwritten */
    /* first code of each length (MSB-first canonical code) and first slot in sym[] */
    uint32_t next_code[16];
    {
        uint32_t code = 0;
        uint16_t o = 0;
        for (int l = 1; l < 16; ++l) {
            next_code[l] = code; offs[l] = o;
            if (lane == 0) { limit[l] = (uint16_t)((code + count[l]) << (15 - l)); delta[l] = (int16_t)((int)o - (int)code); }
            code = (code + count[l]) << 1;
            o = (uint16_t)(o + count[l]);
        }
    }
    for (int k = lane; k < (1 << bits); k += XM_INF_LANES) lut[k] = 0;
    XM_INF_SYNC();
    /* lane 0 hands out the codes in symbol order; every entry of the table belongs to one symbol */
    if (lane == 0) {
        for (int s = 0; s < n; ++s) {
            const int l = lens[s];
            if (!l) continue;
            sym[offs[l]++] = (uint16_t)s;
            const uint32_t code = next_code[l]++;
            if (l <= bits) {
                const uint16_t e = (uint16_t)(s | (l << 9));
                for (uint32_t k = inf_bitrev(code, l); k < (1u << bits); k += 1u << l) lut[k] = e;
            }
        }
    }
    XM_INF_SYNC();
    return true;
}

/* the bit reader: aligned 32-bit words of the input, least significant bit first */
struct InfBits {
    const uint32_t *w;
    uint64_t bb;
    int bc;
    uint32_t wi, wmax;
};
XM_HD void inf_bits_init(InfBits &B, const uint8_t *in, uint32_t in_len)
{
    const uintptr_t a = (uintptr_t)in;
    const uint32_t mis = (uint32_t)(a & 3u);
    B.w = (const uint32_t *)(a - mis);
    B.bb = (uint64_t)(B.w[0] >> (8u * mis));
    B.bc = 32 - 8 * (int)mis;
    B.wi = 1;
    B.wmax = (mis + in_len + 3u) / 4u + 2u;       /* a valid stream never asks for more; the buffers are padded for these words */
}
#define XM_INF_NEED(B, n)                                                        \
    do {                                                                         \
        if ((B).bc < (n)) {                                                      \
            if ((B).wi >= (B).wmax) return INF_E_INPUT;                          \
            (B).bb |= (uint64_t)(B).w[(B).wi++] << (B).bc;                       \
            (B).bc += 32;                                                        \
        }                                                                        \
    } while (0)
XM_HD uint32_t inf_take(InfBits &B, int n)
{
    const uint32_t v = (uint32_t)B.bb & ((1u << n) - 1u);
    B.bb >>= n;
    B.bc -= n;
    return v;
}

/* one symbol of a canonical code (at least 15 bits buffered): the table for codes of at most `bits` bits; a longer code is
 * found by its length -- the next 15 bits, first bit on top, lie below limit[l] for the first time at the code's length */
XM_HD int inf_symbol(InfBits &B, const uint16_t *lut, int bits, const uint16_t *limit, const int16_t *delta, const uint16_t *sym)
{
    const uint32_t e = lut[(uint32_t)B.bb & ((1u << bits) - 1u)];
    if (e) {
        const int l = (int)(e >> 9);
        B.bb >>= l;
        B.bc -= l;
        return (int)(e & 0x1ffu);
    }
    const uint32_t v = inf_bitrev((uint32_t)B.bb & 0x7fffu, 15);
    for (int l = bits + 1; l < 16; ++l) {
        if (v < (uint32_t)limit[l]) {
            B.bb >>= l;
            B.bc -= l;
            return (int)sym[(int)(v >> (15 - l)) + (int)delta[l]];
        }
    }
    return -1;
}

/*
 * Raw DEFLATE stream at `in` (in_len bytes, readable as aligned words up to 12 bytes past its end) -> exactly
 * out_len bytes at `out`.  Every lane of the warp calls it with the same arguments; returns INF_OK or an INF_E_* code
 * (the same in every lane).  base: inf_base_word(0..63).
 */
XM_HD int inflate_raw(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len, InflateTables &T, const uint32_t *base)
{
    const int lane = XM_INF_LANE;
    InfBits B;
    inf_bits_init(B, in, in_len);
    uint32_t pos = 0;
    if (lane == 0) T.fixed = 0;
    XM_INF_SYNC();
    for (;;) {
        XM_INF_NEED(B, 3);
        const uint32_t bfinal = inf_take(B, 1), btype = inf_take(B, 2);
        if (btype == 0) {
            /* stored: LEN, ~LEN at the next byte boundary, then the bytes */
            inf_take(B, B.bc & 7);
            XM_INF_NEED(B, 32);
            const uint32_t len = inf_take(B, 16), nlen = inf_take(B, 16);
            if ((len ^ nlen) != 0xffffu) return INF_E_HEADER;
            /* the byte after the header: words fetched so far minus the bytes still buffered */
            const uint8_t *src = (const uint8_t *)(B.w + B.wi) - (B.bc >> 3);
            if ((uint64_t)(src - in) + len > in_len) return INF_E_INPUT;
            if (pos + len > out_len) return INF_E_OVERRUN;
            for (uint32_t k = (uint32_t)lane; k < len; k += XM_INF_LANES) out[pos + k] = src[k];
            pos += len;
            if ((uint64_t)(src - in) + len == in_len && bfinal) break;          /* nothing may follow: do not touch what lies beyond */
            inf_bits_init(B, src + len, in_len - (uint32_t)(src + len - in));
        } else if (btype == 1 || btype == 2) {
            if (btype == 1) {
                if (!T.fixed) {
                    XM_INF_SYNC();
                    for (int s = lane; s < 288; s += XM_INF_LANES) T.lens[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
                    XM_INF_SYNC();
                    inf_build(T.lens, 288, T.lit_limit, T.lit_delta, T.lit_sym, T.lit_lut, INF_LIT_BITS);
                    XM_INF_SYNC();
                    for (int s = lane; s < 32; s += XM_INF_LANES) T.lens[s] = 5;          /* 30 and 31 complete the code and are refused when met */
                    XM_INF_SYNC();
                    inf_build(T.lens, 32, T.dist_limit, T.dist_delta, T.dist_sym, T.dist_lut, INF_DIST_BITS);
                    if (lane == 0) T.fixed = 1;
                    XM_INF_SYNC();
                }
            } else {
                XM_INF_NEED(B, 14);
                const int nlen = (int)inf_take(B, 5) + 257, ndist = (int)inf_take(B, 5) + 1, ncode = (int)inf_take(B, 4) + 4;
                if (nlen > 286 || ndist > 30) return INF_E_HEADER;
                XM_INF_SYNC();
                if (lane == 0) T.fixed = 0;
                /* the code-length code: 19 lengths of 3 bits in a fixed order; its table lives in dist_lut for a moment */
                uint32_t cl[3] = {0, 0, 0};          /* 19 x 4 bits, by symbol: every lane holds them */
                const uint8_t order_[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                for (int k = 0; k < ncode; ++k) {
                    XM_INF_NEED(B, 3);
                    const int s = order_[k];
                    cl[s / 8] |= inf_take(B, 3) << (4 * (s % 8));
                }
                for (int s = lane; s < 19; s += XM_INF_LANES) T.lens[s] = (uint8_t)((cl[s / 8] >> (4 * (s % 8))) & 7u);
                XM_INF_SYNC();
                if (!inf_build(T.lens, 19, T.dist_limit, T.dist_delta, T.dist_sym, T.dist_lut, 7)) return INF_E_CODES;
                /* the nlen + ndist code lengths, run-length coded; lane 0 writes them */
                int idx = 0;
                uint32_t prev = 0;
                while (idx < nlen + ndist) {
                    XM_INF_NEED(B, 15 + 7);
                    const int s = inf_symbol(B, T.dist_lut, 7, T.dist_limit, T.dist_delta, T.dist_sym);
                    if (s < 0) return INF_E_CODES;
                    uint32_t val = 0;
                    int rep = 1;
                    if (s < 16) { val = (uint32_t)s; prev = val; }
                    else if (s == 16) { if (idx == 0) return INF_E_CODES; val = prev; rep = 3 + (int)inf_take(B, 2); }
                    else if (s == 17) { rep = 3 + (int)inf_take(B, 3); prev = 0; }
                    else { rep = 11 + (int)inf_take(B, 7); prev = 0; }
                    if (idx + rep > nlen + ndist) return INF_E_CODES;
                    /* lens[0..18] held the code-length code's own lengths: its tables are built, every lane is past reading them */
                    if (lane == 0) for (int k = 0; k < rep; ++k) T.lens[idx + k] = (uint8_t)val;
                    idx += rep;
                }
                XM_INF_SYNC();
                if (T.lens[256] == 0) return INF_E_CODES;                    /* no end-of-block code */
                bool any = false;                                            /* a block of literals only may have no distance code */
                for (int k = 0; k < ndist; ++k) any = any || T.lens[nlen + k] != 0;
                XM_INF_SYNC();
                if (!inf_build(T.lens, nlen, T.lit_limit, T.lit_delta, T.lit_sym, T.lit_lut, INF_LIT_BITS)) return INF_E_CODES;
                XM_INF_SYNC();
                if (any) {
                    if (!inf_build(T.lens + nlen, ndist, T.dist_limit, T.dist_delta, T.dist_sym, T.dist_lut, INF_DIST_BITS)) return INF_E_CODES;
                } else {
                    if (lane == 0) for (int l = 0; l < 16; ++l) T.dist_limit[l] = 0;
                    for (int k = lane; k < (1 << INF_DIST_BITS); k += XM_INF_LANES) T.dist_lut[k] = 0;
                }
                XM_INF_SYNC();
            }
            /* The symbols.  Literals wait in a register, four at a time, and leave with one store of lanes 0..3.  A match of at
             * most 32 bytes that does not overlap itself is LOADED when it is met and STORED when the next match (or the end
             * of the block) comes: the decode of the symbols in between hides the load's round trip to L2. */
            uint32_t lit_acc = 0, lit_n = 0;
            uint32_t pend_pos = 0, pend_len = 0, pend_byte = 0;
#define XM_INF_FLUSH()                                                                                              \
    do {                                                                                                            \
        if ((uint32_t)lane < lit_n) out[pos - lit_n + (uint32_t)lane] = (uint8_t)(lit_acc >> (8u * (uint32_t)lane)); \
        lit_acc = 0; lit_n = 0;                                                                                     \
    } while (0)
#define XM_INF_COMMIT()                                                                                             \
    do {                                                                                                            \
        if ((uint32_t)lane < pend_len) out[pend_pos + (uint32_t)lane] = (uint8_t)pend_byte;                         \
        pend_len = 0;                                                                                               \
    } while (0)
            for (;;) {
                XM_INF_NEED(B, 20);
                const int s = inf_symbol(B, T.lit_lut, INF_LIT_BITS, T.lit_limit, T.lit_delta, T.lit_sym);
                if (s < 0) return INF_E_SYMBOL;
                if (s < 256) {
                    if (pos >= out_len) return INF_E_OVERRUN;
                    lit_acc |= (uint32_t)s << (8u * lit_n);
                    ++lit_n; ++pos;
                    if (lit_n == (XM_INF_LANES >= 4 ? 4u : 1u)) XM_INF_FLUSH();
                    continue;
                }
                XM_INF_FLUSH();
                if (s == 256) { XM_INF_COMMIT(); break; }
                if (s > 285) return INF_E_SYMBOL;
                const uint32_t lw = base[s - 257];
                const uint32_t len = (lw >> 8) + inf_take(B, (int)(lw & 15u));
                XM_INF_NEED(B, 28);
                const int d = inf_symbol(B, T.dist_lut, INF_DIST_BITS, T.dist_limit, T.dist_delta, T.dist_sym);
                if (d < 0 || d > 29) return INF_E_DISTANCE;
                const uint32_t dw = base[32 + d];
                const uint32_t dist = (dw >> 8) + inf_take(B, (int)(dw & 15u));
                if (dist > pos) return INF_E_DISTANCE;
                if (pos + len > out_len) return INF_E_OVERRUN;
                XM_INF_COMMIT();
                XM_INF_SYNC();                                            /* the bytes the match reads are written */
                const uint8_t *from = out + pos - dist;
                if (dist >= len) {
                    if (len <= (uint32_t)XM_INF_LANES && XM_INF_LANES > 1) {
                        if ((uint32_t)lane < len) pend_byte = from[lane];
                        pend_pos = pos; pend_len = len;
                    } else
                        for (uint32_t k = (uint32_t)lane; k < len; k += XM_INF_LANES) out[pos + k] = from[k];
                } else {
                    /* the match overlaps what it writes: the source repeats with period dist */
                    for (uint32_t k = (uint32_t)lane; k < len; k += XM_INF_LANES) out[pos + k] = from[k % dist];
                }
                pos += len;
            }
#undef XM_INF_FLUSH
#undef XM_INF_COMMIT
        } else
            return INF_E_HEADER;
        if (bfinal) break;
    }
    XM_INF_SYNC();
    return pos == out_len ? INF_OK : INF_E_SHORT;
}

/* ---- CRC-32 (reflected 0xEDB88320, as in gzip) ------------------------------------------------------------ */
XM_HD uint32_t crc_table_entry(uint32_t k)
{
    uint32_t c = k;
    for (int j = 0; j < 8; ++j) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    return c;
}
/* register after feeding n bytes to it, starting from c */
XM_HD uint32_t crc_feed(const uint32_t *table, uint32_t c, const uint8_t *p, uint32_t n)
{
    for (uint32_t k = 0; k < n; ++k) c = table[(c ^ p[k]) & 0xffu] ^ (c >> 8);
    return c;
}
/* register after n zero bytes */
XM_HD uint32_t crc_zeros(const uint32_t *table, uint32_t c, uint32_t n)
{
    for (uint32_t k = 0; k < n; ++k) c = table[c & 0xffu] ^ (c >> 8);
    return c;
}
/* CRC-32 of n bytes by 32 lanes: the data is cut into slices of L = ceil(n / 32) bytes, right-aligned (the first ones
 * shorter or empty).  Lane i feeds slice i to a register that starts at 0 -- at all ones for the first slice that holds a
 * byte -- and lane j also finds op[j] = the register after L zero bytes from 1 << j.  The register is linear, so feeding a
 * slice behind state c gives zeros_L(c) ^ slice: crc_join folds the slices in order. */
XM_HD void crc_slice(uint32_t n, int lane, uint32_t &lo, uint32_t &hi, uint32_t &L)
{
    L = (n + 31u) / 32u;
    const int64_t h = (int64_t)n - (int64_t)(31 - lane) * (int64_t)L, l = h - (int64_t)L;
    hi = h < 0 ? 0u : (uint32_t)h;
    lo = l < 0 ? 0u : (uint32_t)l;
}
XM_HD uint32_t crc_join(const uint32_t *slice_crc, const uint32_t *op, int first)
{
    uint32_t acc = slice_crc[first];
    for (int i = first + 1; i < 32; ++i) {
        uint32_t z = 0;
        for (int j = 0; j < 32; ++j) if ((acc >> j) & 1u) z ^= op[j];
        acc = z ^ slice_crc[i];
    }
    return ~acc;
}
#if !XM_DEVICE_PASS
/* the same on the host, one lane after the other (tests) */
inline uint32_t crc32_by_lanes(const uint8_t *p, uint32_t n)
{
    uint32_t table[256], slice[32], op[32];
    for (uint32_t k = 0; k < 256; ++k) table[k] = crc_table_entry(k);
    if (!n) return 0;
    int first = -1;
    for (int lane = 0; lane < 32; ++lane) {
        uint32_t lo, hi, L;
        crc_slice(n, lane, lo, hi, L);
        if (hi > lo && first < 0) first = lane;
        slice[lane] = hi > lo ? crc_feed(table, lane == first ? 0xffffffffu : 0u, p + lo, hi - lo) : 0u;
        op[lane] = crc_zeros(table, 1u << lane, L);
    }
    return crc_join(slice, op, first);
}
#endif

#if defined(__CUDACC__)
/* one BGZF block: where its DEFLATE stream lies in the compressed buffer, where its data goes */
struct BgzfDev {
    uint64_t in_off;          /* of the DEFLATE stream (behind the block's header) */
    uint64_t out_off;
    uint32_t in_len;          /* bytes of DEFLATE data */
    uint32_t out_len;         /* ISIZE */
    uint32_t crc;             /* CRC-32 from the trailer */
    uint32_t pad;
};
constexpr int INF_WARPS = 4;
#ifndef XM_INF_OCC
#define XM_INF_OCC 12
#endif

/* status[0]: atomicMin of (block << 8 | code) over the failed blocks (all ones: none) */
__global__ void __launch_bounds__(INF_WARPS * 32, XM_INF_OCC)
k_bgzf_inflate(const uint8_t *comp, const BgzfDev *blocks, uint32_t n, uint8_t *out, unsigned long long *status, int check_crc)
{
    __shared__ InflateTables s_tab[INF_WARPS];
    __shared__ uint32_t s_crc[256];
    __shared__ uint32_t s_op[INF_WARPS][32], s_slice[INF_WARPS][32];
    __shared__ uint32_t s_base[INF_BASE_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t k = threadIdx.x; k < 256; k += blockDim.x) s_crc[k] = crc_table_entry(k);
    if (threadIdx.x < INF_BASE_WORDS) s_base[threadIdx.x] = inf_base_word((int)threadIdx.x);
    __syncthreads();
    const uint32_t b = blockIdx.x * INF_WARPS + (uint32_t)warp;
    if (b >= n) return;
    const BgzfDev blk = blocks[b];
    if (!blk.out_len && blk.in_len <= 2) return;             /* the empty block at the end of the file */
    uint8_t *dst = out + blk.out_off;
    int rc = inflate_raw(comp + blk.in_off, blk.in_len, dst, blk.out_len, s_tab[warp], s_base);
    if (rc == INF_OK && check_crc && blk.out_len) {
        __syncwarp();
        uint32_t lo, hi, L;
        crc_slice(blk.out_len, lane, lo, hi, L);
        const int first = __ffs((int)__ballot_sync(0xffffffffu, hi > lo)) - 1;
        s_slice[warp][lane] = hi > lo ? crc_feed(s_crc, lane == first ? 0xffffffffu : 0u, dst + lo, hi - lo) : 0u;
        s_op[warp][lane] = crc_zeros(s_crc, 1u << lane, L);
        __syncwarp();
        if (crc_join(s_slice[warp], s_op[warp], first) != blk.crc) rc = INF_E_CRC;
    }
    if (rc != INF_OK && lane == 0) atomicMin(status, ((unsigned long long)b << 8) | (unsigned long long)rc);
}
#endif

}  // namespace xm
