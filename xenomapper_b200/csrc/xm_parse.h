/*
 * xm_parse.h -- per-line logic of the read-binning path: byte classes, the
 * exact tokeniser, tag/score extraction, the category decision.
 *
 * Compiled twice from one source: as device code inside the sm_100a kernels
 * (xm_kernels.cu) and as host code inside the CPU emulation harness the tests
 * use to exercise the tile logic without a GPU (tests/emu/xm_emu.cpp).  The
 * host build is test scaffolding only; libxenomapper_b200.so never runs it.
 *
 * Reference semantics restated here (xenomapper/xenomapper.py, "xm.py"):
 *   tokenise       line.strip('\n').split()                      xm.py:103
 *   get_tag        substring match on tokens >= 11, value after
 *                  the last ':'                                  xm.py:186-191
 *   ZS alias       xm.py:204-205
 *   CIGAR score    xm.py:247-255
 *   decision       get_mapping_state                             xm.py:275-289
 *   pair chains    xm.py:423-448 (liberal), 521-550 (conservative)
 */
#pragma once
#include "xm_common.h"
#include <string.h>

namespace xm {

/* ---- portable bit helpers ------------------------------------------- */
XM_HD int ffs32(uint32_t x)
{
#if XM_DEVICE_PASS
    return __ffs((int)x);
#else
    return x ? __builtin_ctz(x) + 1 : 0;
#endif
}
XM_HD int clz32(uint32_t x)
{
#if XM_DEVICE_PASS
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
XM_HD int popc32(uint32_t x)
{
#if XM_DEVICE_PASS
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
XM_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
#if XM_DEVICE_PASS
    return __funnelshift_r(lo, hi, sh);
#else
    sh &= 31;
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}
XM_HD uint32_t umulhi32(uint32_t a, uint32_t b)
{
#if XM_DEVICE_PASS
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
XM_HD uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

/* ---- SIMD-in-register byte classes (four bytes per 32-bit word) ------- */
/* bit 7 of each byte set where the byte is < 0x21 or >= 0x80: every ASCII
 * whitespace, every control byte, every non-ASCII byte.  Clean SAM has only
 * '\t' and '\n' in this class. */
XM_HD uint32_t ctrl_mask(uint32_t w)
{
    uint32_t t = (w & 0x7f7f7f7fu) + 0x5f5f5f5fu;
    return (~t | w) & 0x80808080u;
}
/* bit 7 of each byte set where the byte equals the (ASCII) byte replicated in pat; exact */
XM_HD uint32_t eq_mask(uint32_t w, uint32_t pat)
{
    uint32_t x = w ^ pat;
    uint32_t t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return ~(t | x) & 0x80808080u;
}
/* gather the four bit-7 flags of a word into a nibble (byte 0 -> bit 0) */
XM_HD uint32_t pack4(uint32_t z) { return umulhi32(z, 0x02040810u) & 0xfu; }

/* 4 x u8 dot product with accumulate (IDP4A on the device) */
XM_HD uint32_t dp4a_u(uint32_t a, uint32_t b, uint32_t c)
{
#if XM_DEVICE_PASS
    return __dp4a(a, b, c);
#else
    for (int k = 0; k < 4; ++k) c += ((a >> (8 * k)) & 0xffu) * ((b >> (8 * k)) & 0xffu);
    return c;
#endif
}
/* sixteen bit-7 byte flags (four words) -> 16 mask bits, byte 0 of z0 -> bit 0 */
XM_HD uint32_t pack16(uint32_t z0, uint32_t z1, uint32_t z2, uint32_t z3)
{
    const uint32_t lo = dp4a_u(z1, 0x80402010u, dp4a_u(z0, 0x08040201u, 0u));   /* 128 * bits 0..7 */
    const uint32_t hi = dp4a_u(z3, 0x80402010u, dp4a_u(z2, 0x08040201u, 0u));   /* 128 * bits 8..15 */
    return (lo + (hi << 8)) >> 7;
}
/* the two byte-class masks of 16 staged bytes: W = byte < 0x21 or >= 0x80, T = byte == '\t' */
XM_HD void masks16(const uint4 v, uint32_t &W, uint32_t &T)
{
    W = pack16(ctrl_mask(v.x), ctrl_mask(v.y), ctrl_mask(v.z), ctrl_mask(v.w));
    T = pack16(eq_mask(v.x, 0x09090909u), eq_mask(v.y, 0x09090909u), eq_mask(v.z, 0x09090909u), eq_mask(v.w, 0x09090909u));
}
/* The same for the barrier-free kernels, whose spans give up (exact kernel) when two W bytes touch: the tab test
 * is the three-operation zero-byte test on v ^ '\t'.  Its only false positive is a backspace (0x08) in the byte
 * after a tab inside the same word -- two touching W bytes, so that span never uses the mask. */
XM_HD uint32_t tab_mask_loose(uint32_t w)
{
    const uint32_t x = w ^ 0x09090909u;
    return (x - 0x01010101u) & ~x & 0x80808080u;
}
XM_HD void masks16_span(const uint4 v, uint32_t &W, uint32_t &T)
{
    W = pack16(ctrl_mask(v.x), ctrl_mask(v.y), ctrl_mask(v.z), ctrl_mask(v.w));
    T = pack16(tab_mask_loose(v.x), tab_mask_loose(v.y), tab_mask_loose(v.z), tab_mask_loose(v.w));
}
/* exact newline mask of 16 staged bytes (tiles whose W & ~T bytes are not all '\n') */
XM_HD uint32_t newlines16(const uint4 v)
{
    return pack16(eq_mask(v.x, 0x0a0a0a0au), eq_mask(v.y, 0x0a0a0a0au), eq_mask(v.z, 0x0a0a0a0au), eq_mask(v.w, 0x0a0a0a0au));
}

/* The per-line parse reads a line's bytes straight from global memory, every lane of a warp at an address of
 * its own: such a load occupies the SM's L1 pipeline once per 128-byte line it touches whatever its width, and
 * that pipeline is busier than HBM in the barrier-free kernels (profiles/r01_parse16_and_whatifs.md).  So the parse
 * fetches aligned 16-byte blocks and takes its words and bytes from registers. */
/* aligned 16-byte block; every buffer the parse looks at is readable up to the next multiple of 16 */
XM_HD uint4 ld16(const uint8_t *p)
{
#if XM_DEVICE_PASS
    return *(const uint4 *)p;
#else
    uint4 v;
    memcpy(&v, p, 16);
    return v;
#endif
}
/* byte b (0..15) of a block */
XM_HD uint32_t byte_of16(const uint4 &v, int b)
{
    const uint32_t w = b < 8 ? (b < 4 ? v.x : v.y) : (b < 12 ? v.z : v.w);
    return (w >> (8 * (b & 3))) & 0xffu;
}

XM_HD bool is_w_byte(uint32_t c) { return c < 0x21u || c >= 0x80u; }

XM_HD bool is_ascii_space(uint8_t c) { return c == ' ' || (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x1f); }

/* ---- QNAME hash: 2 x 32-bit lanes over little-endian words ------------ */
struct Hash2 {
    uint32_t a, b;
};
XM_HD void hash_init(Hash2 &h) { h.a = 0x243f6a88u; h.b = 0x85a308d3u; }
XM_HD void hash_word(Hash2 &h, uint32_t w)
{
    /* two multiply-add lanes with unrelated odd multipliers; fmix32 in hash_final spreads the bits */
    h.a = h.a * 0x9e3779b1u + w;
    h.b = (h.b ^ w) * 0x85ebca6bu;
}
XM_HD uint32_t fmix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x;
}
XM_HD void hash_final(Hash2 &h, uint32_t len)
{
#ifdef XM_WEAK_HASH
    /* test builds: every QNAME of one length collides, so only the byte compare behind the hash can tell names apart */
    h.a = len; h.b = ~len;
#else
    h.a = fmix32(h.a ^ len);
    h.b = fmix32(h.b ^ (len * 0x9e3779b1u));
#endif
}

/* ---- decimal integers: the device grammar [+-]?[0-9]+, |v| < 2^31 ------ */
struct NumSt {
    uint32_t v;
    uint32_t st;    /* 0 start, 1 after sign, 2 in digits; bit 8 bad; bit 9 negative */
};
XM_HD void num_reset(NumSt &n) { n.v = 0; n.st = 0; }
XM_HD void num_feed(NumSt &n, uint32_t c)
{
    uint32_t d = c - '0';
    if (d <= 9u) {
        if (n.v > 214748364u || (n.v == 214748364u && d > 7u)) n.st |= 0x100;
        else n.v = n.v * 10u + d;
        n.st = (n.st & ~3u) | 2u;
    } else if ((n.st & 3u) == 0 && (c == '+' || c == '-')) {
        n.st |= 1u | (c == '-' ? 0x200u : 0u);
    } else {
        n.st |= 0x100;
    }
}
XM_HD bool num_ok(const NumSt &n, int32_t &out)
{
    if ((n.st & 0x100) || (n.st & 3u) != 2u) return false;
    out = (n.st & 0x200) ? -(int32_t)n.v : (int32_t)n.v;
    return true;
}

/* ---- CIGAR walk, xm.py:251-255 ----------------------------------------- */
struct CigSt {
    uint32_t run, nd_big;      /* nd_big: bit 0 digits pending, bit 1 run too large, bit 2 unsupported seen */
    uint32_t n_id;
    unsigned long long sum_id, sum_s;
};
XM_HD void cig_reset(CigSt &c) { c.run = 0; c.nd_big = 0; c.n_id = 0; c.sum_id = 0; c.sum_s = 0; }
XM_HD void cig_feed(CigSt &c, uint32_t ch)
{
    uint32_t d = ch - '0';
    if (d <= 9u) {
        if (c.run > 99999999u) c.nd_big |= 2;
        else c.run = c.run * 10u + d;
        c.nd_big |= 1;
    } else {
        if (c.nd_big & 1) {
            if (ch == 'I' || ch == 'D') { if (c.nd_big & 2) c.nd_big |= 4; c.n_id++; c.sum_id += c.run; }
            else if (ch == 'S') { if (c.nd_big & 2) c.nd_big |= 4; c.sum_s += c.run; }
        }
        c.run = 0;
        c.nd_big &= 4;
    }
}
XM_HD bool cig_score(const CigSt &c, int32_t mm, int32_t &out)
{
    if (c.nd_big & 4) return false;
    long long v = -6ll * mm - 5ll * (long long)c.n_id - 3ll * (long long)c.sum_id - 2ll * (long long)c.sum_s;
    if (v > 2147483647ll || v < -2147483647ll) return false;
    out = (int32_t)v;
    return true;
}

/* ---- one parsed line --------------------------------------------------- */
struct LineRec {
    uint32_t s;         /* start, relative to the window origin g0 */
    uint32_t rawbytes;  /* bytes the line occupies in the input, terminator included */
    uint32_t outlen;    /* bytes it occupies in an output: tokens joined by tabs + '\n' */
    uint32_t qs, qlen;  /* QNAME (token 0), start relative to g0 */
    uint32_t h1, h2;    /* QNAME hash */
    uint32_t flags;
    int32_t as, xs;
};

/* byte source for the exact path: the staged window where it covers the
 * address, global memory elsewhere (long lines, secondary-stream lines) */
struct Reader {
    const uint8_t *win;
    const uint8_t *glob;
    uint64_t g0;
    uint32_t wbytes;
    uint64_t len;
    XM_HD uint8_t at(uint64_t p) const
    {
        uint64_t r = p - g0;
        return r < wbytes ? win[r] : glob[p];
    }
};

/*
 * Exact tokeniser + tag extraction for one line starting at global offset gs.
 * Handles every input the reference accepts as ASCII text: any whitespace as
 * separator (runs collapse, leading/trailing dropped), CRLF, a missing final
 * newline, lines of any length.  One pass, O(1) state.
 */
XM_COLD void generic_parse(const Reader &rd, uint64_t gs, int score_src, LineRec &L)
{
    const bool cigar = score_src == SCORE_CIGAR_NM;
    const uint32_t xsch = score_src == SCORE_AS_ZS ? 'Z' : 'X';
    uint32_t flags = 0, ntok = 0, toklen = 0, prevc = 0, tokm = 0;   /* tokm: bit0 AS, bit1 XS, bit2 NM seen in this token */
    uint32_t as_cnt = 0, xs_cnt = 0, nm_cnt = 0, as_ok = 0, xs_ok = 0, nm_ok = 0;
    int32_t as_v = 0, xs_v = 0, nm_v = 0;
    bool in_tok = false, nontab = false, eof = false;
    NumSt num; num_reset(num);
    CigSt cig; cig_reset(cig);
    Hash2 h; hash_init(h);
    uint32_t wacc = 0, qlen = 0;
    uint64_t qs = gs;
    uint64_t p = gs;
    for (;; ++p) {
        uint32_t c = 0;
        bool end = false, ws;
        if (p >= rd.len) { end = true; eof = true; }
        else { c = rd.at(p); end = (c == '\n'); }
        if (end) ws = true;
        else if (c == '\r') {
            ws = true;
            if (!(p + 1 < rd.len && rd.at(p + 1) == '\n')) flags |= F_TEXT;   /* lone CR is a line break for the reference */
        } else if (c >= 0x80) { ws = false; flags |= F_TEXT; }
        else ws = is_ascii_space((uint8_t)c);
        if (ws) {
            if (!end && c != '\t') nontab = true;
            if (in_tok) {
                in_tok = false;
                if (ntok == 1) { if (qlen & 3) hash_word(h, wacc); }
                if (ntok > 11) {
                    if (!cigar && (tokm & 1)) { if (++as_cnt == 1) as_ok = num_ok(num, as_v); }
                    if (tokm & 2) { if (++xs_cnt == 1) xs_ok = num_ok(num, xs_v); }
                    if (cigar && (tokm & 4)) { if (++nm_cnt == 1) nm_ok = num_ok(num, nm_v); }
                }
            }
            if (end) break;
        } else {
            if (!in_tok) {
                in_tok = true; ++ntok; prevc = 0; tokm = 0; num_reset(num);
                if (ntok == 1) qs = p;
            }
            ++toklen;
            if (ntok == 1) {
                wacc |= c << (8 * (qlen & 3));
                if ((++qlen & 3) == 0) { hash_word(h, wacc); wacc = 0; }
            } else if (ntok == 6) {
                if (cigar) cig_feed(cig, c);
            } else if (ntok > 11) {
                if (c == 'S') { if (prevc == 'A') tokm |= 1; if (prevc == xsch) tokm |= 2; }
                else if (c == 'M' && prevc == 'N') tokm |= 4;
                if (c == ':') num_reset(num); else num_feed(num, c);
            }
            prevc = c;
        }
    }
    hash_final(h, qlen);
    uint64_t raw = (p - gs) + (eof ? 0 : 1);
    uint64_t outlen = ntok ? (uint64_t)toklen + ntok : 0;
    if (ntok == 0) flags |= F_BLANK;
    if (nontab || eof || (p - gs) + 1 != outlen) flags |= F_DIRTY;
    if (outlen > META_LEN_MASK || raw > 0xffffffffull) { flags |= F_TEXT; outlen &= META_LEN_MASK; }
    int32_t as = SCORE_ABSENT, xs = SCORE_ABSENT;
    if (cigar) {
        if (nm_cnt) {
            if (!nm_ok || !cig_score(cig, nm_v, as)) { flags |= F_AS_NUM; as = SCORE_ABSENT; }
        }
    } else if (as_cnt == 1) { if (as_ok) as = as_v; else flags |= F_AS_NUM; }
    else if (as_cnt > 1) flags |= F_AS_DUP;
    if (xs_cnt == 1) { if (xs_ok) xs = xs_v; else flags |= F_XS_NUM; }
    else if (xs_cnt > 1) flags |= F_XS_DUP;
    L.s = (uint32_t)(gs - rd.g0);
    L.rawbytes = (uint32_t)raw;
    L.outlen = (uint32_t)outlen;
    L.qs = (uint32_t)(qs - rd.g0);
    L.qlen = qlen;
    L.h1 = h.a; L.h2 = h.b;
    L.flags = flags;
    L.as = as; L.xs = xs;
}

/* masks of the staged window, one bit per byte */
struct WinMasks {
    const uint8_t *win;
    const uint32_t *tbm;   /* T: byte == '\t' */
    const uint32_t *nlm;   /* N: line terminator candidates, W & ~T (W = byte < 0x21 or >= 0x80) */
    const uint16_t *trk;   /* tabs in the window before each mask word */
    uint32_t wbytes;       /* staged data bytes */
    bool adj;              /* somewhere in the window a W byte directly follows another W byte */
    XM_HD uint32_t wsm(int w) const { return tbm[w] | nlm[w]; }
};

XM_HD uint32_t tabs_before(const WinMasks &M, int p)
{
    const int w = p >> 5;
    return (uint32_t)M.trk[w] + (uint32_t)popc32(M.tbm[w] & ((1u << (p & 31)) - 1u));
}
/* window position of the tab with window rank r, known to lie in mask words [wlo, whi] */
XM_HD int tab_select(const WinMasks &M, uint32_t r, int wlo, int whi)
{
    while (whi > wlo) {                 /* largest word whose rank is <= r */
        const int mid = (wlo + whi + 1) >> 1;
        if ((uint32_t)M.trk[mid] <= r) wlo = mid; else whi = mid - 1;
    }
    uint32_t bits = M.tbm[wlo];
    for (uint32_t k = r - (uint32_t)M.trk[wlo]; k; --k) bits &= bits - 1;
    return (wlo << 5) + ffs32(bits) - 1;
}
/* start of the token that holds position q: the byte after the nearest separator below q */
XM_HD int token_start(const WinMasks &M, int q)
{
    int w = q >> 5;
    uint32_t m = M.wsm(w) & ((1u << (q & 31)) - 1u);
    while (!m) { --w; m = M.wsm(w); }
    return (w << 5) + 32 - clz32(m);
}
/* first separator bit at or above q (the line's terminator bounds the search) */
XM_HD int token_end(const WinMasks &M, int q)
{
    int w = q >> 5;
    uint32_t m = M.wsm(w) & (0xffffffffu << (q & 31));
    while (!m) { ++w; m = M.wsm(w); }
    return (w << 5) + ffs32(m) - 1;
}
/* plain integer after the last ':' of the token [ts, te) */
XM_HD bool token_value(const uint8_t *win, int ts, int te, int32_t &out)
{
    /* the usual shape, ':' [+-] one to nine digits at the token's end, read backwards from the block that holds the
     * token's last byte (and the block before it when the value straddles); anything else takes the loop below */
    if (te > ts) {
        int p = te - 1, b = p & 15;
        const uint8_t *blk = win + (p & ~15);
        uint4 v = ld16(blk);
        uint32_t val = 0, mul = 1, nd = 0, c = 0;
        bool fast = true;
        for (;;) {
            c = byte_of16(v, b);
            const uint32_t d = c - '0';
            if (d > 9u) break;
            if (nd == 9u || p == ts) { fast = false; break; }
            val += d * mul; mul *= 10u; ++nd;
            --p;
            if (--b < 0) { blk -= 16; v = ld16(blk); b = 15; }
        }
        if (fast && nd) {
            bool neg = false;
            if ((c == '-' || c == '+') && p > ts) {
                neg = c == '-';
                --p;
                if (--b < 0) { blk -= 16; v = ld16(blk); b = 15; }
                c = byte_of16(v, b);
            }
            if (c == ':') { out = neg ? -(int32_t)val : (int32_t)val; return true; }
        }
    }
    int vs = te;
    while (vs > ts && win[vs - 1] != ':') --vs;
    NumSt n; num_reset(n);
    for (int p = vs; p < te; ++p) num_feed(n, win[p]);
    return num_ok(n, out);
}

/*
 * Fast path for the line [s, e) whose terminator candidate sits at window
 * position e (the first N bit at or after s, taken from the line index).
 * By construction every separator inside [s, e) is a tab; the line is clean --
 * its output equals its raw bytes -- when in addition the terminator is a real
 * '\n' inside the staged data and no two W bytes touch anywhere in [s-1, e]
 * (no empty token, no leading or trailing tab, not a blank line).
 *
 * fast_head does those checks and the QNAME (length, hash); it returns false
 * (L untouched) whenever the line needs the exact byte-wise path.  fast_tail
 * finishes the line: the aux tokens and the scores.  They are separate so that
 * a run-skipping walk can rank its records between the two.
 */
struct FastCtx {
    uint32_t r0, ntab;     /* tabs before the line's first byte, tabs inside the line */
};

XM_HD bool fast_head(const WinMasks &M, int s, int e, LineRec &L, FastCtx &fc)
{
    const uint8_t *win = M.win;
    if ((uint32_t)e >= M.wbytes || e <= s || win[e] != '\n') return false;
    if (s == 0 && (M.wsm(0) & 1u)) return false;            /* the stream opens with a separator */
    const int w0 = s >> 5, w1 = e >> 5;
    if (M.adj) {
        for (int w = w0; w <= w1; ++w) {
            const uint32_t x = M.wsm(w);
            uint32_t a = x & ((x << 1) | (w ? M.wsm(w - 1) >> 31 : 0u));
            if (w == w0) a &= 0xffffffffu << (s & 31);
            if (w == w1) a &= 0xffffffffu >> (31 - (e & 31));
            if (a) return false;
        }
    }
    const uint32_t outlen = (uint32_t)(e - s) + 1u;
    if (outlen > META_LEN_MASK) return false;
    fc.r0 = tabs_before(M, s);
    fc.ntab = tabs_before(M, e) - fc.r0;
    /* QNAME = [s, first tab) */
    int sep0 = e;
    if (fc.ntab) {
        int w = w0;
        uint32_t m = M.tbm[w] & (0xffffffffu << (s & 31));
        while (!m) m = M.tbm[++w];
        sep0 = (w << 5) + ffs32(m) - 1;
    }
    const int qlen = sep0 - s;
    Hash2 h; hash_init(h);
    {
        const int base = s >> 2;
        const uint32_t sh = (uint32_t)(s & 3) * 8u;
        const int nw = (qlen + 3) >> 2;
        /* the words base .. base + nw, fetched a block at a time: word j of the name's words closes hash step j - 1 */
        const uint32_t lastmask = (qlen & 3) ? (1u << (8 * (qlen & 3))) - 1u : 0xffffffffu;
        uint32_t lo = 0;
        int j = ((base >> 2) << 2) - base;              /* index of the block's first word, relative to base */
        auto step = [&](uint32_t wd) {
            const int k = j - 1;
            if (k >= 0 && k < nw) {
                uint32_t w = funnel_r(lo, wd, sh);
                if (k == nw - 1) w &= lastmask;
                hash_word(h, w);
            }
            lo = wd; ++j;
        };
#if defined(XM_WHATIF) && XM_WHATIF == 4
        if (false)
#endif
        for (int qb = base >> 2; qb <= (base + nw) >> 2; ++qb) {
            const uint4 v = ld16(win + (qb << 4));
            step(v.x); step(v.y); step(v.z); step(v.w);
        }
    }
    hash_final(h, (uint32_t)qlen);
    L.s = (uint32_t)s; L.rawbytes = outlen; L.outlen = outlen;
    L.qs = (uint32_t)s; L.qlen = (uint32_t)qlen;
    L.h1 = h.a; L.h2 = h.b; L.flags = 0; L.as = SCORE_ABSENT; L.xs = SCORE_ABSENT;
    return true;
}

XM_HD void fast_tail(const WinMasks &M, int s, int e, int score_src, const FastCtx &fc, LineRec &L)
{
    if (fc.ntab < 11) return;           /* no token with index >= 11: both scores absent (xm.py:186-188) */
#if defined(XM_WHATIF) && XM_WHATIF == 3
    return;
#endif
    const uint8_t *win = M.win;
    const int w0 = s >> 5, w1 = e >> 5;
    /* aux tokens (index >= 11): look for the tag letters a 16-byte block at a time */
    uint32_t flags = 0;
    int32_t as = SCORE_ABSENT, xs = SCORE_ABSENT;
    const bool cigar = score_src == SCORE_CIGAR_NM;
    const uint32_t xsch = score_src == SCORE_AS_ZS ? 'Z' : 'X';
    const int a = tab_select(M, fc.r0 + 10, w0, w1) + 1;
    int as_ts = -1, xs_ts = -1, nm_ts = -1;
    uint32_t as_cnt = 0, xs_cnt = 0;
    const int k0 = a >> 4, k1 = (e - 1) >> 4;
    uint32_t prev_last = 0;                                   /* the byte before the block */
    for (int k = k0; k <= k1; ++k) {
        const uint4 v = ld16(win + (k << 4));
        uint32_t z0 = eq_mask(v.x, 0x53535353u), z1 = eq_mask(v.y, 0x53535353u), z2 = eq_mask(v.z, 0x53535353u), z3 = eq_mask(v.w, 0x53535353u);   /* 'S' */
        if (cigar) { z0 |= eq_mask(v.x, 0x4d4d4d4du); z1 |= eq_mask(v.y, 0x4d4d4d4du); z2 |= eq_mask(v.z, 0x4d4d4d4du); z3 |= eq_mask(v.w, 0x4d4d4d4du); }   /* 'M' */
        uint32_t z = pack16(z0, z1, z2, z3);                  /* one bit per byte */
        if (k == k0) z &= 0xffffu << (a & 15);
        if (k == k1) z &= 0xffffu >> (15 - ((e - 1) & 15));
        const uint32_t carry = prev_last;
        prev_last = v.w >> 24;
        while (z) {
            const int b = ffs32(z) - 1;
            const int p = (k << 4) + b;
            z &= z - 1;
            if (p <= a) continue;
            const uint32_t c1 = byte_of16(v, b), c0 = b ? byte_of16(v, b - 1) : carry;
            if (c1 == 'S') {
                if (c0 == 'A' && !cigar) {
                    const int ts = token_start(M, p - 1);
                    if (as_cnt == 0) { as_ts = ts; as_cnt = 1; } else if (ts != as_ts) as_cnt = 2;
                }
                if (c0 == xsch) {
                    const int ts = token_start(M, p - 1);
                    if (xs_cnt == 0) { xs_ts = ts; xs_cnt = 1; } else if (ts != xs_ts) xs_cnt = 2;
                }
            } else if (c0 == 'N' && nm_ts < 0) {
                nm_ts = token_start(M, p - 1);
            }
        }
    }
    if (cigar) {
        if (nm_ts >= 0) {
            int32_t mm;
            const bool ok = token_value(win, nm_ts, token_end(M, nm_ts), mm);
            const int sep4 = tab_select(M, fc.r0 + 4, w0, w1), sep5 = tab_select(M, fc.r0 + 5, w0, w1);
            CigSt cg; cig_reset(cg);
            for (int p = sep4 + 1; p < sep5; ++p) cig_feed(cg, win[p]);
            if (!ok || !cig_score(cg, mm, as)) { flags |= F_AS_NUM; as = SCORE_ABSENT; }
        }
    } else if (as_cnt == 1) {
        if (!token_value(win, as_ts, token_end(M, as_ts), as)) { flags |= F_AS_NUM; as = SCORE_ABSENT; }
    } else if (as_cnt > 1) flags |= F_AS_DUP;
    if (xs_cnt == 1) {
        if (!token_value(win, xs_ts, token_end(M, xs_ts), xs)) { flags |= F_XS_NUM; xs = SCORE_ABSENT; }
    } else if (xs_cnt > 1) flags |= F_XS_DUP;
    L.flags = flags; L.as = as; L.xs = xs;
}

/* exact QNAME comparison of two tokens given by global offsets */
XM_HD bool names_equal(const Reader &rd, uint64_t a, uint32_t alen, uint64_t b, uint32_t blen)
{
    if (alen != blen) return false;
    const uint64_t ra = a - rd.g0, rb = b - rd.g0;
    if (ra + alen + 4 <= rd.wbytes && rb + blen + 4 <= rd.wbytes) {
        /* both inside the window: compare word-wise through aligned loads */
        const uint32_t *w32 = (const uint32_t *)rd.win;
        const int ba = (int)(ra >> 2), bb = (int)(rb >> 2);
        const uint32_t sa = (uint32_t)(ra & 3) * 8u, sb = (uint32_t)(rb & 3) * 8u;
        const int nw = (int)((alen + 3) >> 2);
        for (int k = 0; k < nw; ++k) {
            uint32_t x = funnel_r(w32[ba + k], w32[ba + k + 1], sa) ^ funnel_r(w32[bb + k], w32[bb + k + 1], sb);
            if (k == nw - 1 && (alen & 3)) x &= (1u << (8 * (alen & 3))) - 1u;
            if (x) return false;
        }
        return true;
    }
    for (uint32_t k = 0; k < alen; ++k)
        if (rd.at(a + k) != rd.at(b + k)) return false;
    return true;
}

/* ---- the category decision, xm.py:275-289, on integer scores ----------- */
/* SCORE_ABSENT plays -inf; `thr` encodes min_score: AS > min_score <=> AS >= thr. */
XM_HD int mapping_state(int32_t as1, int32_t xs1, int32_t as2, int32_t xs2, long long thr)
{
    const bool g1 = as1 != SCORE_ABSENT && (long long)as1 >= thr;
    const bool g2 = as2 != SCORE_ABSENT && (long long)as2 >= thr;
    if (!g1 && !g2) return UA;
    if (g1 && (!g2 || as1 > as2)) return (xs1 == 0 || as1 > xs1) ? PS : PM;
    if (as1 == as2) return UR;
    return (xs2 == 0 || as2 > xs2) ? SS : SM;
}

/* liberal chain xm.py:423-448: PS > SS > PM > SM > UR > UA */
XM_HD int pair_bin_liberal(int f, int r)
{
    /* priority rank of each state; the pair takes the better ranked of the two */
    const uint32_t rank = (0u << (4 * PS)) | (1u << (4 * SS)) | (2u << (4 * PM)) | (3u << (4 * SM)) | (4u << (4 * UR)) | (5u << (4 * UA));
    uint32_t rf = (rank >> (4 * f)) & 15u, rr = (rank >> (4 * r)) & 15u;
    return rf <= rr ? f : r;
}
/* conservative chain xm.py:521-550 */
XM_HD int pair_bin_conservative(int f, int r)
{
    if (f == UA || r == UA) return UA;
    const bool fp = (f == PS || f == PM), fs = (f == SS || f == SM);
    const bool rp = (r == PS || r == PM), rs = (r == SS || r == SM);
    if (f == UR || r == UR || (fp && rs) || (fs && rp)) return UR;
    return pair_bin_liberal(f, r);
}

}  // namespace xm
