/*
 * xm_walk.h -- host orchestration of one resident walk: scratch sizing, the
 * two kernel launches, the re-runs (tile-geometry overflow, compact-array
 * growth, exact-prefix re-run after a failing record) and the mapping of
 * device error words onto the reference's exception classes.
 *
 * Written against a small Backend interface so that the CUDA runtime
 * (xm_api.cu) and the CPU emulation harness of the tests (tests/emu) run the
 * same logic.  Backend:
 *     void *alloc(size_t)            scratch in the backend's memory space
 *     void  release(void *)
 *     int   zero(void *, size_t)
 *     int   fill64(void *, unsigned long long v, size_t n)
 *     int   read(void *host_dst, const void *src, size_t)     (D2H, synchronous)
 *     int   read_words(const void *const src[], int n, unsigned long long *dst)   n 8-byte words in one round trip
 *     int   write(void *dst, const void *host_src, size_t)    (H2D, synchronous)
 *     std::string last_error()
 *     int   scan(const ScanArgs &, bool small)
 *     int   classify(const ClassifyArgs &, bool small)
 *     int   sync()
 *     void  tick(int which) / float elapsed(...)               optional timing
 */
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/xenomapper_b200.h"
#include "xm_common.h"

namespace xm {

/* AS > min_score  <=>  AS >= threshold, for integer AS (xm.py:275-284 compare with > and <=) */
inline long long score_threshold(double min_score)
{
    if (min_score == -INFINITY) return -(1ll << 40);
    if (min_score == INFINITY) return 1ll << 40;
    const double f = floor(min_score);
    if (f >= 2147483647.0) return 1ll << 40;
    if (f < -2147483648.0) return -(1ll << 40);
    return (long long)f + 1;
}

/* ---- host-side diagnosis of a failing record's text ------------------------ */
inline bool host_is_space(unsigned char c) { return c == ' ' || (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x1f); }

inline bool utf8_valid(const unsigned char *s, size_t n)
{
    size_t i = 0;
    while (i < n) {
        unsigned char c = s[i];
        if (c < 0x80) { i++; continue; }
        int k;
        uint32_t v, lo;
        if (c >= 0xc2 && c < 0xe0) { k = 1; v = c & 0x1f; lo = 0x80; }
        else if (c >= 0xe0 && c < 0xf0) { k = 2; v = c & 0x0f; lo = 0x800; }
        else if (c >= 0xf0 && c < 0xf5) { k = 3; v = c & 0x07; lo = 0x10000; }
        else return false;
        for (int j = 1; j <= k; j++) {
            if (i + j >= n || (s[i + j] & 0xc0) != 0x80) return false;
            v = (v << 6) | (s[i + j] & 0x3f);
        }
        if (v < lo || v > 0x10ffff || (v >= 0xd800 && v <= 0xdfff)) return false;
        i += k + 1;
    }
    return true;
}

/* PEP 515: every '_' sits between two digits; returns false when violated */
inline bool strip_underscores(const std::string &in, std::string &out)
{
    out.clear();
    for (size_t i = 0; i < in.size(); i++) {
        if (in[i] == '_') {
            if (i == 0 || i + 1 >= in.size() || !isdigit((unsigned char)in[i - 1]) || !isdigit((unsigned char)in[i + 1])) return false;
            continue;
        }
        out.push_back(in[i]);
    }
    return true;
}
inline bool ci_equal(const char *s, const char *lit)
{
    for (; *lit; s++, lit++) if (tolower((unsigned char)*s) != *lit) return false;
    return *s == 0;
}
/* would Python's float() accept it?  (xm.py:191) */
inline bool py_float_accepts(const std::string &raw)
{
    std::string t;
    if (!strip_underscores(raw, t)) return false;
    const char *p = t.c_str();
    if (*p == '+' || *p == '-') p++;
    if (ci_equal(p, "inf") || ci_equal(p, "infinity") || ci_equal(p, "nan")) return true;
    int nd = 0;
    while (isdigit((unsigned char)*p)) { p++; nd++; }
    if (*p == '.') { p++; while (isdigit((unsigned char)*p)) { p++; nd++; } }
    if (!nd) return false;
    if (*p == 'e' || *p == 'E') {
        p++;
        if (*p == '+' || *p == '-') p++;
        if (!isdigit((unsigned char)*p)) return false;
        while (isdigit((unsigned char)*p)) p++;
    }
    return *p == 0;
}
/* would Python's int() accept it?  (xm.py:250) */
inline bool py_int_accepts(const std::string &raw)
{
    std::string t;
    if (!strip_underscores(raw, t)) return false;
    const char *p = t.c_str();
    if (*p == '+' || *p == '-') p++;
    if (!isdigit((unsigned char)*p)) return false;
    while (isdigit((unsigned char)*p)) p++;
    return *p == 0;
}

/* class of a "value is not a plain 31-bit integer" report for one line and one lookup */
inline int diagnose_number(const std::string &line, int score_src, bool want_xs)
{
    for (unsigned char c : line) if (c >= 0x80) return XM_ERR_UNSUPPORTED;
    std::vector<std::string> tok;
    size_t i = 0;
    while (i < line.size()) {
        while (i < line.size() && host_is_space((unsigned char)line[i])) i++;
        size_t j = i;
        while (j < line.size() && !host_is_space((unsigned char)line[j])) j++;
        if (j > i) tok.push_back(line.substr(i, j - i));
        i = j;
    }
    const bool nm = score_src == SCORE_CIGAR_NM && !want_xs;
    const char *tag = want_xs ? (score_src == SCORE_AS_ZS ? "ZS" : "XS") : (nm ? "NM" : "AS");
    for (size_t k = 11; k < tok.size(); k++) {
        if (tok[k].find(tag) == std::string::npos) continue;
        const size_t c = tok[k].rfind(':');
        const std::string val = c == std::string::npos ? tok[k] : tok[k].substr(c + 1);
        const bool ok = nm ? py_int_accepts(val) : py_float_accepts(val);
        return ok ? XM_ERR_UNSUPPORTED : XM_ERR_VALUE;      /* first match: duplicates were reported as such */
    }
    return XM_ERR_UNSUPPORTED;
}

/* ---- scratch --------------------------------------------------------------- */
struct Scratch {
    unsigned long long *chain1_s = nullptr, *chain1_p = nullptr, *chain2 = nullptr;
    uint64_t cap_tiles_s = 0, cap_tiles_p = 0;
    SCompact sc{nullptr, nullptr, nullptr};
    uint64_t sc_cap = 0;
    uint64_t *p_start = nullptr;       /* chunked walks: byte offset of every yielded primary record */
    uint64_t p_start_cap = 0;
    Globals *g = nullptr;
    uint32_t last_kernels = 0;       /* bit 0 k_scan2, bit 1 k_classify2, bit 2 k_scan, bit 3 k_classify, bit 4 the row kernels ran in the last resident walk */
    /* the walk over rows (xm_emit.cuh): rows of the primary stream, per-tile bin totals */
    SCompact scp{nullptr, nullptr, nullptr};
    uint64_t scp_cap = 0;
    unsigned long long *tile_tot = nullptr;
    uint64_t cap_tot = 0;
    double line_bytes[2] = {0, 0};   /* mean bytes per record of the last walk's streams: sizes the row arrays of the next one */
    int short_lines[2] = {-1, -1};   /* the span kernels' geometry per stream (primary, secondary): -1 not looked at yet, 1 spans of half the size */
};

template <class BE>
inline void scratch_release(BE &be, Scratch &s)
{
    be.release(s.chain1_s); be.release(s.chain1_p); be.release(s.chain2);
    be.release(s.sc.start); be.release(s.sc.rec); be.release(s.sc.meta);
    be.release(s.g); be.release(s.p_start);
    be.release(s.scp.start); be.release(s.scp.rec); be.release(s.scp.meta); be.release(s.tile_tot);
    s = Scratch();
}

template <class BE>
inline bool scratch_reserve(BE &be, Scratch &s, uint64_t tiles_s, uint64_t tiles_p, uint64_t sc_cap)
{
    if (!s.g) { s.g = (Globals *)be.alloc(sizeof(Globals)); if (!s.g) return false; }
    if (tiles_s > s.cap_tiles_s) {
        be.release(s.chain1_s);
        s.chain1_s = (unsigned long long *)be.alloc(tiles_s * 8);
        if (!s.chain1_s) return false;
        s.cap_tiles_s = tiles_s;
    }
    if (tiles_p > s.cap_tiles_p) {
        be.release(s.chain1_p); be.release(s.chain2);
        s.chain1_p = (unsigned long long *)be.alloc(tiles_p * 8);
        s.chain2 = (unsigned long long *)be.alloc(tiles_p * 8 * C2_SLOTS);
        if (!s.chain1_p || !s.chain2) return false;
        s.cap_tiles_p = tiles_p;
    }
    if (sc_cap > s.sc_cap) {
        be.release(s.sc.start); be.release(s.sc.rec); be.release(s.sc.meta);
        s.sc.start = (uint64_t *)be.alloc((sc_cap + 1) * 8);
        s.sc.rec = (uint4 *)be.alloc(sc_cap * 16 + 16);
        s.sc.meta = (uint32_t *)be.alloc(sc_cap * 4 + 16);
        if (!s.sc.start || !s.sc.rec || !s.sc.meta) { s.sc_cap = 0; return false; }
        s.sc_cap = sc_cap;
    }
    return true;
}

/* rows of the primary stream and tile totals of the row walk */
template <class BE>
inline bool scratch_reserve_rows(BE &be, Scratch &s, uint64_t scp_cap, uint64_t tiles)
{
    if (scp_cap > s.scp_cap) {
        be.release(s.scp.start); be.release(s.scp.rec); be.release(s.scp.meta);
        s.scp.start = (uint64_t *)be.alloc((scp_cap + 1) * 8);
        s.scp.rec = (uint4 *)be.alloc(scp_cap * 16 + 16);
        s.scp.meta = (uint32_t *)be.alloc(scp_cap * 4 + 16);
        if (!s.scp.start || !s.scp.rec || !s.scp.meta) { s.scp_cap = 0; return false; }
        s.scp_cap = scp_cap;
    }
    if (tiles > s.cap_tot) {
        be.release(s.tile_tot);
        s.tile_tot = (unsigned long long *)be.alloc(tiles * 8 * C2_SLOTS);
        if (!s.tile_tot) { s.cap_tot = 0; return false; }
        s.cap_tot = tiles;
    }
    return true;
}

/* records a stream of `len` bytes is expected to hold: from the previous walk's density, else from the first 256 KiB */
template <class BE>
inline bool estimate_records(BE &be, Scratch &sc, int k, const StreamBuf &B, uint64_t &need, double *mean_out = nullptr)
{
    double mean = sc.line_bytes[k];
    if (mean <= 0) {
        const size_t n = (size_t)(B.len < (256u << 10) ? B.len : (256u << 10));
        std::vector<char> smp(n);
        if (be.read(smp.data(), B.p, n)) return false;
        uint64_t nl = 0;
        for (size_t q = 0; q < n; q++) nl += smp[q] == '\n';
        mean = nl ? (double)n / (double)nl : (double)n;
    }
    need = (uint64_t)((double)B.len / mean * 1.05) + 4096;
    if (mean_out) *mean_out = mean;
    return true;
}

/* one line of a resident stream, fetched to the host (error paths only) */
template <class BE>
inline std::string fetch_line(BE &be, const StreamBuf &B, uint64_t off)
{
    std::string out;
    const size_t blk = 1 << 16;
    std::vector<char> buf(blk);
    while (off < B.len && out.size() < (64u << 20)) {
        const size_t n = (size_t)((B.len - off) < blk ? (B.len - off) : blk);
        be.read(buf.data(), B.p + off, n);
        const void *nl = memchr(buf.data(), '\n', n);
        if (nl) { out.append(buf.data(), (const char *)nl - buf.data()); return out; }
        out.append(buf.data(), n);
        off += n;
    }
    return out;
}
/* start of the line that precedes the one starting at off */
template <class BE>
inline uint64_t prev_line_off(BE &be, const StreamBuf &B, uint64_t off)
{
    if (off < 2) return 0;
    uint64_t end = off - 1;                 /* the newline that ends the previous line */
    const size_t blk = 1 << 16;
    std::vector<char> buf(blk);
    while (end > 0) {
        const size_t n = (size_t)(end < blk ? end : blk);
        be.read(buf.data(), B.p + (end - n), n);
        for (size_t k = n; k > 0; k--) if (buf[k - 1] == '\n') return end - n + k;
        end -= n;
    }
    return 0;
}

/* chunked walks (xm_stream.h): what the resident walk is told about its buffers and what it reports back */
struct WalkCtl {
    int halo = 0;                      /* record 0 of both buffers is the previous step's last record: context only */
    bool want_tail = false;            /* report the offsets below */
    uint64_t n_stream[2] = {0, 0};     /* records in each buffer (before its first blank line) */
    bool stopped[2] = {false, false};  /* a blank line ends the stream inside the buffer */
    uint64_t end_off[2] = {0, 0};      /* byte offset just past the last counted record of each buffer */
    uint64_t last_p = 0, last_s = 0;   /* byte offset of the last yielded record in each buffer */
};

struct WalkTimes {
    float ms_scan = 0, ms_classify = 0, ms_total = 0;
    uint32_t launches = 0;
};

/*
 * The resident walk.  P, S, out[] are in the backend's memory space.
 * Returns an xm_status; res is always filled as far as known.
 */
template <class BE>
inline int walk_resident(BE &be, Scratch &sc, const StreamBuf &P, const StreamBuf &S, const xm_opts &o,
                         uint8_t *const out[6], const uint64_t out_cap[6], uint32_t debug, xm_result *res, std::string &errmsg,
                         WalkCtl *ctl = nullptr)
{
    memset(res, 0, sizeof *res);
    res->err_stream = -1;
    if (o.mode < 0 || o.mode > 2 || o.score_src < 0 || o.score_src > 2) { errmsg = "bad mode / score_src"; return res->status = XM_ERR_ARG; }
    if (o.min_score != o.min_score) { errmsg = "min_score is NaN"; return res->status = XM_ERR_UNSUPPORTED; }
    if (P.len == 0 || S.len == 0) return res->status = XM_OK;       /* readline() == '' on the first call: nothing is yielded */

    bool small = (debug & DBG_SMALL_TILES) != 0;
    sc.last_kernels = 0;
    uint64_t limit = ~0ull;
    uint64_t sc_need = 0;
    Globals G;
    bool no_scan2 = false, no_cls2 = false, no_rows = false, rows_walk = false;
    uint64_t scp_need = 0;
    int code_first = 0;
    unsigned long long err_first = NO_ERROR;
    for (int attempt = 0; attempt < 8; ++attempt) {
        const uint64_t tile = small ? (uint64_t)CfgSmall::TILE : (uint64_t)CfgBig::TILE;
        const uint64_t nt_s = (S.len + tile - 1) / tile, nt_p = (P.len + tile - 1) / tile;
        if (nt_s > 0xffffffffull || nt_p > 0xffffffffull) { errmsg = "stream too large for one call"; return res->status = XM_ERR_ARG; }
        if (!sc_need && !estimate_records(be, sc, 1, S, sc_need)) { errmsg = "sample read failed"; return res->status = XM_ERR_CUDA; }
        for (int k = 0; k < 2; ++k)
            if (sc.short_lines[k] < 0) {
                /* first walk on this context: mean LINE length of the first 64 KiB decides the span kernels' geometry; a walk
                 * whose spans overflow switches it (below), and it stays for the walks that follow */
                const StreamBuf &B = k ? S : P;
                const size_t n = (size_t)(B.len < (64u << 10) ? B.len : (64u << 10));
                std::vector<char> smp(n);
                if (be.read(smp.data(), B.p, n)) { errmsg = "sample read failed"; return res->status = XM_ERR_CUDA; }
                uint64_t nl = 0;
                for (size_t q = 0; q < n; q++) nl += smp[q] == '\n';
                sc.short_lines[k] = (nl && (double)n / (double)nl < 300.0) ? 1 : 0;
            }
        const uint64_t nt_s2 = small ? 0 : be.scan2_tiles(S.len);               /* the barrier-free scan has its own tiling */
        const uint64_t nt_sc = nt_s > nt_s2 ? nt_s : nt_s2;
        const uint64_t nt_p2 = small ? 0 : be.scan2_tiles(P.len);               /* the row walk scans the primary stream with it too */
        const uint64_t nt_pc = nt_p > nt_p2 ? nt_p : nt_p2;
        if (!scratch_reserve(be, sc, nt_sc, nt_pc, sc_need)) { errmsg = "out of device memory for scratch"; return res->status = XM_ERR_NOMEM; }
        if (ctl && ctl->want_tail && sc.p_start_cap < sc.sc_cap) {
            be.release(sc.p_start);
            sc.p_start = (uint64_t *)be.alloc(sc.sc_cap * 8 + 16);
            if (!sc.p_start) { sc.p_start_cap = 0; errmsg = "out of device memory for scratch"; return res->status = XM_ERR_NOMEM; }
            sc.p_start_cap = sc.sc_cap;
        }
        Globals init;
        memset(&init, 0, sizeof init);
        init.err = NO_ERROR;
        init.limit_off = ~0ull;
        be.tick(0);                                  /* the step starts here: scratch init is part of it */
        if (be.write(sc.g, &init, sizeof init) || be.zero(sc.chain1_s, nt_sc * 8) || be.zero(sc.chain1_p, nt_pc * 8) ||
            be.zero(sc.chain2, nt_pc * 8 * C2_SLOTS)) { errmsg = "scratch init failed"; return res->status = XM_ERR_CUDA; }

        ScanArgs sa;
        memset(&sa, 0, sizeof sa);
        sa.S = S; sa.sc = sc.sc; sa.sc_cap = sc.sc_cap; sa.chain1 = sc.chain1_s; sa.g = sc.g; sa.ntiles = (uint32_t)nt_s;
        sa.score_src = o.score_src; sa.skip = o.skip_repeated ? 1 : 0; sa.stream_id = 1; sa.debug = debug;
        ClassifyArgs ca;
        memset(&ca, 0, sizeof ca);
        ca.P = P; ca.S = S; ca.sc = sc.sc; ca.sc_cap = sc.sc_cap; ca.chain1 = sc.chain1_p; ca.chain2 = sc.chain2;
        ca.g = sc.g; ca.ntiles = (uint32_t)nt_p; ca.mode = o.mode; ca.score_src = o.score_src; ca.skip = sa.skip;
        ca.thr = score_threshold(o.min_score); ca.enabled = o.enabled_bins & 0x3f; ca.limit = limit; ca.debug = debug;
        for (int b = 0; b < 6; ++b) { ca.out[b] = out[b]; ca.out_cap[b] = ((ca.enabled >> b) & 1u) ? out_cap[b] : 0; }
        if (ctl) { ca.halo = ctl->halo; if (ctl->want_tail) { ca.p_start = sc.p_start; ca.p_start_cap = sc.p_start_cap; } }

        /* ---- the walk over rows (xm_emit.cuh): both streams scanned into rows, then size / prefix / emit --------
         * Clean, error-free inputs only: a span or a row that needs the exact kernels raises Globals::pad and the
         * attempt starts over on the exact pair. */
        if (!small && !(debug & DBG_FORCE_GENERIC) && (debug & DBG_ROWS) && !no_rows && limit == ~0ull && be.rows_enabled()) {
            if (!scp_need && !estimate_records(be, sc, 0, P, scp_need)) { errmsg = "sample read failed"; return res->status = XM_ERR_CUDA; }
            {
                const uint64_t rows_p = scp_need > sc.scp_cap ? scp_need : sc.scp_cap;
                const uint64_t rec_bound = sc.sc_cap < rows_p ? sc.sc_cap : rows_p;         /* the walk yields min(records of both streams) */
                if (!scratch_reserve_rows(be, sc, scp_need, (rec_bound + EM_TILE - 1) / EM_TILE + 1)) { errmsg = "out of device memory for scratch"; return res->status = XM_ERR_NOMEM; }
            }
            ScanArgs s2 = sa;
            s2.sc = sc.sc; s2.sc_cap = sc.sc_cap; s2.short_lines = sc.short_lines[1] > 0 ? 1 : 0;
            ScanArgs p2 = sa;
            p2.short_lines = sc.short_lines[0] > 0 ? 1 : 0;
            p2.S = P; p2.sc = sc.scp; p2.sc_cap = sc.scp_cap; p2.chain1 = sc.chain1_p; p2.stream_id = 0;
            p2.want_same = (o.mode != MODE_SE && !sa.skip) ? 1 : 0;
            be.tick(3);
            int r2 = be.scan2(s2);
            be.tick(4);
            if (r2 == 0) r2 = be.scan2(p2);
            if (r2 > 0) { errmsg = "scan kernel launch failed: " + be.last_error(); return res->status = XM_ERR_CUDA; }
            if (r2 == 0) {
                be.tick(1);
                if (be.read(&G, sc.g, sizeof G)) { errmsg = "kernel execution failed: " + be.last_error(); return res->status = XM_ERR_CUDA; }
                res->n_launches += 2;
                if (G.pad) {
                    /* a span overflowed (or the input is not clean): spans of half the size first, then the other walks */
                    if (!s2.short_lines || !p2.short_lines) { sc.short_lines[0] = sc.short_lines[1] = 1; continue; }
                    no_rows = true; continue;
                }
                if (G.n_stream[1] > sc.sc_cap) { sc_need = G.n_stream[1] + 1; continue; }
                if (G.n_stream[0] > sc.scp_cap) { scp_need = G.n_stream[0] + 1; continue; }
                const uint64_t n = G.n_stream[0] < G.n_stream[1] ? G.n_stream[0] : G.n_stream[1];
                EmitArgs ea;
                memset(&ea, 0, sizeof ea);
                ea.P = P; ea.S = S; ea.rp = sc.scp; ea.rs = sc.sc; ea.n = n;
                ea.mode = o.mode; ea.skip = sa.skip; ea.halo = ctl ? ctl->halo : 0; ea.exact_names = (debug & DBG_EXACT_NAMES) ? 1 : 0;
                ea.thr = ca.thr; ea.enabled = ca.enabled; ea.g = sc.g; ea.tile_tot = sc.tile_tot;
                ea.ntiles = (uint32_t)((n + EM_TILE - 1) / EM_TILE);
                for (int b = 0; b < 6; ++b) { ea.out[b] = ca.out[b]; ea.out_cap[b] = ca.out_cap[b]; }
                if ((uint64_t)ea.ntiles > sc.cap_tot) { errmsg = "tile totals not sized"; return res->status = XM_ERR_ARG; }
                if (be.size(ea)) { errmsg = "size kernel launch failed: " + be.last_error(); return res->status = XM_ERR_CUDA; }
                be.tick(5);
                if (be.prefix(ea) || be.emit(ea)) { errmsg = "emit kernel launch failed: " + be.last_error(); return res->status = XM_ERR_CUDA; }
                be.tick(2);
                if (be.read(&G, sc.g, sizeof G)) { errmsg = "kernel execution failed: " + be.last_error(); return res->status = XM_ERR_CUDA; }
                res->n_launches += 3;
                if (G.pad) { no_rows = true; continue; }
                sc.last_kernels = 1u | 16u;
                res->ms_scan = be.elapsed(3, 1); res->ms_classify = be.elapsed(1, 2); res->ms_total += be.elapsed(0, 2);
                res->ms_kernel[0] = be.elapsed(3, 4); res->ms_kernel[1] = be.elapsed(4, 1); res->ms_kernel[2] = be.elapsed(1, 5); res->ms_kernel[3] = be.elapsed(5, 2);
                /* what the fused kernels leave in Globals and the code below reads */
                if (n > 0) {
                    const void *at[2] = {sc.scp.start, sc.scp.start + n};
                    unsigned long long v[2] = {0, 0};
                    be.read_words(at, 2, v);
                    G.bytes_in[0] = v[1] - v[0];
                }
                rows_walk = true;
                break;
            }
            /* r2 < 0: the backend has no row kernels */
            if (be.write(sc.g, &init, sizeof init) || be.zero(sc.chain1_s, nt_sc * 8) || be.zero(sc.chain1_p, nt_pc * 8)) { errmsg = "scratch init failed"; return res->status = XM_ERR_CUDA; }
        }
        be.tick(3);
        /* clean inputs take the barrier-free scan (xm_scan2.cuh); it says so when a span needs the exact kernel */
        bool scanned = false, have_gs = false;
        Globals Gs;
        for (int geom = 0; geom < 2 && !scanned && !small && !(debug & DBG_FORCE_GENERIC) && !no_scan2; ++geom) {
            ScanArgs s2 = sa;
            s2.ntiles = (uint32_t)be.scan2_tiles(S.len);
            s2.short_lines = sc.short_lines[1] > 0 ? 1 : 0;
            const int r2 = s2.ntiles ? be.scan2(s2) : -1;
            if (r2 > 0) { errmsg = "scan kernel launch failed: " + be.last_error(); return res->status = XM_ERR_CUDA; }
            if (r2 < 0) break;
            Globals G2;
            if (be.read(&G2, sc.g, sizeof G2)) { errmsg = "kernel execution failed: " + be.last_error(); return res->status = XM_ERR_CUDA; }
            res->n_launches += 1;
            if (!G2.pad) { scanned = true; Gs = G2; have_gs = true; sc.last_kernels |= 1u; break; }
            /* declined: once more with spans of half the size (a stretch of short lines overflows the 64 lines a span holds),
             * then the exact kernel -- for the re-runs of this call too */
            if (s2.short_lines) no_scan2 = true; else sc.short_lines[1] = 1;
            if (be.write(sc.g, &init, sizeof init) || be.zero(sc.chain1_s, nt_sc * 8)) { errmsg = "scratch init failed"; return res->status = XM_ERR_CUDA; }
            be.tick(3);
        }
        if (!scanned) {
            sc.last_kernels |= 4u;
            if (be.scan(sa, small)) { errmsg = "scan kernel launch failed: " + be.last_error(); return res->status = XM_ERR_CUDA; }
            res->n_launches += 1;
        }
        be.tick(1);
        /* and the barrier-free classify (k_classify2) when nothing but clean, error-free input has been seen so far */
        bool classified = false;
        for (int geom = 0; geom < 2 && !classified && !small && !(debug & DBG_FORCE_GENERIC) && !no_cls2 && limit == ~0ull; ++geom) {
            if (!have_gs) { if (be.read(&Gs, sc.g, sizeof Gs)) { errmsg = "kernel execution failed: " + be.last_error(); return res->status = XM_ERR_CUDA; } have_gs = true; }
            ca.short_lines = sc.short_lines[0] > 0 ? 1 : 0;
            const int r3 = be.classify2(ca);
            if (r3 > 0) { errmsg = "classify kernel launch failed: " + be.last_error(); return res->status = XM_ERR_CUDA; }
            if (r3 < 0) break;
            be.tick(2);
            if (be.read(&G, sc.g, sizeof G)) { errmsg = "kernel execution failed: " + be.last_error(); return res->status = XM_ERR_CUDA; }
            res->n_launches += 1;
            if (!G.pad) { classified = true; sc.last_kernels |= 2u; break; }
            /* back to the state the scan left; spans of half the size, then the exact kernel */
            if (ca.short_lines) no_cls2 = true; else sc.short_lines[0] = 1;
            Gs.pad = 0;
            if (be.write(sc.g, &Gs, sizeof Gs) || be.zero(sc.chain1_p, nt_pc * 8) || be.zero(sc.chain2, nt_pc * 8 * C2_SLOTS)) { errmsg = "scratch init failed"; return res->status = XM_ERR_CUDA; }
            be.tick(1);
        }
        if (!classified) {
            sc.last_kernels |= 8u;
            if (be.classify(ca, small)) { errmsg = "classify kernel launch failed: " + be.last_error(); return res->status = XM_ERR_CUDA; }
            be.tick(2);
            if (be.sync()) { errmsg = "kernel execution failed: " + be.last_error(); return res->status = XM_ERR_CUDA; }
            res->n_launches += 1;
            if (be.read(&G, sc.g, sizeof G)) { errmsg = "result read failed"; return res->status = XM_ERR_CUDA; }
        }
        res->ms_scan = be.elapsed(3, 1); res->ms_classify = be.elapsed(1, 2); res->ms_total += be.elapsed(0, 2);
        res->ms_kernel[0] = res->ms_scan; res->ms_kernel[1] = res->ms_classify;       /* the scan / classify pair */

#ifdef XM_PHASE_TIMING
        {
            static const char *nm[16] = {"stage", "mask+count", "scan+trk", "select", "parse", "heads", "front-end", "pre-look1", "look1", "decide", "bins", "items", "runs", "resolve2", "copy", "rare+end"};
            for (int k = 0; k < 2; ++k) {
                const unsigned long long nt = k ? nt_p : nt_s;
                fprintf(stderr, "[phases %s, cycles per tile, %llu tiles]", k ? "classify" : "scan", nt);
                for (int q = 0; q < (k ? 16 : 8); ++q) fprintf(stderr, " %s=%.0f", nm[q], (double)G.phase[12 * k + q] / (double)nt);
                if (k) fprintf(stderr, " | resolve2 windows/tile=%.2f polls/tile=%.1f per window: first-load=%.0f +spin=%.0f +fold=%.0f", (double)G.phase[30] / (double)nt, (double)G.phase[31] / (double)nt, (double)G.phase[32] / (double)G.phase[30], (double)G.phase[33] / (double)G.phase[30], (double)G.phase[34] / (double)G.phase[30]);
                fprintf(stderr, "\n");
            }
        }
#endif
        if (G.overflow && !small) { small = true; continue; }           /* pathologically short lines: 1 KiB tiles cannot overflow */
        if (G.n_stream[1] > sc.sc_cap) { sc_need = G.n_stream[1] + 1; continue; }
        if (G.err != NO_ERROR && limit == ~0ull) {
            /* re-run cut at the failing record: outputs and counts of everything before it, like the reference's streaming writes */
            err_first = G.err;
            code_first = (int)((G.err >> 2) & 63);
            limit = G.err >> 8;
            continue;
        }
        break;
    }

    unsigned long long n = G.n_stream[0] < G.n_stream[1] ? G.n_stream[0] : G.n_stream[1];
    if (limit < n) n = limit;
    res->n_records = n;
    if (ctl) {
        for (int k = 0; k < 2; ++k) {
            ctl->n_stream[k] = G.n_stream[k];
            ctl->end_off[k] = G.end_off[k];
            ctl->stopped[k] = G.end_off[k] < (k ? S.len : P.len);
        }
    }
    /* the few words of the row arrays the host still needs, in one round trip: where the last yielded record of each
     * stream starts (chunked walks carry from there), first and one-past-last offset of the secondary rows */
    if (n > 0) {
        const bool tail = ctl && ctl->want_tail;
        const void *at[4] = {sc.sc.start, sc.sc.start + n, sc.sc.start + (n - 1), (rows_walk ? sc.scp.start : (tail ? sc.p_start : sc.sc.start)) + (n - 1)};
        unsigned long long v[4] = {0, 0, 0, 0};
        be.read_words(at, tail ? 4 : 2, v);
        res->bytes_in[1] = v[1] - v[0];
        if (tail) { ctl->last_s = v[2]; ctl->last_p = v[3]; }
    }
    for (int k = 0; k < 2; ++k)
        if (G.n_stream[k] > 1000 && G.end_off[k] > 0) sc.line_bytes[k] = (double)G.end_off[k] / (double)G.n_stream[k];
    for (int k = 0; k < 36; ++k) res->counts[k] = G.counts[k];
    for (int b = 0; b < 6; ++b) res->out_len[b] = G.out_len[b];
    res->bytes_in[0] = G.bytes_in[0];
    for (int b = 0; b < 6; ++b)
        if (((o.enabled_bins >> b) & 1u) && G.out_len[b] > out_cap[b]) {
            errmsg = "output buffer too small for bin " + std::to_string(b);
            return res->status = XM_ERR_ARG;
        }
    if (err_first == NO_ERROR) return res->status = XM_OK;

    /* map the device report onto the reference's exception class */
    const int stream = (int)(err_first & 1), prev = (int)((err_first >> 1) & 1);
    res->err_record = limit;
    res->err_stream = stream;
    int status;
    if (code_first == EC_ASSERT) { status = XM_ERR_ASSERT; errmsg = "QNAMEs differ at record " + std::to_string(limit); }
    else if (code_first == EC_DUP) { status = XM_ERR_VALUE; errmsg = "SAM line has multiple values of a score tag at record " + std::to_string(limit); }
    else {
        /* fetch the text of the line the report points at */
        uint64_t off = 0;
        const StreamBuf &B = stream ? S : P;
        if (stream) { unsigned long long v = 0; be.read(&v, sc.sc.start + (limit - (prev ? 1 : 0)), 8); off = v; }
        else { off = G.limit_off; if (prev) off = prev_line_off(be, P, off); }
        const std::string line = fetch_line(be, B, off);
        if (code_first == EC_TEXT) {
            bool ascii = true;
            for (unsigned char c : line) if (c >= 0x80) ascii = false;
            status = (!ascii && !utf8_valid((const unsigned char *)line.data(), line.size())) ? XM_ERR_UNICODE : XM_ERR_UNSUPPORTED;
            errmsg = status == XM_ERR_UNICODE ? "invalid UTF-8 in record " : "text outside the device grammar (non-ASCII byte, lone CR or oversized line) in record ";
            errmsg += std::to_string(limit);
        } else {
            status = diagnose_number(line, o.score_src, code_first == EC_NUM_XS);
            errmsg = (status == XM_ERR_VALUE ? "score tag value is not a number at record " : "score tag value outside the device grammar (plain 31-bit integer) at record ") + std::to_string(limit);
        }
    }
    return res->status = status;
}

/*
 * Index pass of the sharded walk: scans one resident byte range that starts at a record boundary and reports how
 * many records it yields (run heads only with skip), whether a blank line ends the stream inside it, and -- for
 * each queried record index -- the byte offset at which that record starts (the offset just past the last
 * counted record for indices at or beyond the count).
 */
template <class BE>
inline int index_resident(BE &be, Scratch &sc, const StreamBuf &B, bool skip, uint32_t debug, uint32_t n_queries,
                          const uint64_t *record_index, uint64_t *byte_offset, xm_shard_info *info, std::string &errmsg)
{
    memset(info, 0, sizeof *info);
    info->stop_at = ~0ull;
    for (uint32_t q = 0; q < n_queries; ++q) byte_offset[q] = 0;
    if (!B.len) return XM_OK;
    bool small = (debug & DBG_SMALL_TILES) != 0;
    uint64_t need = n_queries ? B.len / 64 + 4096 : 0;         /* per-record arrays are only needed to answer queries */
    for (int attempt = 0; attempt < 4; ++attempt) {
        const uint64_t tile = small ? (uint64_t)CfgSmall::TILE : (uint64_t)CfgBig::TILE;
        const uint64_t nt = (B.len + tile - 1) / tile;
        if (nt > 0xffffffffull) { errmsg = "stream too large for one call"; return XM_ERR_ARG; }
        if (!scratch_reserve(be, sc, nt, 0, need)) { errmsg = "out of device memory for scratch"; return XM_ERR_NOMEM; }
        Globals init;
        memset(&init, 0, sizeof init);
        init.err = NO_ERROR;
        if (be.write(sc.g, &init, sizeof init) || be.zero(sc.chain1_s, nt * 8)) { errmsg = "scratch init failed: " + be.last_error(); return XM_ERR_CUDA; }
        ScanArgs sa;
        memset(&sa, 0, sizeof sa);
        sa.S = B; sa.sc = sc.sc; sa.sc_cap = n_queries ? sc.sc_cap : 0;
        sa.chain1 = sc.chain1_s; sa.g = sc.g; sa.ntiles = (uint32_t)nt;
        sa.skip = skip ? 1 : 0; sa.stream_id = 1; sa.debug = debug;
        if (be.scan(sa, small) || be.sync()) { errmsg = "index kernel failed: " + be.last_error(); return XM_ERR_CUDA; }
        Globals G;
        if (be.read(&G, sc.g, sizeof G)) { errmsg = "result read failed: " + be.last_error(); return XM_ERR_CUDA; }
        if (G.overflow && !small) { small = true; continue; }
        if (n_queries && G.n_stream[1] > sc.sc_cap) { need = G.n_stream[1] + 1; continue; }
        info->n_records = G.n_stream[1];
        info->first_start = 0;
        info->end_off = G.end_off[1];
        if (G.end_off[1] < B.len) info->stop_at = G.n_stream[1];
        for (uint32_t q = 0; q < n_queries; ++q) {
            if (record_index[q] < G.n_stream[1]) {
                unsigned long long v = 0;
                if (be.read(&v, sc.sc.start + record_index[q], 8)) { errmsg = "result read failed: " + be.last_error(); return XM_ERR_CUDA; }
                byte_offset[q] = v;
            } else byte_offset[q] = G.end_off[1];
        }
        return XM_OK;
    }
    errmsg = "record arrays could not be sized";
    return XM_ERR_NOMEM;
}

}  // namespace xm
