/*
 * xm_shard.h -- the walk across the GPUs of one box, one process per GPU.
 *
 * The reference is one sequential pass (xm.py:95-118, 291-556), but record i
 * depends only on record i-1 (pair predicate xm.py:402, run-skipping reader
 * xm.py:110-114) and on where the walk stops (xm.py:105), so it shards by
 * RECORD INDEX of the yielded sequence.  Every rank holds a BYTE shard of each
 * stream -- bytes [len*r/W, len*(r+1)/W) of the record region, cut anywhere --
 * and nobody reads a byte twice to find the record boundaries:
 *
 *   A. line heads: the bytes up to a shard's first newline belong to the line
 *      the rank before it ends with; they are sent there (a few hundred bytes);
 *   B. context lines: walks that look at the previous record (pairs, run
 *      skipping) get the last line of the rank before in front of their shard;
 *   C. both shards are scanned into rows by k_scan2 (the work a single GPU does
 *      anyway): record counts per rank and stream fall out of it;
 *   D. all-gather of the counts: every rank knows the record index at which
 *      each byte shard starts in each stream, and N = min(N1, N2).  Rank r
 *      walks the records its PRIMARY shard holds, [c_r, c_r+1);
 *   E. the secondary rows (and the text behind them, for the bins that emit
 *      secondary lines) of those records sit in the rank's own secondary shard
 *      except for slivers at both ends, which the neighbours send: rows and
 *      text of a few hundred records.  The record before c_r comes along as
 *      context for the first pair unit;
 *   F. k_size / k_prefix / k_emit over the rank's records (xm_emit.cuh);
 *   G. all-gather of status, counts, histogram and the six bin lengths: the
 *      histogram is summed, the bins are concatenated in rank order.
 *
 * Written against the Backend of xm_walk.h plus a small Comm interface
 *     int rank(), size()
 *     int all_gather(const void *send, void *recv, size_t bytes)       host memory, `bytes` per rank
 *     int exchange(const Xfer *sends, int ns, const Xfer *recvs, int nr) pointers in the backend's memory space
 *     std::string last_error()
 * so that the NCCL runtime (xm_api.cu) and the CPU harness of the tests
 * (tests/emu, gloo) run the same logic.  Clean inputs only (the row kernels);
 * anything else is reported as XM_SHARD_DECLINED by every rank together and the
 * caller gathers the shards on one rank for the exact walk.
 */
#pragma once
#include <stdlib.h>

#include <algorithm>
#include <chrono>

#include "xm_walk.h"

namespace xm {

struct ShardBuf {
    uint8_t *p;                 /* first byte of this rank's byte shard, in the backend's memory space */
    uint64_t len;
    uint64_t front_room;        /* writable bytes before p ... */
    uint64_t back_room;         /* ... and after p + len (received line heads, slivers, 16 bytes of read slack) */
};

struct Xfer {
    int peer;
    void *ptr;
    uint64_t bytes;
};

constexpr int XM_SHARD_DECLINED = -2;       /* some rank's input needs the exact kernels: no rank has written anything */
constexpr uint64_t ROW_MARGIN_MIN = 1u << 16;

/* rows of one stream with room for slivers in front of row 0 and behind the last one */
struct ShardRows {
    SCompact alloc{nullptr, nullptr, nullptr};     /* what was allocated */
    SCompact rows{nullptr, nullptr, nullptr};      /* alloc + margin: row 0 of the rank's own scan */
    uint64_t cap = 0, margin = 0;
    unsigned long long *chain1 = nullptr;
    uint64_t cap_tiles = 0;
};
struct ShardScratch {
    int short_lines = 0;             /* the scans take spans of half the size (set when a span overflowed) */
    ShardRows r[2];
    Globals *g = nullptr;
    unsigned long long *tile_tot = nullptr;
    uint64_t cap_tot = 0;
    double line_bytes[2] = {0, 0};
};

template <class BE>
inline void shard_release(BE &be, ShardScratch &s)
{
    for (auto &r : s.r) { be.release(r.alloc.start); be.release(r.alloc.rec); be.release(r.alloc.meta); be.release(r.chain1); }
    be.release(s.g); be.release(s.tile_tot);
    s = ShardScratch();
}

template <class BE>
inline bool shard_reserve_rows(BE &be, ShardRows &r, uint64_t cap, uint64_t tiles)
{
    if (cap > r.cap) {
        be.release(r.alloc.start); be.release(r.alloc.rec); be.release(r.alloc.meta);
        const uint64_t margin = std::max<uint64_t>(ROW_MARGIN_MIN, cap / 16);
        const uint64_t total = cap + 2 * margin;
        r.alloc.start = (uint64_t *)be.alloc((total + 1) * 8);
        r.alloc.rec = (uint4 *)be.alloc(total * 16 + 16);
        r.alloc.meta = (uint32_t *)be.alloc(total * 4 + 16);
        if (!r.alloc.start || !r.alloc.rec || !r.alloc.meta) { r.cap = 0; return false; }
        r.rows.start = r.alloc.start + margin; r.rows.rec = r.alloc.rec + margin; r.rows.meta = r.alloc.meta + margin;
        r.cap = cap; r.margin = margin;
    }
    if (tiles > r.cap_tiles) {
        be.release(r.chain1);
        r.chain1 = (unsigned long long *)be.alloc(tiles * 8);
        if (!r.chain1) { r.cap_tiles = 0; return false; }
        r.cap_tiles = tiles;
    }
    return true;
}

/* offset just past the first '\n' of [p, p + len), len if there is none */
template <class BE>
inline uint64_t first_line_end(BE &be, const uint8_t *p, uint64_t len)
{
    std::vector<char> buf(1 << 16);
    for (uint64_t off = 0; off < len;) {
        const size_t n = (size_t)std::min<uint64_t>(len - off, off ? buf.size() : 4096);
        if (be.read(buf.data(), p + off, n)) return len;
        const void *nl = memchr(buf.data(), '\n', n);
        if (nl) return off + (uint64_t)((const char *)nl - buf.data()) + 1;
        off += n;
    }
    return len;
}
/* offset at which the last line of [p, p + len) starts (the line may lack its newline at the end of a stream) */
template <class BE>
inline uint64_t last_line_start(BE &be, const uint8_t *p, uint64_t len)
{
    if (!len) return 0;
    std::vector<char> buf(1 << 16);
    uint64_t end = len;
    {
        char c = 0;
        be.read(&c, p + len - 1, 1);
        if (c == '\n') end = len - 1;          /* look for the newline before the closing one */
    }
    size_t blk = 4096;
    while (end > 0) {
        const size_t n = (size_t)std::min<uint64_t>(end, blk);
        be.read(buf.data(), p + (end - n), n);
        for (size_t k = n; k > 0; --k) if (buf[k - 1] == '\n') return end - n + k;
        end -= n;
        blk = buf.size();
    }
    return 0;
}

struct ShardPhase {
    std::chrono::steady_clock::time_point t0;
    void start() { t0 = std::chrono::steady_clock::now(); }
    float ms() const { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

/*
 * The sharded walk of one rank.  in[0] / in[1]: this rank's byte shard of the primary / secondary stream.
 * out / out_cap as for the resident walk.  res: counts and n_records of the WHOLE job, out_len of this rank;
 * st: where this rank's bytes go.  Returns an xm_status, or XM_SHARD_DECLINED on every rank together.
 */
struct NoOutAlloc {
    int operator()(const uint64_t *, uint8_t **, uint64_t *) const { return 0; }
    explicit operator bool() const { return false; }
};
/* alloc_out(len[6], ptr[6], cap[6]): called once the six bin lengths of this rank are known, before anything is
 * copied, when the caller wants the bins sized exactly (it fills ptr / cap; nonzero return = out of memory) */
template <class BE, class CM, class OutAlloc = NoOutAlloc>
inline int walk_sharded(BE &be, CM &cm, ShardScratch &sc, const ShardBuf in[2], const xm_opts &o, uint8_t *const out[6],
                        const uint64_t out_cap[6], uint32_t debug, xm_result *res, xm_shard_stats *st, std::string &err,
                        OutAlloc alloc_out = OutAlloc())
{
    memset(res, 0, sizeof *res);
    memset(st, 0, sizeof *st);
    res->err_stream = -1;
    st->first_bad_rank = -1;
    const int W = cm.size(), r = cm.rank();
    const bool skip = (o.skip_repeated & 1) != 0, paired = o.mode != MODE_SE;
    const bool need_ctx = skip || paired;
    if (o.mode < 0 || o.mode > 2 || o.score_src < 0 || o.score_src > 2) { err = "bad mode / score_src"; return res->status = XM_ERR_ARG; }
    if (o.min_score != o.min_score) { err = "min_score is NaN"; return res->status = XM_ERR_UNSUPPORTED; }
    ShardPhase whole, ph;
    whole.start();
    /* a local failure must not leave the other ranks waiting in a collective: it travels with the next gather */
    int local_rc = XM_OK;
    auto fail = [&](int code, const std::string &msg) { if (!local_rc) { local_rc = code; err = msg; } };
    auto gather = [&](const std::vector<uint64_t> &mine, std::vector<uint64_t> &all) {
        all.assign(mine.size() * (size_t)W, 0);
        st->n_collectives++;
        ShardPhase c; c.start();
        const int rc = cm.all_gather(mine.data(), all.data(), mine.size() * 8);
        st->comm_ms += c.ms();
        if (rc) fail(XM_ERR_CUDA, "all-gather failed: " + cm.last_error());
        return rc;
    };
    auto swap_bytes = [&](std::vector<Xfer> &sends, std::vector<Xfer> &recvs) {
        ShardPhase c; c.start();
        for (auto &x : sends) st->sent_bytes += x.bytes;
        for (auto &x : recvs) st->sliver_bytes += x.bytes;
        const int rc = (sends.empty() && recvs.empty()) ? 0 : cm.exchange(sends.data(), (int)sends.size(), recvs.data(), (int)recvs.size());
        st->comm_ms += c.ms();
        if (rc) fail(XM_ERR_CUDA, "sliver exchange failed: " + cm.last_error());
        return rc;
    };
    /* every gather carries the sender's status in word 0; any rank's failure ends the walk on all of them */
    auto any_failed = [&](const std::vector<uint64_t> &all, size_t stride) {
        for (int q = 0; q < W; ++q) if (all[(size_t)q * stride]) { if (!local_rc) { local_rc = (int)all[(size_t)q * stride]; err = "rank " + std::to_string(q) + " failed"; } return true; }
        return false;
    };

    /* ---- A. line heads ------------------------------------------------------------------------------------- */
    ph.start();
    uint64_t head[2] = {0, 0};
    for (int s = 0; s < 2; ++s) head[s] = r == 0 ? 0 : first_line_end(be, in[s].p, in[s].len);
    std::vector<uint64_t> mineA = {(uint64_t)local_rc, head[0], head[1], in[0].len, in[1].len}, allA;
    if (gather(mineA, allA) || any_failed(allA, 5)) return res->status = local_rc;
    auto headq = [&](int q, int s) { return allA[(size_t)q * 5 + 1 + s]; };
    auto lenq = [&](int q, int s) { return allA[(size_t)q * 5 + 3 + s]; };
    auto allhead = [&](int q, int s) { return q > 0 && headq(q, s) == lenq(q, s); };      /* the whole shard is the tail of an earlier line */
    auto head_dest = [&](int q, int s) { int d = q - 1; while (d > 0 && allhead(d, s)) --d; return d; };
    uint64_t appended[2] = {0, 0};
    {
        std::vector<Xfer> sends, recvs;
        for (int s = 0; s < 2; ++s) {
            if (r > 0 && head[s] > 0) sends.push_back(Xfer{head_dest(r, s), in[s].p, head[s]});
            for (int q = r + 1; q < W; ++q) {
                if (headq(q, s) == 0 || head_dest(q, s) != r) continue;
                recvs.push_back(Xfer{q, in[s].p + in[s].len + appended[s], headq(q, s)});
                appended[s] += headq(q, s);
            }
            if (appended[s] + 16 > in[s].back_room) fail(XM_ERR_NOMEM, "shard buffer has no room behind it for the line that ends in the next shard");
        }
        std::vector<uint64_t> okA = {(uint64_t)local_rc}, allok;
        if (gather(okA, allok) || any_failed(allok, 1)) return res->status = local_rc;
        if (swap_bytes(sends, recvs)) { /* reported with the next gather */ }
    }
    uint64_t a_lo[2], a_len[2];                 /* this rank's whole lines: [p + a_lo, p + a_lo + a_len) */
    for (int s = 0; s < 2; ++s) { a_lo[s] = head[s]; a_len[s] = in[s].len - head[s] + appended[s]; }

    /* ---- B. context lines, filler ---------------------------------------------------------------------------- */
    uint64_t last_len[2] = {0, 0}, last_off[2] = {0, 0};
    if (need_ctx && !local_rc)
        for (int s = 0; s < 2; ++s)
            if (a_len[s]) { last_off[s] = last_line_start(be, in[s].p + a_lo[s], a_len[s]); last_len[s] = a_len[s] - last_off[s]; }
    std::vector<uint64_t> mineB = {(uint64_t)local_rc, a_len[0], a_len[1], last_len[0], last_len[1]}, allB;
    if (gather(mineB, allB) || any_failed(allB, 5)) return res->status = local_rc;
    uint64_t ctx_len[2] = {0, 0};
    {
        std::vector<Xfer> sends, recvs;
        for (int s = 0; s < 2 && need_ctx; ++s) {
            auto alen = [&](int q) { return allB[(size_t)q * 5 + 1 + s]; };
            auto llen = [&](int q) { return allB[(size_t)q * 5 + 3 + s]; };
            auto ctx_src = [&](int q) { int d = q - 1; while (d >= 0 && alen(d) == 0) --d; return d; };      /* nearest lower rank with lines */
            if (a_len[s]) {
                const int q = ctx_src(r);
                if (q >= 0) {
                    ctx_len[s] = llen(q);
                    if (ctx_len[s] + 32 > a_lo[s] + in[s].front_room) fail(XM_ERR_NOMEM, "shard buffer has no room in front of it for the context line");
                    else recvs.push_back(Xfer{q, in[s].p + a_lo[s] - ctx_len[s], ctx_len[s]});
                }
            }
            for (int q = r + 1; q < W && a_len[s]; ++q)
                if (alen(q) && ctx_src(q) == r) sends.push_back(Xfer{q, in[s].p + a_lo[s] + last_off[s], last_len[s]});
        }
        std::vector<uint64_t> okB = {(uint64_t)local_rc}, allok;
        if (gather(okB, allok) || any_failed(allok, 1)) return res->status = local_rc;
        swap_bytes(sends, recvs);
    }
    /* the scan starts on a 16-byte boundary: a filler line of 2..17 bytes in front, dropped with the context line's row */
    uint64_t scan_lo[2], scan_len[2], drop[2];   /* [p + scan_lo, +scan_len); leading rows that are not this rank's records */
    for (int s = 0; s < 2; ++s) {
        uint64_t lo = a_lo[s] - ctx_len[s];       /* may be "negative": counted from p */
        drop[s] = ctx_len[s] ? 1 : 0;
        const uint64_t lead = (uint64_t)((uintptr_t)(in[s].p + lo) & 15u);
        if (lead && a_len[s]) {
            const uint64_t L = lead >= 2 ? lead : 17;
            if (L + 16 > a_lo[s] - ctx_len[s] + in[s].front_room) { fail(XM_ERR_NOMEM, "shard buffer has no room in front of it"); }
            else {
                char first = 0;
                be.read(&first, in[s].p + lo, 1);
                char filler[17];
                memset(filler, first == 'x' ? 'y' : 'x', sizeof filler);      /* its QNAME is not the next line's */
                filler[L - 1] = '\n';
                if (be.write(in[s].p + lo - L, filler, (size_t)L)) fail(XM_ERR_CUDA, "filler write failed: " + be.last_error());
                lo -= L;
                drop[s] += 1;
            }
        }
        scan_lo[s] = lo;
        scan_len[s] = a_len[s] ? (a_lo[s] + a_len[s]) - lo : 0;
    }
    st->align_ms = ph.ms();

    /* ---- C. both shards into rows ------------------------------------------------------------------------------ */
    ph.start();
    Globals G;
    memset(&G, 0, sizeof G);
    uint8_t *ebase[2];                          /* row offsets count from here: a 16-byte boundary at or before the front room */
    uint64_t elen[2];
    for (int s = 0; s < 2; ++s) {
        uint8_t *lo = in[s].p - in[s].front_room;
        ebase[s] = lo + ((16 - ((uintptr_t)lo & 15)) & 15);
        elen[s] = (uint64_t)((in[s].p + in[s].len + in[s].back_room) - ebase[s]);
    }
    if (!local_rc) {
        if (!sc.g) sc.g = (Globals *)be.alloc(sizeof(Globals));
        bool ok = sc.g != nullptr;
        for (int s = 0; s < 2 && ok; ++s) {
            uint64_t need = 0;
            const StreamBuf B{in[s].p + scan_lo[s], scan_len[s]};
            if (scan_len[s]) {
                double mean = sc.line_bytes[s];
                if (mean <= 0) {
                    const size_t n = (size_t)std::min<uint64_t>(B.len, 256u << 10);
                    std::vector<char> smp(n);
                    be.read(smp.data(), B.p, n);
                    uint64_t nl = 0;
                    for (size_t k = 0; k < n; ++k) nl += smp[k] == '\n';
                    mean = nl ? (double)n / (double)nl : (double)n;
                }
                need = (uint64_t)((double)B.len / mean * 1.05) + 4096;
            }
            ok = shard_reserve_rows(be, sc.r[s], std::max<uint64_t>(need, 4096), be.scan2_tiles(scan_len[s]) + 1);
        }
        if (!ok) fail(XM_ERR_NOMEM, "out of device memory for the row arrays");
    }
    bool declined = false;
    if (!local_rc) {
        for (int attempt = 0; attempt < 4 && !local_rc; ++attempt) {
            Globals init;
            memset(&init, 0, sizeof init);
            init.err = NO_ERROR;
            init.limit_off = ~0ull;
            if (be.write(sc.g, &init, sizeof init)) { fail(XM_ERR_CUDA, "scratch init failed: " + be.last_error()); break; }
            int launched = 0;
            for (int s = 1; s >= 0 && !local_rc; --s) {
                if (!scan_len[s]) continue;
                if (be.zero(sc.r[s].chain1, sc.r[s].cap_tiles * 8)) { fail(XM_ERR_CUDA, "scratch init failed: " + be.last_error()); break; }
                ScanArgs a;
                memset(&a, 0, sizeof a);
                a.S = StreamBuf{in[s].p + scan_lo[s], scan_len[s]};
                a.sc = sc.r[s].rows; a.sc_cap = sc.r[s].cap; a.chain1 = sc.r[s].chain1; a.g = sc.g;
                a.score_src = o.score_src; a.skip = skip ? 1 : 0; a.stream_id = s; a.debug = 0;
                a.want_same = (s == 0 && paired && !skip) ? 1 : 0;
                a.start_bias = (uint64_t)((in[s].p + scan_lo[s]) - ebase[s]);
                a.short_lines = sc.short_lines;
                const int rc = be.scan2(a);
                if (rc < 0) { declined = true; break; }
                if (rc > 0) { fail(XM_ERR_CUDA, "scan kernel launch failed: " + be.last_error()); break; }
                ++launched;
            }
            if (local_rc || declined) break;
            if (be.read(&G, sc.g, sizeof G)) { fail(XM_ERR_CUDA, "kernel execution failed: " + be.last_error()); break; }
            res->n_launches += (uint32_t)launched;
            if (G.pad) {
                if (!sc.short_lines) { sc.short_lines = 1; continue; }        /* short reads overflow the spans: half the size, once */
                declined = true; break;
            }
            bool grown = false;
            for (int s = 0; s < 2; ++s)
                if (G.n_stream[s] > sc.r[s].cap) {
                    if (!shard_reserve_rows(be, sc.r[s], G.n_stream[s] + 1, sc.r[s].cap_tiles)) fail(XM_ERR_NOMEM, "out of device memory for the row arrays");
                    grown = true;
                }
            if (!grown) break;
        }
        for (int s = 0; s < 2; ++s)
            if (G.n_stream[s] > 1000 && scan_len[s]) sc.line_bytes[s] = (double)scan_len[s] / (double)G.n_stream[s];
    }
    st->index_ms = ph.ms();
    /* records of this rank's own lines, per stream (the filler and context rows are not its own) */
    uint64_t own[2] = {0, 0};
    for (int s = 0; s < 2; ++s) if (scan_len[s] && !local_rc && !declined) own[s] = G.n_stream[s] > drop[s] ? G.n_stream[s] - drop[s] : 0;

    /* ---- D. counts -------------------------------------------------------------------------------------------- */
    std::vector<uint64_t> mineC = {(uint64_t)local_rc, declined ? 1u : 0u, own[0], own[1]}, allC;
    if (gather(mineC, allC) || any_failed(allC, 4)) return res->status = local_rc;
    for (int q = 0; q < W; ++q) if (allC[(size_t)q * 4 + 1]) { err = "input needs the exact kernels (rank " + std::to_string(q) + ")"; return XM_SHARD_DECLINED; }
    std::vector<uint64_t> base[2];
    for (int s = 0; s < 2; ++s) {
        base[s].assign((size_t)W + 1, 0);
        for (int q = 0; q < W; ++q) base[s][(size_t)q + 1] = base[s][(size_t)q] + allC[(size_t)q * 4 + 2 + s];
    }
    const uint64_t N = std::min(base[0][(size_t)W], base[1][(size_t)W]);
    std::vector<uint64_t> cut((size_t)W + 1);
    for (int q = 0; q <= W; ++q) cut[(size_t)q] = std::min(base[0][(size_t)q], N);
    cut[(size_t)W] = N;
    auto has_ctx = [&](int q) { return need_ctx && cut[(size_t)q] > 0 && cut[(size_t)q + 1] > cut[(size_t)q]; };
    auto need_lo = [&](int q) { return cut[(size_t)q] - (has_ctx(q) ? 1 : 0); };
    auto need_hi = [&](int q) { return cut[(size_t)q + 1]; };

    /* ---- E. slivers: rows and text of the records a rank walks but another rank holds ---------------------------- */
    ph.start();
    struct Piece { int s, holder, consumer; uint64_t lo, hi; };
    std::vector<Piece> pieces;
    for (int s = 0; s < 2; ++s)
        for (int c = 0; c < W; ++c) {
            if (need_hi(c) <= cut[(size_t)c]) continue;                         /* walks nothing */
            for (int q = 0; q < W; ++q) {
                if (q == c) continue;
                const uint64_t lo = std::max(need_lo(c), base[s][(size_t)q]), hi = std::min(need_hi(c), base[s][(size_t)q + 1]);
                if (lo < hi) pieces.push_back(Piece{s, q, c, lo, hi});
            }
        }
    /* text bytes of the pieces this rank holds: first and last row offsets */
    auto row_start = [&](int s, long long local) {         /* byte offset (from ebase) of local row `local`; the end of the scan for the row past the last */
        unsigned long long v = 0;
        be.read(&v, sc.r[s].rows.start + local, 8);
        return (uint64_t)v;
    };
    std::vector<uint64_t> mineD((size_t)1 + 2 * (size_t)W, 0), allD;
    std::vector<std::pair<uint64_t, uint64_t>> held;       /* (text offset, text bytes) of the pieces this rank sends, in `pieces` order */
    for (auto &pc : pieces) {
        if (pc.holder != r) continue;
        const uint64_t l0 = pc.lo - base[pc.s][(size_t)r] + drop[pc.s], l1 = pc.hi - base[pc.s][(size_t)r] + drop[pc.s];
        const uint64_t t0 = row_start(pc.s, (long long)l0), t1 = row_start(pc.s, (long long)l1);
        held.push_back({t0, t1 - t0});
        mineD[(size_t)1 + (size_t)pc.s * (size_t)W + (size_t)pc.consumer] += t1 - t0;
    }
    mineD[0] = (uint64_t)local_rc;
    if (gather(mineD, allD) || any_failed(allD, mineD.size())) return res->status = local_rc;
    auto text_bytes = [&](int s, int holder, int consumer) { return allD[(size_t)holder * mineD.size() + 1 + (size_t)s * (size_t)W + (size_t)consumer]; };
    /* where this rank's own rows of its range start, and where the slivers go around them */
    long long first_row[2] = {0, 0};             /* row (relative to rows[0] of the own scan) that is record need_lo(r) after the exchange */
    uint64_t walked_bytes[2] = {0, 0};           /* text of the rows this rank walks: its own part and the slivers */
    uint64_t n_walk = need_hi(r) > need_lo(r) && need_hi(r) > cut[(size_t)r] ? need_hi(r) - need_lo(r) : 0;
    {
        std::vector<Xfer> sends, recvs;
        struct Rebase { int s; long long row; uint64_t n; uint64_t delta; };
        std::vector<Rebase> rebase;
        size_t hk = 0;
        uint64_t back_used[2] = {0, 0};
        for (int s = 0; s < 2; ++s) {
            /* own rows of the needed range: local rows [own_lo, own_hi) */
            const uint64_t lo = std::max(need_lo(r), base[s][(size_t)r]), hi = std::min(need_hi(r), base[s][(size_t)r + 1]);
            long long own_lo = (long long)drop[s], own_hi = (long long)drop[s];
            if (n_walk && lo < hi) { own_lo = (long long)(lo - base[s][(size_t)r] + drop[s]); own_hi = (long long)(hi - base[s][(size_t)r] + drop[s]); }
            else if (n_walk) {
                /* nothing of the range is held here: the slivers go behind this rank's rows, out of the way of what it sends */
                own_lo = own_hi = (long long)(own[s] + drop[s]);
            }
            uint64_t before = 0, after = 0;
            for (auto &pc : pieces) if (pc.s == s && pc.consumer == r) { if (pc.holder < r) before += pc.hi - pc.lo; else after += pc.hi - pc.lo; }
            if (before > sc.r[s].margin + (uint64_t)own_lo || (uint64_t)own_hi + after > sc.r[s].cap + sc.r[s].margin)
                fail(XM_ERR_NOMEM, "record-index shard and byte shard differ by more rows than the row arrays have room for");
            first_row[s] = own_lo - (long long)before;
            if (n_walk && own_hi > own_lo && !local_rc) walked_bytes[s] = row_start(s, own_hi) - row_start(s, own_lo);
            long long at_lo = own_lo - (long long)before, at_hi = own_hi;
            /* text of the slivers: the records just below the rank's own go right in front of its first line (over the
             * context line and the filler, which the scan no longer needs), the records just above right behind its last
             * line -- a pair unit whose two lines come from different ranks still finds them next to each other */
            uint64_t front_top = (uint64_t)((in[s].p + a_lo[s]) - ebase[s]);                /* grows downwards */
            uint64_t back_bot = (uint64_t)((in[s].p + a_lo[s] + a_len[s]) - ebase[s]);
            uint64_t lower_text = 0;
            for (auto &pc : pieces) if (pc.s == s && pc.consumer == r && pc.holder < r) lower_text += text_bytes(s, pc.holder, r);
            if (lower_text + 16 > front_top) fail(XM_ERR_NOMEM, "shard buffer has no room in front of it for the record slivers");
            uint64_t lower_at = front_top - std::min(lower_text, front_top);                /* lower pieces in rank order, ending at the own lines */
            for (auto &pc : pieces) {
                if (pc.s != s) continue;
                const uint64_t nrows = pc.hi - pc.lo;
                if (pc.holder == r) {
                    const uint64_t l0 = pc.lo - base[s][(size_t)r] + drop[s];
                    sends.push_back(Xfer{pc.consumer, sc.r[s].rows.start + l0, nrows * 8});
                    sends.push_back(Xfer{pc.consumer, sc.r[s].rows.rec + l0, nrows * 16});
                    sends.push_back(Xfer{pc.consumer, sc.r[s].rows.meta + l0, nrows * 4});
                    if (held[hk].second) sends.push_back(Xfer{pc.consumer, ebase[s] + held[hk].first, held[hk].second});
                    ++hk;
                } else if (pc.consumer == r) {
                    const uint64_t tb = text_bytes(s, pc.holder, r);
                    long long row;
                    uint64_t toff;
                    if (pc.holder < r) {
                        row = at_lo; at_lo += (long long)nrows;
                        toff = lower_at; lower_at += tb;
                        if (local_rc) continue;
                    } else {
                        row = at_hi; at_hi += (long long)nrows;
                        toff = back_bot + back_used[s];
                        back_used[s] += tb;
                        if (toff + tb + 16 > elen[s]) { fail(XM_ERR_NOMEM, "shard buffer has no room behind it for the record slivers"); continue; }
                    }
                    walked_bytes[s] += tb;
                    recvs.push_back(Xfer{pc.holder, sc.r[s].rows.start + row, nrows * 8});
                    recvs.push_back(Xfer{pc.holder, sc.r[s].rows.rec + row, nrows * 16});
                    recvs.push_back(Xfer{pc.holder, sc.r[s].rows.meta + row, nrows * 4});
                    if (tb) recvs.push_back(Xfer{pc.holder, ebase[s] + toff, tb});
                    rebase.push_back(Rebase{s, row, nrows, toff});
                }
            }
        }
        std::vector<uint64_t> okE = {(uint64_t)local_rc}, allok;
        if (gather(okE, allok) || any_failed(allok, 1)) return res->status = local_rc;
        swap_bytes(sends, recvs);
        /* received rows carry their holder's offsets: move them to where their text landed */
        for (auto &rb : rebase) {
            if (local_rc) break;
            unsigned long long first = 0;
            be.read(&first, sc.r[rb.s].rows.start + rb.row, 8);
            if (be.add64(sc.r[rb.s].rows.start + rb.row, rb.n, (unsigned long long)(rb.delta - first))) fail(XM_ERR_CUDA, "row rebase failed: " + be.last_error());
        }
    }
    st->sliver_ms = ph.ms();

    /* ---- F. the walk over this rank's rows ---------------------------------------------------------------------- */
    ph.start();
    Globals Gw;
    memset(&Gw, 0, sizeof Gw);
    int walk_rc = XM_OK;
    bool late_decline = false;
    if (!local_rc) {
        Globals init;
        memset(&init, 0, sizeof init);
        init.err = NO_ERROR;
        init.limit_off = ~0ull;
        EmitArgs ea;
        memset(&ea, 0, sizeof ea);
        ea.P = StreamBuf{ebase[0], elen[0]}; ea.S = StreamBuf{ebase[1], elen[1]};
        ea.rp.start = sc.r[0].rows.start + first_row[0]; ea.rp.rec = sc.r[0].rows.rec + first_row[0]; ea.rp.meta = sc.r[0].rows.meta + first_row[0];
        ea.rs.start = sc.r[1].rows.start + first_row[1]; ea.rs.rec = sc.r[1].rows.rec + first_row[1]; ea.rs.meta = sc.r[1].rows.meta + first_row[1];
        ea.n = n_walk;
        ea.mode = o.mode; ea.skip = skip ? 1 : 0; ea.halo = (n_walk && has_ctx(r)) ? 1 : 0; ea.exact_names = (debug & DBG_EXACT_NAMES) ? 1 : 0;
        ea.thr = score_threshold(o.min_score); ea.enabled = o.enabled_bins & 0x3f; ea.g = sc.g;
        ea.ntiles = (uint32_t)((n_walk + EM_TILE - 1) / EM_TILE);
        if ((uint64_t)ea.ntiles + 1 > sc.cap_tot) {
            be.release(sc.tile_tot);
            sc.tile_tot = (unsigned long long *)be.alloc(((uint64_t)ea.ntiles + 1) * 8 * C2_SLOTS);
            sc.cap_tot = sc.tile_tot ? (uint64_t)ea.ntiles + 1 : 0;
            if (!sc.tile_tot) fail(XM_ERR_NOMEM, "out of device memory for the tile totals");
        }
        ea.tile_tot = sc.tile_tot;
        for (int b = 0; b < 6; ++b) { ea.out[b] = out[b]; ea.out_cap[b] = ((ea.enabled >> b) & 1u) && out[b] ? out_cap[b] : 0; }
        if (!local_rc) {
            if (be.write(sc.g, &init, sizeof init) || be.size(ea) || be.read(&Gw, sc.g, sizeof Gw)) fail(XM_ERR_CUDA, "size kernel failed: " + be.last_error());
            else if (Gw.pad) late_decline = true;
            else if (be.prefix(ea) || be.read(&Gw, sc.g, sizeof Gw)) fail(XM_ERR_CUDA, "prefix kernel failed: " + be.last_error());
            res->n_launches += 2;
        }
        /* a QNAME mismatch or a flagged row on any rank: nobody writes, the caller takes the exact walk */
        std::vector<uint64_t> okF = {(uint64_t)local_rc, late_decline ? 1u : 0u}, allok;
        if (gather(okF, allok) || any_failed(allok, 2)) return res->status = local_rc;
        for (int q = 0; q < W; ++q) if (allok[(size_t)q * 2 + 1]) { err = "input needs the exact kernels (rank " + std::to_string(q) + ")"; return XM_SHARD_DECLINED; }
        if (alloc_out) {
            uint64_t want[6], cap6[6] = {0, 0, 0, 0, 0, 0};
            uint8_t *ptr6[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
            for (int b = 0; b < 6; ++b) want[b] = ((ea.enabled >> b) & 1u) ? Gw.out_len[b] : 0;
            if (alloc_out(want, ptr6, cap6)) { walk_rc = XM_ERR_NOMEM; err = "out of memory for this rank's bins"; }
            else for (int b = 0; b < 6; ++b) { ea.out[b] = ptr6[b]; ea.out_cap[b] = ((ea.enabled >> b) & 1u) ? cap6[b] : 0; }
        }
        for (int b = 0; b < 6; ++b)
            if (ea.out_cap[b] < Gw.out_len[b] && ((ea.enabled >> b) & 1u) && ea.out[b]) { walk_rc = XM_ERR_ARG; err = "output buffer too small for bin " + std::to_string(b); }
        if (be.emit(ea) || be.sync()) fail(XM_ERR_CUDA, "emit kernel failed: " + be.last_error());
        res->n_launches += 1;
    } else {
        std::vector<uint64_t> okF = {(uint64_t)local_rc, 0}, allok;
        gather(okF, allok);
        return res->status = local_rc;
    }
    st->walk_ms = ph.ms();

    /* ---- G. the whole job -------------------------------------------------------------------------------------- */
    std::vector<uint64_t> mineG((size_t)3 + 6 + 36, 0), allG;
    mineG[0] = (uint64_t)(local_rc ? local_rc : walk_rc);
    mineG[1] = n_walk - (n_walk && has_ctx(r) ? 1 : 0);
    for (int b = 0; b < 6; ++b) mineG[(size_t)3 + (size_t)b] = ((o.enabled_bins >> b) & 1u) ? Gw.out_len[b] : 0;
    for (int k = 0; k < 36; ++k) mineG[(size_t)9 + (size_t)k] = Gw.counts[k];
    if (gather(mineG, allG)) return res->status = local_rc;
    const size_t SG = mineG.size();
    for (int q = 0; q < W; ++q) {
        if (allG[(size_t)q * SG] && st->first_bad_rank < 0) st->first_bad_rank = q;
        res->n_records += allG[(size_t)q * SG + 1];
        for (int k = 0; k < 36; ++k) res->counts[k] += allG[(size_t)q * SG + 9 + (size_t)k];
        for (int b = 0; b < 6; ++b) {
            if (q < r) st->out_offset[b] += allG[(size_t)q * SG + 3 + (size_t)b];
            st->out_total[b] += allG[(size_t)q * SG + 3 + (size_t)b];
        }
    }
    for (int b = 0; b < 6; ++b) res->out_len[b] = mineG[(size_t)3 + (size_t)b];
    st->rec_lo = cut[(size_t)r]; st->rec_hi = cut[(size_t)r + 1];
    st->n_records_total = N;
    for (int s = 0; s < 2; ++s) res->bytes_in[s] = walked_bytes[s];       /* text behind the rows this rank walked (its context record included) */
    st->total_ms = whole.ms();
    res->ms_total = st->total_ms;
    res->ms_scan = st->index_ms; res->ms_classify = st->walk_ms;
    if (st->first_bad_rank >= 0) {
        const int code = (int)allG[(size_t)st->first_bad_rank * SG];
        if (st->first_bad_rank != r) err = "rank " + std::to_string(st->first_bad_rank) + " failed";
        return res->status = code;
    }
    return res->status = XM_OK;
}

}  // namespace xm
