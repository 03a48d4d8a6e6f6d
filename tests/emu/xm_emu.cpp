/*
 * xm_emu.cpp -- TEST SCAFFOLDING ONLY.  Runs the tile programs of
 * xenomapper_b200/csrc/xm_tile.h on the CPU, one emulated CTA at a time, tiles
 * in order, behind the same host orchestration (xm_walk.h) the CUDA runtime
 * uses.  It lets the CPU test-suite exercise the kernels' parsing, indexing,
 * look-back bookkeeping and offset arithmetic without a GPU.  It is not part
 * of libxenomapper_b200.so and nothing in xenomapper_b200/ loads it: the
 * product has no CPU path.
 */
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../xenomapper_b200/csrc/xm_tile.h"
#include "../../xenomapper_b200/csrc/xm_walk.h"
#include "../../xenomapper_b200/csrc/xm_stream.h"

int64_t xm::xm_pread_all(int, void *, uint64_t, int64_t) { return -1; }     /* the emulation only streams from memory */

namespace {

using namespace xm;

template <class C>
struct EmuTile {
    std::vector<uint8_t> smem;
    std::vector<ThreadState<C>> threads;
    TileCtx<C> ctx;
    EmuTile() : smem(TileLayout<C>::total + 64), threads(C::THREADS)
    {
        uint8_t *b = smem.data();
        b += (16 - ((uintptr_t)b & 15)) & 15;
        ctx.m = carve<C>(b);
        ctx.emu = threads.data();
    }
    void reset()
    {
        /* poison, so that stale state from the previous tile cannot make a test pass by accident */
        memset(smem.data(), 0xA5, smem.size());
        memset((void *)threads.data(), 0x5A, threads.size() * sizeof(ThreadState<C>));
    }
};

struct EmuBackend {
    void *alloc(size_t n) { return calloc(1, n + 64); }
    void release(void *p) { free(p); }
    int zero(void *p, size_t n) { memset(p, 0, n); return 0; }
    int write(void *d, const void *s, size_t n) { memcpy(d, s, n); return 0; }
    int read(void *d, const void *s, size_t n) { memcpy(d, s, n); return 0; }
    int read_words(const void *const src[], int n, unsigned long long *dst) { for (int k = 0; k < n; ++k) memcpy(dst + k, src[k], 8); return 0; }
    int sync() { return 0; }
    int upload(void *d, const void *s, size_t n) { memcpy(d, s, n); return 0; }
    int upload_wait() { return 0; }
    int copy_dd(void *d, const void *s, size_t n) { memmove(d, s, n); return 0; }
    void tick(int) {}
    float elapsed(int, int) { return 0.f; }
    std::string last_error() { return ""; }
    template <class C>
    void scan_t(const ScanArgs &a)
    {
        EmuTile<C> t;
        for (uint32_t k = 0; k < a.ntiles; ++k) { t.reset(); scan_tile<C>(t.ctx, a, k); }
    }
    template <class C>
    void classify_t(const ClassifyArgs &a)
    {
        EmuTile<C> t;
        for (uint32_t k = 0; k < a.ntiles; ++k) { t.reset(); classify_tile<C>(t.ctx, a, k); }
    }
    int scan(const ScanArgs &a, bool small) { if (small) scan_t<CfgSmall>(a); else scan_t<CfgBig>(a); return 0; }
    int scan2(const ScanArgs &) { return -1; }      /* the warp-synchronous scan exists on the device only */
    uint64_t scan2_tiles(uint64_t) { return 0; }
    int classify2(const ClassifyArgs &) { return -1; }
    bool rows_enabled() { return false; }          /* the row kernels exist on the device only */
    int size(const EmitArgs &) { return -1; }
    int prefix(const EmitArgs &) { return -1; }
    int emit(const EmitArgs &) { return -1; }
    int classify(const ClassifyArgs &a, bool small) { if (small) classify_t<CfgSmall>(a); else classify_t<CfgBig>(a); return 0; }
};

}  // namespace

extern "C" int xm_emu_classify(const void *prim, uint64_t plen, const void *sec, uint64_t slen, const xm_opts *o,
                               uint32_t debug, void *const out[6], const uint64_t cap[6], xm_result *res,
                               char *errbuf, size_t errcap)
{
    /* device buffers are readable up to the next multiple of 16: give the emulation the same slack */
    std::vector<uint8_t> p(plen + 64, 0xEE), s(slen + 64, 0xEE);
    if (plen) memcpy(p.data(), prim, plen);
    if (slen) memcpy(s.data(), sec, slen);
    EmuBackend be;
    Scratch sc;
    std::string msg;
    uint8_t *o6[6];
    for (int b = 0; b < 6; ++b) o6[b] = (uint8_t *)out[b];
    /* bit 1 of the reader flags: record 0 is context (sharded walks), exactly as xm_classify_device decodes it */
    WalkCtl ctl;
    ctl.halo = (o->skip_repeated >> 1) & 1;
    xm_opts oo = *o;
    oo.skip_repeated &= 1;
    const int rc = walk_resident(be, sc, StreamBuf{p.data(), plen}, StreamBuf{s.data(), slen}, oo, o6, cap, debug, res, msg, &ctl);
    if (ctl.halo && res->n_records) res->n_records -= 1;
    scratch_release(be, sc);
    if (errbuf && errcap) { strncpy(errbuf, msg.c_str(), errcap - 1); errbuf[errcap - 1] = 0; }
    return rc;
}

/* the chunked walk (xm_stream.h) over the emulated kernels: `chunk` new bytes per stream and step */
extern "C" int xm_emu_classify_stream(const void *prim, uint64_t plen, const void *sec, uint64_t slen, const xm_opts *o,
                                      uint32_t debug, uint64_t chunk, void *const out[6], const uint64_t cap[6], xm_result *res,
                                      char *errbuf, size_t errcap)
{
    EmuBackend be;
    Scratch sc;
    HostIn in[2];
    static const uint8_t nothing = 0;
    in[0].mem = plen ? (const uint8_t *)prim : &nothing; in[0].len = plen;
    in[1].mem = slen ? (const uint8_t *)sec : &nothing; in[1].len = slen;
    StreamPlan plan;
    plan.chunk = chunk;
    const uint64_t dcap = 2 * chunk + 64;
    std::vector<uint8_t> ibuf[2][2], sbuf[2][2], obuf[2][6];
    DevIn dev[2];
    for (int s = 0; s < 2; ++s) {
        for (int k = 0; k < 2; ++k) {
            ibuf[s][k].assign(dcap + 64, 0xEE); dev[s].buf[k] = ibuf[s][k].data();
            sbuf[s][k].assign(dcap + 64, 0xEE); dev[s].stage[k] = sbuf[s][k].data();
        }
        dev[s].cap = dcap;
    }
    uint64_t ocap[6];
    uint8_t *outs[2][6];
    for (int b = 0; b < 6; ++b) {
        ocap[b] = 4 * dcap + 64;
        for (int k = 0; k < 2; ++k) { obuf[k][b].assign(ocap[b], 0); outs[k][b] = obuf[k][b].data(); }
    }
    uint64_t filled[6] = {0, 0, 0, 0, 0, 0};
    bool overflow = false;
    auto emit = [&](int, int b, const uint8_t *src, uint64_t n) {
        if (b < 0) return;
        if (filled[b] + n > cap[b]) { overflow = true; return; }
        memcpy((uint8_t *)out[b] + filled[b], src, n);
        filled[b] += n;
    };
    auto emit_wait = [&](int) {};
    std::string msg;
    int rc = walk_stream(be, sc, in, dev, outs, ocap, *o, debug, plan, emit, emit_wait, res, msg);
    scratch_release(be, sc);
    if (overflow) { rc = XM_ERR_ARG; msg = "test output buffer too small"; }
    if (errbuf && errcap) { strncpy(errbuf, msg.c_str(), errcap - 1); errbuf[errcap - 1] = 0; }
    return rc;
}


/* the index pass of the sharded walk (xm_walk.h index_resident) over the emulated scan kernel */
extern "C" int xm_emu_index(const void *buf, uint64_t len, int skip, uint32_t debug, uint32_t nq, const uint64_t *q, uint64_t *off,
                            xm_shard_info *info, char *errbuf, size_t errcap)
{
    std::vector<uint8_t> b(len + 64, 0xEE);
    if (len) memcpy(b.data(), buf, len);
    EmuBackend be;
    Scratch sc;
    std::string msg;
    const int rc = index_resident(be, sc, StreamBuf{b.data(), len}, skip != 0, debug, nq, q, off, info, msg);
    scratch_release(be, sc);
    if (errbuf && errcap) { strncpy(errbuf, msg.c_str(), errcap - 1); errbuf[errcap - 1] = 0; }
    return rc;
}

/* ================================================================================================================
 * The walk across ranks (xm_shard.h) on the CPU: the same host logic as the NCCL runtime, with
 *   - scalar stand-ins for the device-only row kernels (k_scan2 -> rows, k_size, k_prefix, k_emit): they follow the
 *     same contract (rows of clean lines, Globals::pad for anything else) so that the orchestration -- line heads,
 *     context lines, filler, counts, slivers, placement in the bins -- is what gets tested;
 *   - a Comm whose all-gather and send/receive are callbacks into the test process (torch.distributed, gloo).
 * ================================================================================================================ */
#include "../../xenomapper_b200/csrc/xm_shard.h"

namespace {

struct EmuRowsBackend : EmuBackend {
    int add64(void *p, uint64_t n, unsigned long long d) { unsigned long long *q = (unsigned long long *)p; for (uint64_t k = 0; k < n; ++k) q[k] += d; return 0; }
    uint64_t scan2_tiles(uint64_t len) { return len / 4096 + 1; }
    /* rows of every line (run heads with skip); a blank line raises pad, flagged lines keep their flags for k_size */
    int scan2(const ScanArgs &a)
    {
        const Reader rd{a.S.p, a.S.p, 0, (uint32_t)std::min<uint64_t>(a.S.len, 0xffffffffu), a.S.len};
        uint64_t n = 0, off = 0;
        LineRec prev;
        bool have_prev = false;
        while (off < a.S.len) {
            LineRec L;
            generic_parse(rd, off, a.score_src, L);
            if (L.flags & F_BLANK) { if (getenv("XM_EMU_DEBUG")) fprintf(stderr, "[emu scan] blank line at %llu of %llu\n", (unsigned long long)off, (unsigned long long)a.S.len); a.g->pad = 1; break; }
            bool same = false;
            if (have_prev && (a.skip || a.want_same) && prev.qlen == L.qlen && memcmp(a.S.p + prev.qs, a.S.p + L.qs, L.qlen) == 0) same = true;
            if (!(a.skip && same)) {
                if (n < a.sc_cap) {
                    a.sc.start[n] = a.start_bias + off;
                    a.sc.rec[n] = make_uint4((uint32_t)L.as, (uint32_t)L.xs, L.h1, L.h2);
                    a.sc.meta[n] = (L.outlen & META_LEN_MASK) | ((L.flags & 0x3fu) << META_LEN_BITS) | (same ? META_SAME : 0u);
                }
                ++n;
            }
            prev = L; have_prev = true;
            off += L.rawbytes;
        }
        a.g->n_stream[a.stream_id] = n;
        a.g->end_off[a.stream_id] = a.S.len;
        if (n <= a.sc_cap) a.sc.start[n] = a.start_bias + a.S.len;
        return 0;
    }
    /* the decision of one record from its rows (xm_emit.cuh rows_decide, scalar) */
    struct Dec { uint32_t key, bin, plen, slen; uint64_t psrc, ssrc; };
    static Dec decide(const EmitArgs &a, uint64_t i, bool check, bool &bad)
    {
        Dec d{36, NO_BIN, 0, 0, 0, 0};
        auto state = [&](uint64_t k) { const uint4 p = a.rp.rec[k], s = a.rs.rec[k]; return mapping_state((int32_t)p.x, (int32_t)p.y, (int32_t)s.x, (int32_t)s.y, a.thr); };
        const uint32_t pm = a.rp.meta[i], sm = a.rs.meta[i];
        const uint64_t ps = a.rp.start[i], ss = a.rs.start[i];
        if (check) {
            const uint4 p = a.rp.rec[i], s = a.rs.rec[i];
            if (((pm | sm) & META_FLAGS) || p.z != s.z || p.w != s.w) bad = true;
            else {
                const uint8_t *x = a.P.p + ps, *y = a.S.p + ss;
                for (;; ++x, ++y) { if (*x != *y) { bad = true; break; } if (*x < 0x21) break; }
            }
        }
        const int st = state(i);
        const uint32_t pout = pm & META_LEN_MASK, sout = sm & META_LEN_MASK;
        if (a.mode == MODE_SE) {
            if (!(a.halo && i == 0)) {
                d.key = (uint32_t)st; d.bin = (uint32_t)st;
                const bool pside = st == PS || st == PM || st == UA || st == UR, sside = st == SS || st == SM || st == UR;
                d.plen = pside ? pout : 0; d.slen = sside ? sout : 0; d.psrc = ps; d.ssrc = ss;
            }
        } else if (!a.skip && (pm & META_SAME) && i > 0) {
            const int pst = state(i - 1);
            const uint32_t ppout = a.rp.meta[i - 1] & META_LEN_MASK, psout = a.rs.meta[i - 1] & META_LEN_MASK;
            d.key = (uint32_t)(pst * 6 + st);
            d.bin = (uint32_t)(a.mode == MODE_PE_CONSERVATIVE ? pair_bin_conservative(pst, st) : pair_bin_liberal(pst, st));
            const bool pside = d.bin == PS || d.bin == PM || d.bin == UA || d.bin == UR, sside = d.bin == SS || d.bin == SM || d.bin == UR;
            d.plen = pside ? ppout + pout : 0; d.slen = sside ? psout + sout : 0; d.psrc = a.rp.start[i - 1]; d.ssrc = a.rs.start[i - 1];
            if (check && (d.psrc + ppout != ps || d.ssrc + psout != ss)) bad = true;
        }
        if (d.bin != NO_BIN && !((a.enabled >> d.bin) & 1u)) { d.plen = 0; d.slen = 0; }
        return d;
    }
    int size(const EmitArgs &a)
    {
        bool bad = false;
        unsigned long long tot[6] = {0, 0, 0, 0, 0, 0};
        for (uint64_t i = 0; i < a.n; ++i) {
            const bool was = bad;
            const Dec d = decide(a, i, true, bad);
            if (bad && !was && getenv("XM_EMU_DEBUG"))
                fprintf(stderr, "[emu size] record %llu of %llu bad: pmeta %x smeta %x pstart %llu sstart %llu P:'%.24s' S:'%.24s'\n", (unsigned long long)i, (unsigned long long)a.n,
                        a.rp.meta[i], a.rs.meta[i], (unsigned long long)a.rp.start[i], (unsigned long long)a.rs.start[i], a.P.p + a.rp.start[i], a.S.p + a.rs.start[i]);
            if (d.key < 36) a.g->counts[d.key]++;
            if (d.bin < 6) tot[d.bin] += d.plen + d.slen;
        }
        if (bad) a.g->pad = 1;
        for (int b = 0; b < 6; ++b) a.tile_tot[b] = tot[b];      /* one tile */
        return 0;
    }
    int prefix(const EmitArgs &a) { for (int b = 0; b < 6; ++b) { a.g->out_len[b] = a.tile_tot[b]; a.tile_tot[b] = 0; } return 0; }
    int emit(const EmitArgs &a)
    {
        bool bad = false;
        unsigned long long at[6] = {0, 0, 0, 0, 0, 0};
        for (uint64_t i = 0; i < a.n; ++i) {
            const Dec d = decide(a, i, false, bad);
            if (d.bin >= 6) continue;
            if (d.plen && at[d.bin] + d.plen <= a.out_cap[d.bin]) memcpy(a.out[d.bin] + at[d.bin], a.P.p + d.psrc, d.plen);
            at[d.bin] += d.plen;
            if (d.slen && at[d.bin] + d.slen <= a.out_cap[d.bin]) memcpy(a.out[d.bin] + at[d.bin], a.S.p + d.ssrc, d.slen);
            at[d.bin] += d.slen;
        }
        return 0;
    }
};

typedef int (*emu_all_gather_fn)(const void *send, void *recv, uint64_t bytes);
typedef int (*emu_exchange_fn)(const Xfer *sends, int ns, const Xfer *recvs, int nr);
struct EmuComm {
    int my, n;
    emu_all_gather_fn ag;
    emu_exchange_fn ex;
    int rank() const { return my; }
    int size() const { return n; }
    std::string last_error() const { return "callback failed"; }
    int all_gather(const void *send, void *recv, size_t bytes) { if (n == 1) { memcpy(recv, send, bytes); return 0; } return ag(send, recv, bytes); }
    int exchange(const Xfer *sends, int ns, const Xfer *recvs, int nr) { return ex(sends, ns, recvs, nr); }
};

}  // namespace

extern "C" int xm_emu_classify_sharded(const void *prim, uint64_t plen, const void *sec, uint64_t slen, const xm_opts *o, int rank, int world,
                                       emu_all_gather_fn ag, emu_exchange_fn ex, uint64_t room, void *const out[6], const uint64_t cap[6],
                                       xm_result *res, xm_shard_stats *st, char *errbuf, size_t errcap)
{
    /* shard buffers with room on both sides, deliberately misaligned so that the filler line is exercised */
    std::vector<uint8_t> buf[2];
    const void *src[2] = {prim, sec};
    const uint64_t len[2] = {plen, slen};
    ShardBuf in[2];
    for (int s = 0; s < 2; ++s) {
        buf[s].assign(len[s] + 2 * room + 96, 0xEE);
        uint8_t *p = buf[s].data() + room + 16;
        p += (16 - ((uintptr_t)p & 15)) & 15;
        p += (rank * 5 + s * 3) & 15;
        if (len[s]) memcpy(p, src[s], len[s]);
        in[s] = ShardBuf{p, len[s], room, room};
    }
    EmuRowsBackend be;
    EmuComm cm{rank, world, ag, ex};
    ShardScratch sc;
    std::string msg;
    uint8_t *o6[6];
    for (int b = 0; b < 6; ++b) o6[b] = (uint8_t *)out[b];
    const int rc = walk_sharded(be, cm, sc, in, *o, o6, cap, 0, res, st, msg);
    shard_release(be, sc);
    if (errbuf && errcap) { strncpy(errbuf, msg.c_str(), errcap - 1); errbuf[errcap - 1] = 0; }
    return rc;
}

/* ---- BGZF inflate and the BAM record chain (csrc/xm_inflate.h, csrc/xm_bamchain.h) on the CPU ------------------------ */
#include "../../xenomapper_b200/csrc/xm_inflate.h"
#include "../../xenomapper_b200/csrc/xm_bamchain.h"

/* `in` must be readable from the aligned word below it to 12 bytes past its end */
extern "C" int xm_emu_inflate(const void *in, uint32_t in_len, void *out, uint32_t out_len)
{
    std::vector<xm::InflateTables> t(1);
    uint32_t base[xm::INF_BASE_WORDS];
    for (int k = 0; k < xm::INF_BASE_WORDS; ++k) base[k] = xm::inf_base_word(k);
    return xm::inflate_raw((const uint8_t *)in, in_len, (uint8_t *)out, out_len, t[0], base);
}

extern "C" uint32_t xm_emu_crc32(const void *p, uint32_t n) { return xm::crc32_by_lanes((const uint8_t *)p, n); }

/* the parallel chain, one emulated thread per segment; returns the number of records (-1: corrupt, -2: rec[] too small) */
extern "C" int64_t xm_emu_bam_chain(const void *data, uint64_t have, uint64_t seg_bytes, uint64_t first, uint32_t n_ref, uint64_t *rec,
                                    uint64_t cap, uint64_t *end, uint32_t *repaired, uint64_t stop, uint64_t *guess)
{
    using namespace xm;
    const uint8_t *d = (const uint8_t *)data;
    const uint32_t n_seg = (uint32_t)((have + seg_bytes - 1) / seg_bytes);
    std::vector<ChainSeg> seg(n_seg);
    std::vector<uint64_t> base(n_seg + 1);
    for (uint32_t k = 0; k < n_seg; ++k) {
        const uint64_t lo = (uint64_t)k * seg_bytes, hi = std::min(lo + seg_bytes, have);
        if (first != CHAIN_NONE && lo + seg_bytes <= first) { seg[k] = ChainSeg{CHAIN_NONE, CHAIN_NONE, 0, 0}; continue; }
        seg[k] = chain_segment(d, have, lo, hi, (first != CHAIN_NONE && k == first / seg_bytes) ? first : CHAIN_NONE, n_ref);
    }
    uint64_t n_rec = 0;
    auto repair = [&](uint32_t k, uint64_t entry, uint64_t hi) {
        return chain_segment(d, have, (uint64_t)k * seg_bytes, hi, entry, n_ref);
    };
    if (stop == 0 || stop > have) stop = have;
    if (first == CHAIN_NONE) {
        /* a part of a file: the chain starts at the first guess */
        for (uint32_t k = 0; k < n_seg && first == CHAIN_NONE; ++k) first = seg[k].entry;
        if (first == CHAIN_NONE) { *end = have; *repaired = 0; return 0; }
    }
    *guess = first;
    if (!chain_confirm(seg.data(), base.data(), n_seg, seg_bytes, first, have, stop, repair, *end, n_rec, *repaired)) return -1;
    if (n_rec > cap) return -2;
    for (uint32_t k = 0; k < n_seg; ++k) {
        const uint64_t lo = (uint64_t)k * seg_bytes, hi = std::min(std::min(lo + seg_bytes, have), stop);
        if (seg[k].count) chain_emit(d, have, hi, seg[k].entry, seg[k].count, rec + base[k]);
    }
    return (int64_t)n_rec;
}

/* ---- the host half of the device BGZF compressor (csrc/xm_deflate.h): the plan -------------------------------------- */
#include "../../xenomapper_b200/csrc/xm_deflate.h"

/* `data` as ONE raw DEFLATE block of literals coded with the plan made from `sample` (hist: 316 token counts, or NULL for the
 * first guess): header bits as the kernel copies them, then the literal codes and the end-of-block code.  Returns the byte count. */
extern "C" uint64_t xm_emu_deflate_literals(const void *sample, uint64_t ns, const uint32_t *hist, const void *data, uint64_t n, void *out, uint64_t cap,
                                            uint8_t *lit_len, uint8_t *dist_len)
{
    using namespace xm;
    DeflatePlan P;
    if (hist) deflate_plan_hist(hist, P); else deflate_plan((const uint8_t *)sample, ns, P);
    for (int s = 0; s < 286; ++s) lit_len[s] = (uint8_t)(P.lit[s] >> 16);
    for (int s = 0; s < 30; ++s) dist_len[s] = (uint8_t)(P.dist[s] >> 16);
    std::vector<uint8_t> bits;                       /* one byte per bit: simple and obviously right */
    for (uint32_t b = 16; b < P.hdr_bits; ++b) bits.push_back((P.hdr_words[b >> 5] >> (b & 31)) & 1u);
    auto put = [&](uint32_t e) { for (uint32_t k = 0; k < (e >> 16); ++k) bits.push_back((e >> k) & 1u); };
    for (uint64_t k = 0; k < n; ++k) put(P.lit[((const uint8_t *)data)[k]]);
    put(P.lit[256]);
    const uint64_t nbytes = (bits.size() + 7) / 8;
    if (nbytes > cap) return 0;
    memset(out, 0, nbytes);
    for (size_t b = 0; b < bits.size(); ++b) if (bits[b]) ((uint8_t *)out)[b >> 3] |= (uint8_t)(1u << (b & 7));
    return nbytes;
}
