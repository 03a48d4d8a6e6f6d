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
    std::vector<uint8_t> ibuf[2][2], obuf[2][6];
    DevIn dev[2];
    for (int s = 0; s < 2; ++s) {
        for (int k = 0; k < 2; ++k) { ibuf[s][k].assign(dcap + 64, 0xEE); dev[s].buf[k] = ibuf[s][k].data(); }
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
