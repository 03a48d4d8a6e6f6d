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
    const int rc = walk_resident(be, sc, StreamBuf{p.data(), plen}, StreamBuf{s.data(), slen}, *o, o6, cap, debug, res, msg);
    scratch_release(be, sc);
    if (errbuf && errcap) { strncpy(errbuf, msg.c_str(), errcap - 1); errbuf[errcap - 1] = 0; }
    return rc;
}
