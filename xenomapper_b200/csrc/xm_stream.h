/*
 * xm_stream.h -- the chunked walk: two host streams of any length are sent
 * through the GPU in bounded chunks, records aligned by index, the six bins
 * appended in order.  This is what replaces the reference's lockstep reader
 * on real files (getReadPairs, xm.py:95-118, driven by main(), xm.py:702-740).
 *
 * Each step stages the next bytes of both streams behind whatever the
 * previous step left unconsumed (the "carry"), runs the resident walk on the
 * two device buffers and hands the bins' new bytes to the sink.
 *
 *   - A chunk that is not the end of its stream is cut after its last '\n',
 *     so a step only ever sees complete lines; the cut-off tail is sent again
 *     with the next chunk.
 *   - The two streams hold different numbers of records per byte.  A step
 *     yields n = min(records of both buffers); the surplus records of the
 *     longer buffer are carried (device-to-device) to the front of its next
 *     buffer.
 *   - The last yielded record is carried too, as record 0 of the next step
 *     ("halo"): the pair predicate (xm.py:402) and the run-skipping reader
 *     (xm.py:110-114) look at the previous record.  The kernels do not
 *     classify a halo record again.
 *   - The walk ends like the reference's (xm.py:105): at the first blank line
 *     or at the end of either stream.
 *
 * Written against the same Backend interface as xm_walk.h plus
 *     int upload(void *dev_dst, const void *host_src, size_t n)   asynchronous H2D
 *     int upload_wait()
 *     int copy_dd(void *dev_dst, const void *dev_src, size_t n)
 * so that the CPU emulation harness of the tests runs the same logic.
 */
#pragma once
#include <string.h>

#include <algorithm>
#include <string>

#include "xm_walk.h"

namespace xm {

/* one input stream on the host */
struct HostIn {
    const uint8_t *mem = nullptr;      /* memory source (pageable or pinned) ... */
    uint64_t len = 0;
    int fd = -1;                       /* ... or a seekable descriptor: bytes [off, off + len) */
    int64_t off = 0;
    uint8_t *stage = nullptr;          /* pinned staging for descriptor sources, `stage_cap` bytes */
    uint64_t stage_cap = 0;
    uint64_t pos = 0;                  /* next byte to send */
};

/* device side of one input stream: two buffers used alternately */
struct DevIn {
    uint8_t *buf[2] = {nullptr, nullptr};
    uint64_t cap = 0;
    uint64_t len = 0;                  /* bytes of the current buffer */
};

struct StreamPlan {
    uint64_t chunk = 256ull << 20;     /* new bytes per stream and step */
};

/* bytes of [p, p + n) up to and including the last '\n', 0 if there is none */
inline uint64_t cut_after_last_newline(const uint8_t *p, uint64_t n)
{
    while (n > 0 && p[n - 1] != '\n') --n;
    return n;
}

/* read n bytes of a descriptor source at absolute offset `at` into dst; returns bytes read or -1 */
int64_t xm_pread_all(int fd, void *dst, uint64_t n, int64_t at);      /* xm_api.cu / the emulation harness */

/*
 * The chunked walk.  `outs[set][b]` are two sets of six device output buffers of `out_cap[b]` bytes each;
 * `emit(set, bin, dev_ptr, nbytes)` is called for every bin with new bytes after a step and must have consumed
 * (or queued a copy of) them before the same set is written again two steps later -- `emit_wait(set)` is called
 * for that.  Returns an xm_status; *res holds the totals of the whole walk.
 */
template <class BE, class Emit, class EmitWait>
inline int walk_stream(BE &be, Scratch &sc, HostIn in[2], DevIn dev[2], uint8_t *const outs[2][6], const uint64_t out_cap[6],
                       const xm_opts &o, uint32_t debug, const StreamPlan &plan, Emit emit, EmitWait emit_wait,
                       xm_result *res, std::string &errmsg, int first_is_context = 0)
{
    memset(res, 0, sizeof *res);
    res->err_stream = -1;
    uint64_t carry_off[2] = {0, 0}, carry_len[2] = {0, 0};      /* in the previous buffer */
    bool final_sent[2] = {false, false};
    int halo = first_is_context ? 1 : 0;                      /* sharded walks: the caller's buffers open with the record before its range */
    for (uint64_t step = 0;; ++step) {
        const int cur = (int)(step & 1), prev = cur ^ 1;
        uint64_t staged = 0;                                   /* new bytes this step, both streams */
        /* stage: carry to the front, new bytes behind it */
        for (int s = 0; s < 2; ++s) {
            if (carry_len[s] && be.copy_dd(dev[s].buf[cur], dev[s].buf[prev] + carry_off[s], carry_len[s])) { errmsg = "carry copy failed: " + be.last_error(); return res->status = XM_ERR_CUDA; }
            const uint64_t room = dev[s].cap - carry_len[s], remaining = in[s].len - in[s].pos;
            uint64_t take = std::min<uint64_t>(std::min<uint64_t>(room, plan.chunk), remaining);
            const uint8_t *src = nullptr;
            if (take) {
                if (in[s].mem) src = in[s].mem + in[s].pos;
                else {
                    if (take > in[s].stage_cap) take = in[s].stage_cap;
                    const int64_t got = xm_pread_all(in[s].fd, in[s].stage, take, in[s].off + (int64_t)in[s].pos);
                    if (got < 0) { errmsg = "read failed"; return res->status = XM_ERR_IO; }
                    if ((uint64_t)got < take) { in[s].len = in[s].pos + (uint64_t)got; take = (uint64_t)got; }     /* the file is shorter than announced */
                    src = in[s].stage;
                }
            }
            const bool is_final = in[s].pos + take == in[s].len;
            if (!is_final) {
                uint64_t cut = cut_after_last_newline(src, take);
                if (cut == 0 && in[s].mem) {
                    /* a line longer than the chunk: take it whole if the buffer has room for it */
                    const uint64_t lim = std::min<uint64_t>(room, in[s].len - in[s].pos);
                    const void *nl = lim > take ? memchr(src + take, '\n', lim - take) : nullptr;
                    if (nl) cut = (uint64_t)((const uint8_t *)nl - src) + 1;
                }
                if (cut == 0 && carry_len[s] == 0 && take > 0) {
                    errmsg = "a line is longer than the staging buffer (" + std::to_string(dev[s].cap) + " bytes)";
                    return res->status = XM_ERR_UNSUPPORTED;
                }
                take = cut;             /* 0: no complete line fits behind the carry this time; the other stream moves on */
            }
            const bool is_final2 = in[s].pos + take == in[s].len;
            if (take && be.upload(dev[s].buf[cur] + carry_len[s], src, take)) { errmsg = "H2D copy failed: " + be.last_error(); return res->status = XM_ERR_CUDA; }
            in[s].pos += take;
            staged += take;
            final_sent[s] = is_final2;
            dev[s].len = carry_len[s] + take;
        }
        if (be.upload_wait()) { errmsg = "H2D copy failed: " + be.last_error(); return res->status = XM_ERR_CUDA; }
        if (step >= 2) emit_wait(cur);

        /* walk the two buffers */
        xm_result r;
        WalkCtl ctl;
        ctl.halo = halo;
        ctl.want_tail = true;
        std::string msg;
        const int rc = walk_resident(be, sc, StreamBuf{dev[0].buf[cur], dev[0].len}, StreamBuf{dev[1].buf[cur], dev[1].len}, o,
                                     outs[cur], out_cap, debug, &r, msg, &ctl);
        res->n_launches += r.n_launches;
        res->ms_scan += r.ms_scan; res->ms_classify += r.ms_classify; res->ms_total += r.ms_total;
        if (rc == XM_ERR_CUDA || rc == XM_ERR_NOMEM || rc == XM_ERR_ARG) { errmsg = msg; return res->status = rc; }
        const uint64_t n = r.n_records;                              /* includes the halo record */
        const uint64_t fresh = n - (uint64_t)(n ? halo : 0);
        for (int k = 0; k < 36; ++k) res->counts[k] += r.counts[k];
        for (int b = 0; b < 6; ++b) {
            if (r.out_len[b]) emit(cur, b, outs[cur][b], r.out_len[b]);
            res->out_len[b] += r.out_len[b];
        }
        if (rc != XM_OK) {
            /* a failing record: everything before it has been produced (the reference's streaming writes) */
            errmsg = msg;
            res->err_stream = r.err_stream;
            res->err_record = res->n_records + (r.err_record - (uint64_t)halo);
            res->n_records += r.err_record - (uint64_t)halo;
            return res->status = rc;
        }
        res->n_records += fresh;
        /* the end of the walk (xm.py:105): the lockstep reader reaches a blank line, or the end of a stream */
        bool done = false;
        for (int s = 0; s < 2; ++s) {
            const bool drained = ctl.n_stream[s] == n;                   /* every record of this buffer was yielded */
            if (drained && ctl.stopped[s]) done = true;
            if (drained && final_sent[s] && in[s].pos == in[s].len) done = true;
        }
        for (int s = 0; s < 2; ++s) {
            const uint64_t last = s ? ctl.last_s : ctl.last_p;
            if (n > 0) res->bytes_in[s] += done ? (ctl.n_stream[s] == n ? ctl.end_off[s] : last) : last;
        }
        if (done) break;
        if (fresh == 0 && (dev[0].len == dev[0].cap || dev[1].len == dev[1].cap)) {
            errmsg = "no complete record fits the staging buffers";
            return res->status = XM_ERR_UNSUPPORTED;
        }
        if (fresh == 0 && staged == 0) {
            /* nothing was yielded and nothing new could be staged behind the carry: the next step would be this one
             * again.  A line that does not fit the staging chunk of a descriptor source ends here. */
            errmsg = "a line is longer than the staging buffer (" + std::to_string(plan.chunk) + " bytes per step)";
            return res->status = XM_ERR_UNSUPPORTED;
        }
        /* carry: from the last yielded record (the next step's halo) to the end of each buffer */
        if (n > 0) {
            carry_off[0] = ctl.last_p; carry_off[1] = ctl.last_s;
            halo = 1;
        } else {
            carry_off[0] = carry_off[1] = 0;
            halo = halo;            /* nothing was yielded: the buffers (halo included) are carried whole */
        }
        for (int s = 0; s < 2; ++s) carry_len[s] = dev[s].len - carry_off[s];
    }
    return res->status = XM_OK;
}

}  // namespace xm
