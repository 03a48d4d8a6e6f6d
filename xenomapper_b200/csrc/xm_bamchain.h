/*
 * xm_bamchain.h -- where the alignment records of an inflated BAM stream start, found in parallel.
 *
 * A BAM record opens with its own length (block_size, SAM specification section 4.2), so the
 * record starts form a chain that can only be followed from the front.  The stream is cut into
 * segments of SEG bytes; one thread per segment GUESSES the first record start inside its
 * segment (the first offset whose fixed fields are consistent, and whose successors' are too)
 * and follows the chain from there to the segment's end.  The host then checks the guesses in
 * order: segment k's chain is the true one if and only if it begins where the chain that came
 * from segment k-1 left off -- by induction from the first record, which is known.  A segment
 * whose guess is not confirmed is followed again from its true entry point.  Nothing about the
 * result is heuristic: the guess only decides how much of the work runs in parallel.
 *
 *   chain_segment()  one segment: entry (given or guessed) -> exit, record count, error flag
 *   chain_emit()     one segment again, writing the offsets behind the segment's base index
 *
 * Plain host/device code; tests/test_inflate.py runs it on the CPU against the serial chain.
 */
#pragma once
#include <stdint.h>

#include "xm_common.h"

namespace xm {

constexpr uint64_t CHAIN_NONE = ~0ull;
enum { CHAIN_OK = 0, CHAIN_CORRUPT = 1 };

struct ChainSeg {
    uint64_t entry;        /* first record start at or behind the segment's first byte (CHAIN_NONE: none found) */
    uint64_t exit;         /* first chain position at or behind the segment's end: the next segment's entry (or an incomplete record / the end) */
    uint32_t count;        /* record starts in [entry, segment end) that are whole inside the data */
    uint32_t flag;         /* CHAIN_CORRUPT: the chain from `entry` met a block_size below 32 */
};

XM_HD uint32_t chain_u32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

/* could a record start at byte p?  (fixed fields consistent with each other and with the reference list) */
XM_HD bool chain_plausible(const uint8_t *d, uint64_t have, uint64_t p, uint32_t n_ref)
{
    if (p + 36 > have) return false;
    const uint8_t *r = d + p;
    const uint32_t bs = chain_u32(r);
    if (bs < 32 || bs > (1u << 27)) return false;
    const int32_t ref = (int32_t)chain_u32(r + 4), pos = (int32_t)chain_u32(r + 8);
    const int32_t nref = (int32_t)chain_u32(r + 24), npos = (int32_t)chain_u32(r + 28);
    if (ref < -1 || ref >= (int32_t)n_ref || nref < -1 || nref >= (int32_t)n_ref || pos < -1 || npos < -1) return false;
    const uint32_t l_rn = r[12], n_cig = (uint32_t)r[16] | ((uint32_t)r[17] << 8), l_seq = chain_u32(r + 20);
    if (l_rn == 0 || l_seq > (1u << 27)) return false;
    const uint64_t fixed = 32ull + l_rn + 4ull * n_cig + ((uint64_t)l_seq + 1) / 2 + l_seq;
    if (fixed > bs) return false;
    const uint64_t nul = p + 36 + l_rn - 1;
    if (nul < have && d[nul] != 0) return false;
    return true;
}

/* Segment [lo, hi) of the `have` inflated bytes.  entry == CHAIN_NONE: guess it.  Records that are not whole inside the
 * data end the chain (exit = their start). */
XM_HD ChainSeg chain_segment(const uint8_t *d, uint64_t have, uint64_t lo, uint64_t hi, uint64_t entry, uint32_t n_ref)
{
    ChainSeg s;
    s.entry = entry; s.exit = CHAIN_NONE; s.count = 0; s.flag = CHAIN_OK;
    if (entry == CHAIN_NONE) {
        for (uint64_t q = lo; q < hi; ++q) {
            if (!chain_plausible(d, have, q, n_ref)) continue;
            /* two successors must look like records as well (or lie beyond the data) */
            uint64_t p = q;
            bool ok = true;
            for (int k = 0; k < 2 && ok; ++k) {
                p += 4ull + chain_u32(d + p);
                if (p + 36 > have) break;
                ok = chain_plausible(d, have, p, n_ref);
            }
            if (ok) { s.entry = q; break; }
        }
        if (s.entry == CHAIN_NONE) return s;
    }
    uint64_t p = s.entry;
    while (p < hi) {
        if (p + 4 > have) break;
        const uint32_t bs = chain_u32(d + p);
        if (bs < 32) { s.flag = CHAIN_CORRUPT; break; }
        if (p + 4 + (uint64_t)bs > have) break;
        ++s.count;
        p += 4ull + bs;
    }
    s.exit = p;
    return s;
}

/* the offsets of the segment's records (a confirmed segment): rec[0 .. count) */
XM_HD void chain_emit(const uint8_t *d, uint64_t have, uint64_t hi, uint64_t entry, uint32_t count, uint64_t *rec)
{
    uint64_t p = entry;
    for (uint32_t k = 0; k < count && p < hi && p + 4 <= have; ++k) {
        rec[k] = p;
        p += 4ull + chain_u32(d + p);
    }
}

#if defined(__CUDACC__)
/* entry_in: per segment, CHAIN_NONE (guess) or the entry to follow; n_seg threads */
__global__ void k_bam_chain(const uint8_t *d, uint64_t have, uint64_t seg_bytes, uint64_t first, uint32_t n_seg, uint32_t n_ref,
                            const uint64_t *entry_in, ChainSeg *seg)
{
    /* first == CHAIN_NONE: no record start is known (a rank's part of a file): every segment guesses */
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_seg) return;
    const uint64_t lo = (uint64_t)k * seg_bytes, hi = lo + seg_bytes < have ? lo + seg_bytes : have;
    uint64_t entry = entry_in ? entry_in[k] : CHAIN_NONE;
    const bool known = first != CHAIN_NONE;
    if (known && k == first / seg_bytes && !entry_in) entry = first;          /* the first record is known */
    if (known && !entry_in && lo + seg_bytes <= first) { ChainSeg s; s.entry = CHAIN_NONE; s.exit = CHAIN_NONE; s.count = 0; s.flag = 0; seg[k] = s; return; }
    seg[k] = chain_segment(d, have, lo, hi, entry, n_ref);
}
/* follow ONE segment from a given entry (the host's repair of an unconfirmed guess) */
__global__ void k_bam_chain_one(const uint8_t *d, uint64_t have, uint64_t lo, uint64_t hi, uint64_t entry, uint32_t n_ref, ChainSeg *out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = chain_segment(d, have, lo, hi, entry, n_ref);
}
__global__ void k_bam_chain_emit(const uint8_t *d, uint64_t have, uint64_t seg_bytes, uint32_t n_seg, const ChainSeg *seg, const uint64_t *base,
                                 uint64_t *rec)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_seg) return;
    const ChainSeg s = seg[k];
    if (!s.count) return;
    const uint64_t lo = (uint64_t)k * seg_bytes, hi = lo + seg_bytes < have ? lo + seg_bytes : have;
    chain_emit(d, have, hi, s.entry, s.count, rec + base[k]);
}
#endif

/*
 * The host's part: confirm the guesses in order.  seg[] comes from the device; `repair(k, entry, hi)` must follow segment k
 * from `entry` up to byte hi and return the result (one small launch; rare).  Leaves seg[k].count = 0 for segments no
 * record starts in, fills base[] (exclusive prefix of the counts) and returns the chain's end (the first byte that is not
 * part of a whole record); n_rec receives the number of records.  false: corrupt chain.
 * stop < have: only records that START before byte `stop` are counted (a rank's part of a file shared between GPUs);
 * `end` is then the first record start at or behind stop when the data reaches that far, and less than stop when it
 * does not (the caller fetches more).
 */
template <class Repair>
inline bool chain_confirm(ChainSeg *seg, uint64_t *base, uint32_t n_seg, uint64_t seg_bytes, uint64_t first, uint64_t have, uint64_t stop,
                          Repair &&repair, uint64_t &end, uint64_t &n_rec, uint32_t &n_repaired)
{
    uint64_t cur = first;
    n_rec = 0;
    n_repaired = 0;
    bool open = true;                       /* the chain still goes on (no incomplete record met) */
    for (uint32_t k = 0; k < n_seg; ++k) {
        const uint64_t lo = (uint64_t)k * seg_bytes, whole = lo + seg_bytes < have ? lo + seg_bytes : have;
        const uint64_t hi = whole < stop ? whole : stop;
        base[k] = n_rec;
        if (!open || lo >= stop || cur >= hi || cur < lo) { seg[k].count = 0; continue; }         /* no record of ours starts here */
        if (seg[k].entry != cur || hi != whole) { seg[k] = repair(k, cur, hi); ++n_repaired; }
        if (seg[k].flag == CHAIN_CORRUPT) return false;
        n_rec += seg[k].count;
        cur = seg[k].exit;
        if (cur < hi) open = false;          /* stopped inside the segment: an incomplete record (or the end of the data) */
    }
    end = cur;
    return true;
}

}  // namespace xm
