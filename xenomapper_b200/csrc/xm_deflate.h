/*
 * xm_deflate.h -- BGZF output deflated on the GPU (the six bins as blocked gzip, SAM/BAM specification 4.1).
 *
 * The reference writes SAM text and leaves compression to a pipe into `samtools view -bS`
 * (README.md:138, xm.py:590-592).  With XM_OUT_BGZF the bins of a descriptor walk leave the
 * device compressed: less to copy over PCIe, nothing to deflate on the host.
 *
 *   host    deflate_plan(): ONE prefix code per bin and step -- literal counts from a sample of
 *           the bin's bytes, fixed pseudo-counts for the length and distance symbols -- built
 *           with a length limit of 15 (counts folded the way miniz does), and the bits of the
 *           dynamic-block header that announces it (every code length sent as a 4-bit symbol:
 *           168 bytes per 64 KiB member).  SAM text is DNA letters, quality characters, digits
 *           and tabs: a code fitted to them carries most of what per-block codes would.
 *   k_bgzf_deflate   one warp per member (0xff00 input bytes).  32 positions per step, one per
 *           lane: 4-byte hash into a per-warp table of last positions in shared memory, the
 *           candidate (and the run candidate, distance 1) verified and extended by each lane,
 *           matches of four bytes or more selected greedily in position order by ballots, the
 *           tokens' bits placed by a warp scan of their lengths and ORed into a staging word
 *           window in shared memory, whole words stored coalesced.  CRC-32 by lanes
 *           (xm_inflate.h).  A member that would not fit leaves as a stored block.
 *   k_bgzf_pack      the members, written into slots of 64 KiB, moved next to each other.
 *
 * Any inflater reads the result; tests/test_bgzf.py checks it with zlib/gzip and with the
 * device inflater of xm_inflate.h.
 */
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "xm_common.h"
#include "xm_inflate.h"

namespace xm {

constexpr uint32_t DEF_IN_MAX = 0xff00;            /* most input bytes per member, as htslib */
constexpr uint32_t DEF_IN_MIN = 0x4000;            /* fewest: small bins are cut finer so that every SM has members to work on */
constexpr uint32_t DEF_SLOT = 0x10000 + 256;       /* bytes of scratch per member of DEF_IN_MAX bytes */
constexpr uint32_t DEF_MEMBER_MAX = 0x10000;       /* BSIZE is 16 bits */
constexpr int DEF_HASH_BITS = 12;
constexpr int DEF_MIN_MATCH = 4, DEF_MAX_MATCH = 258;

/* what the kernel needs of the plan */
struct DeflatePlan {
    uint32_t lit[286];        /* bit-reversed code | code length << 16, literal/length alphabet */
    uint32_t dist[30];
    uint32_t hdr_words[64];   /* the member's words 4.. (byte 16 on): bytes 16, 17 zero (BSIZE, patched), then the block header's bits */
    uint32_t hdr_bits;        /* bits in hdr_words, the 16 of BSIZE included */
    uint32_t pad;
};

/* ---- host: the plan ------------------------------------------------------------------------------------------ */
/* code lengths (1..max_len, every symbol gets one) of a Huffman code for freq[0..n) */
inline void huffman_lengths(const uint64_t *freq, int n, int max_len, uint8_t *len)
{
    struct Node { uint64_t w; int left, right; };
    std::vector<Node> nodes;
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return freq[a] < freq[b]; });
    /* two-queue construction over the sorted leaves */
    nodes.reserve(2 * n);
    for (int i = 0; i < n; ++i) nodes.push_back(Node{std::max<uint64_t>(freq[order[i]], 1), -1, -1});
    size_t q1 = 0, q2 = (size_t)n;
    auto pop = [&]() {
        int pick;
        if (q1 < (size_t)n && (q2 >= nodes.size() || nodes[q1].w <= nodes[q2].w)) pick = (int)q1++;
        else pick = (int)q2++;
        return pick;
    };
    while ((size_t)n - q1 + (nodes.size() - q2) > 1) {
        const int a = pop(), b = pop();
        nodes.push_back(Node{nodes[a].w + nodes[b].w, a, b});
    }
    std::vector<int> depth(nodes.size(), 0);
    for (size_t k = nodes.size(); k-- > 0;) {
        if (nodes[k].left >= 0) { depth[nodes[k].left] = depth[k] + 1; depth[nodes[k].right] = depth[k] + 1; }
    }
    /* fold the depths into at most max_len (miniz: tdefl_huffman_enforce_max_code_size) */
    std::vector<int> count(max_len + 2, 0);
    for (int i = 0; i < n; ++i) count[std::min(std::max(depth[i], 1), max_len)]++;
    uint64_t total = 0;
    for (int l = max_len; l > 0; --l) total += (uint64_t)count[l] << (max_len - l);
    while (total > (1ull << max_len)) {
        count[max_len]--;
        for (int l = max_len - 1; l > 0; --l)
            if (count[l]) { count[l]--; count[l + 1] += 2; break; }
        total--;
    }
    /* the most frequent symbols get the shortest codes: leaves are sorted by rising weight */
    int at = n - 1;
    for (int l = 1; l <= max_len; ++l)
        for (int k = 0; k < count[l]; ++k) len[order[at--]] = (uint8_t)l;
}

inline uint32_t bitrev_host(uint32_t v, int n)
{
    uint32_t r = 0;
    for (int k = 0; k < n; ++k) r |= ((v >> k) & 1u) << (n - 1 - k);
    return r;
}
/* canonical codes (RFC 1951, 3.2.2) of the lengths, bit-reversed for an LSB-first stream: out[s] = code | len << 16 */
inline void canonical_codes(const uint8_t *len, int n, uint32_t *out)
{
    uint32_t count[16] = {0}, next[16] = {0};
    for (int s = 0; s < n; ++s) count[len[s]]++;
    count[0] = 0;
    uint32_t code = 0;
    for (int l = 1; l < 16; ++l) { code = (code + count[l - 1]) << 1; next[l] = code; }
    for (int s = 0; s < n; ++s) out[s] = len[s] ? (bitrev_host(next[len[s]]++, len[s]) | ((uint32_t)len[s] << 16)) : 0u;
}

/* the plan of given symbol counts (zero counts are raised to one: every symbol stays encodable) */
inline void deflate_plan_counts(const uint64_t *lf, const uint64_t *df, DeflatePlan &P);

/* first guess from a sample of the bin's bytes (n may be 0: a flat code): literal counts from the sample, made-up counts
 * for the length and distance symbols */
inline void deflate_plan(const uint8_t *sample, uint64_t n, DeflatePlan &P)
{
    uint64_t lf[286], df[30];
    for (int s = 0; s < 256; ++s) lf[s] = 1;
    for (uint64_t k = 0; k < n; ++k) lf[sample[k]] += 16;
    const uint64_t total = 16 * n + 256;
    lf[256] = 1;                                                    /* end of block: once per member */
    /* matches: about one token in six on SAM text, short ones most often */
    for (int s = 257; s < 286; ++s) lf[s] = std::max<uint64_t>(1, total / 6 / (uint64_t)(4 + (s - 257) * (s - 257) / 4) / 4);
    for (int s = 0; s < 30; ++s) df[s] = 1 + (uint64_t)(s >= 8 ? 8 : 1) * (uint64_t)(s >= 16 ? 2 : 1);       /* distances: the far ones are the common ones */
    deflate_plan_counts(lf, df, P);
}
/* second guess: the tokens k_bgzf_deflate made of the bin's first members with the first plan (hist: 286 + 30 counts) */
inline void deflate_plan_hist(const uint32_t *hist, DeflatePlan &P)
{
    uint64_t lf[286], df[30];
    for (int s = 0; s < 286; ++s) lf[s] = 4ull * hist[s] + 1;
    for (int s = 0; s < 30; ++s) df[s] = 4ull * hist[286 + s] + 1;
    deflate_plan_counts(lf, df, P);
}
inline void deflate_plan_counts(const uint64_t *lf, const uint64_t *df, DeflatePlan &P)
{
    uint8_t ll[286], dl[30];
    huffman_lengths(lf, 286, 15, ll);
    huffman_lengths(df, 30, 15, dl);
    canonical_codes(ll, 286, P.lit);
    canonical_codes(dl, 30, P.dist);
    /* header bits: BFINAL 1, BTYPE 10, HLIT 29, HDIST 29, HCLEN 15; the code-length code gives symbols 0..15 four bits
     * each and leaves out 16, 17, 18; then the 316 lengths, one 4-bit symbol each */
    memset(P.hdr_words, 0, sizeof P.hdr_words);
    uint32_t bit = 16;
    auto put = [&](uint32_t v, int nb) {
        for (int k = 0; k < nb; ++k, ++bit) if ((v >> k) & 1u) P.hdr_words[bit >> 5] |= 1u << (bit & 31);
    };
    put(1, 1); put(2, 2); put(29, 5); put(29, 5); put(15, 4);
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    for (int k = 0; k < 19; ++k) put(order[k] < 16 ? 4u : 0u, 3);
    for (int s = 0; s < 286; ++s) put(bitrev_host(ll[s], 4), 4);
    for (int s = 0; s < 30; ++s) put(bitrev_host(dl[s], 4), 4);
    P.hdr_bits = bit;
    P.pad = 0;
}

#if defined(__CUDACC__)
constexpr int DEF_WARPS = 4;
#ifndef XM_DEF_OCC
#define XM_DEF_OCC 5
#endif
struct DeflateSmem {
    uint16_t htab[DEF_WARPS][1 << DEF_HASH_BITS];
    uint32_t stage[DEF_WARPS][68];
    uint32_t lit[286], dist[30];
    uint32_t crc[256];
    uint32_t op[DEF_WARPS][32], slice[DEF_WARPS][32];
    uint32_t hist[316];              /* tokens by symbol, when the launch collects them */
};

__device__ __forceinline__ uint32_t def_load4(const uint32_t *W, uint32_t off)
{
    const uint32_t w0 = W[off >> 2], w1 = W[(off >> 2) + 1];
    return __funnelshift_r(w0, w1, (off & 3u) * 8u);
}

/* input bytes per member for a bin of n bytes: DEF_IN_MAX when that still gives every SM several members (a member is one
 * warp's work and takes milliseconds whatever else runs), finer otherwise; a multiple of 256 */
inline uint32_t deflate_member_bytes(uint64_t n, uint32_t sm_count)
{
    const uint64_t want = (uint64_t)sm_count * 40;                /* two waves of 20 warps per SM */
    uint64_t per = (n + want - 1) / want;
    per = (per + 255) & ~255ull;
    return (uint32_t)std::min<uint64_t>(DEF_IN_MAX, std::max<uint64_t>(DEF_IN_MIN, per));
}
inline uint32_t deflate_slot_bytes(uint32_t in_per) { return in_per + 256 + 256; }

/* src: 4-byte aligned, readable 8 bytes past its end.  member m covers [m * in_per, ...); its gzip member is built in
 * slot + m * slot_bytes and its size goes to sizes[m]. */
__global__ void __launch_bounds__(DEF_WARPS * 32, XM_DEF_OCC)
k_bgzf_deflate(const uint8_t *src, uint64_t n_total, uint32_t n_members, uint32_t in_per, uint32_t slot_bytes, const DeflatePlan *plan, uint8_t *slot,
               uint32_t *sizes, uint32_t *hist)
{
    __shared__ DeflateSmem S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t k = threadIdx.x; k < 286; k += blockDim.x) S.lit[k] = plan->lit[k];
    if (threadIdx.x < 30) S.dist[threadIdx.x] = plan->dist[threadIdx.x];
    for (uint32_t k = threadIdx.x; k < 256; k += blockDim.x) S.crc[k] = crc_table_entry(k);
    for (uint32_t k = threadIdx.x; k < 316; k += blockDim.x) S.hist[k] = 0;
    __syncthreads();
    const uint32_t m = blockIdx.x * DEF_WARPS + (uint32_t)warp;
    if (m >= n_members) { if (hist) __syncthreads(); return; }
    const uint64_t lo64 = (uint64_t)m * in_per;
    const uint32_t n = (uint32_t)((n_total - lo64) < in_per ? (n_total - lo64) : in_per);
    const uint8_t *in = src + lo64;
    const uint32_t *W = (const uint32_t *)in;                      /* in_per is a multiple of 4 */
    uint8_t *out = slot + (uint64_t)m * slot_bytes;
    uint32_t *OW = (uint32_t *)out;
    uint16_t *htab = S.htab[warp];
    uint32_t *stage = S.stage[warp];
    for (int k = lane; k < (1 << DEF_HASH_BITS); k += 32) htab[k] = 0xffff;
    for (int k = lane; k < 68; k += 32) stage[k] = 0;

    /* the stream starts at word 4 of the member (byte 16): BSIZE's two bytes, then the block header */
    const uint32_t hdr_bits = plan->hdr_bits;
    uint32_t wpos = 4 + (hdr_bits >> 5);                            /* next word of the member to be stored */
    uint32_t carry_bits = hdr_bits & 31u;
    for (uint32_t k = (uint32_t)lane; k < (hdr_bits >> 5); k += 32) OW[4 + k] = plan->hdr_words[k];
    __syncwarp();
    if (lane == 0) stage[0] = plan->hdr_words[hdr_bits >> 5];
    __syncwarp();
    const uint32_t room = slot_bytes < DEF_MEMBER_MAX ? slot_bytes : DEF_MEMBER_MAX;
    const uint32_t word_limit = (room - 8 - 160) / 4;               /* room for a step (32 tokens), EOB, the trailer */
    bool overflow = false;
    uint32_t cu = 0;                                                /* first position without a token */

    for (uint32_t base = 0; base < n; base += 32) {
        if (cu >= base + 32) continue;                              /* inside a long match */
        if (wpos >= word_limit) { overflow = true; break; }
        const uint32_t p = base + (uint32_t)lane;
        const bool valid = p < n, can = p + 4 <= n;
        uint32_t v = 0, hsh = 0, cand = 0xffff;
        if (can) { v = def_load4(W, p); hsh = (v * 2654435761u) >> (32 - DEF_HASH_BITS); cand = htab[hsh]; }
        __syncwarp();
        if (can) htab[hsh] = (uint16_t)p;
        uint32_t mlen = 0, mdist = 0;
        if (can && p >= cu) {
            const uint32_t maxl = n - p < (uint32_t)DEF_MAX_MATCH ? n - p : (uint32_t)DEF_MAX_MATCH;
            uint32_t c = cand;
            if (!(c != 0xffff && c < p && p - c <= 32768u && def_load4(W, c) == v)) c = (p >= 1 && def_load4(W, p - 1) == v) ? p - 1 : 0xffffffffu;
            if (c != 0xffffffffu) {
                uint32_t l = 4;
                while (l + 4 <= maxl && def_load4(W, c + l) == def_load4(W, p + l)) l += 4;
                while (l < maxl && in[c + l] == in[p + l]) ++l;
                mlen = l; mdist = p - c;
            }
        }
        /* what the match would cost, and what its bytes cost as literals: the literal bits of this step's bytes are summed
         * by a warp scan, a match that runs past the step is charged at the average of its part inside.  DNA letters get
         * two- or three-bit codes from the bin's plan: a four-letter match 30 000 bytes back (26 bits) must not replace them */
        uint32_t b1 = 0, n1 = 0, b2 = 0, n2 = 0, h_lsym = 0, h_dsym = 0;
        const uint32_t le0 = valid ? S.lit[can ? (v & 0xffu) : (uint32_t)in[p]] : 0u;
        if (mlen >= (uint32_t)DEF_MIN_MATCH) {
            const uint32_t l = mlen - 3;
            uint32_t lsym, lex = 0, lexv = 0;
            if (mlen == 258) lsym = 28;
            else if (l < 8) lsym = l;
            else { lex = (uint32_t)(29 - __clz((int)l)); lsym = 4 * lex + 4 + ((l >> lex) & 3u); lexv = l & ((1u << lex) - 1u); }
            const uint32_t le = S.lit[257 + lsym], lb = le >> 16;
            b1 = (le & 0xffffu) | (lexv << lb); n1 = lb + lex; h_lsym = lsym;
            const uint32_t d = mdist - 1;
            uint32_t dsym, dex = 0, dexv = 0;
            if (d < 4) dsym = d;
            else { dex = (uint32_t)(30 - __clz((int)d)); dsym = 2 * dex + 2 + ((d >> dex) & 1u); dexv = d & ((1u << dex) - 1u); }
            const uint32_t de = S.dist[dsym], db = de >> 16;
            b2 = (de & 0xffffu) | (dexv << db); n2 = db + dex; h_dsym = dsym;
        }
        {
            uint32_t lc = le0 >> 16;                               /* inclusive scan of the literal bits */
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, lc, o); if (lane >= o) lc += y; }
            const uint32_t inside = mlen < 32u - (uint32_t)lane ? mlen : 32u - (uint32_t)lane;       /* bytes of the match inside the step */
            const uint32_t upto = __shfl_sync(0xffffffffu, lc, (lane + (int)(inside ? inside : 1u) - 1) & 31);
            if (mlen >= (uint32_t)DEF_MIN_MATCH) {
                const uint32_t lit_in = upto - (lc - (le0 >> 16));
                const uint32_t lit_est = lit_in * mlen / inside;
                if (n1 + n2 + 2u >= lit_est) mlen = 0;
            }
        }
        /* greedy selection in position order */
        const uint32_t M = __ballot_sync(0xffffffffu, mlen >= (uint32_t)DEF_MIN_MATCH);
        bool covered = p < cu, selected = false;
        uint32_t cur = cu > base ? cu : base;
        for (;;) {
            const uint32_t off = cur - base;
            if (off >= 32) break;
            const uint32_t cl = M & (0xffffffffu << off);
            if (!cl) break;
            const int L = __ffs((int)cl) - 1;
            const uint32_t ll = __shfl_sync(0xffffffffu, mlen, L);
            if (lane == L) selected = true;
            cur = base + (uint32_t)L + ll;
            if (lane > L && p < cur) covered = true;
        }
        const uint32_t step_end = base + 32 < n ? base + 32 : n;
        cu = cur > step_end ? cur : step_end;
        /* the tokens' bits */
        if (!selected) {
            b2 = 0; n2 = 0;
            if (valid && !covered) { b1 = le0 & 0xffffu; n1 = le0 >> 16; }
            else { b1 = 0; n1 = 0; }
        }
        if (hist) {
            if (selected) { atomicAdd(&S.hist[257 + h_lsym], 1u); atomicAdd(&S.hist[286 + h_dsym], 1u); }
            else if (valid && !covered) atomicAdd(&S.hist[can ? (v & 0xffu) : (uint32_t)in[p]], 1u);
        }
        const uint32_t nb = n1 + n2;
        uint32_t inc = nb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
        const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
        uint32_t pos = carry_bits + inc - nb;
        if (n1) {
            atomicOr(&stage[pos >> 5], b1 << (pos & 31u));
            if ((pos & 31u) + n1 > 32u) atomicOr(&stage[(pos >> 5) + 1], b1 >> (32u - (pos & 31u)));
            pos += n1;
        }
        if (n2) {
            atomicOr(&stage[pos >> 5], b2 << (pos & 31u));
            if ((pos & 31u) + n2 > 32u) atomicOr(&stage[(pos >> 5) + 1], b2 >> (32u - (pos & 31u)));
        }
        __syncwarp();
        const uint32_t bits = carry_bits + tot, full = bits >> 5;       /* at most 31 + 32 * 48 bits: 49 words */
        const uint32_t s0 = (uint32_t)lane < full ? stage[lane] : 0u, s1 = (uint32_t)lane + 32u < full ? stage[lane + 32] : 0u;
        const uint32_t last = stage[full];
        __syncwarp();
        if ((uint32_t)lane < full) OW[wpos + lane] = s0;
        if ((uint32_t)lane + 32u < full) OW[wpos + lane + 32] = s1;
        stage[lane] = 0; stage[lane + 32] = 0;
        if (lane < 4) stage[64 + lane] = 0;
        __syncwarp();
        if (lane == 0) stage[0] = last;
        __syncwarp();
        wpos += full;
        carry_bits = bits & 31u;
    }

    uint32_t total;                                                 /* bytes of the member */
    if (!overflow) {
        /* end of block, the last partial word */
        const uint32_t e = S.lit[256];
        uint32_t w0 = stage[0] | ((e & 0xffffu) << carry_bits), w1 = carry_bits + (e >> 16) > 32u ? (e & 0xffffu) >> (32u - carry_bits) : 0u;
        const uint32_t bits = carry_bits + (e >> 16);
        if (lane == 0) { OW[wpos] = w0; OW[wpos + 1] = w1; }
        const uint32_t end_byte = wpos * 4 + (bits + 7) / 8;        /* first byte behind the DEFLATE data */
        total = end_byte + 8;
        if (total > DEF_MEMBER_MAX || end_byte - 18 >= n + 5) overflow = true;      /* a stored block is smaller */
    }
    __syncwarp();
    if (overflow) {
        /* stored: BFINAL 1, BTYPE 00, LEN, ~LEN, the bytes */
        if (lane == 0) { out[18] = 1; out[19] = (uint8_t)n; out[20] = (uint8_t)(n >> 8); out[21] = (uint8_t)~n; out[22] = (uint8_t)(~n >> 8); }
        for (uint32_t k = (uint32_t)lane; k < n; k += 32) out[23 + k] = in[k];
        total = 18 + 5 + n + 8;
    }
    __syncwarp();
    /* CRC-32 of the input by lanes */
    uint32_t lo, hi, L;
    crc_slice(n, lane, lo, hi, L);
    const int first = __ffs((int)__ballot_sync(0xffffffffu, hi > lo)) - 1;
    S.slice[warp][lane] = hi > lo ? crc_feed(S.crc, lane == first ? 0xffffffffu : 0u, in + lo, hi - lo) : 0u;
    S.op[warp][lane] = crc_zeros(S.crc, 1u << lane, L);
    __syncwarp();
    if (lane == 0) {
        const uint32_t crc = n ? crc_join(S.slice[warp], S.op[warp], first) : 0u;
        const uint8_t head[16] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 'B', 'C', 2, 0};
        for (int k = 0; k < 16; ++k) out[k] = head[k];
        out[16] = (uint8_t)((total - 1) & 0xffu); out[17] = (uint8_t)((total - 1) >> 8);
        uint8_t *t = out + total - 8;
        for (int k = 0; k < 4; ++k) { t[k] = (uint8_t)(crc >> (8 * k)); t[4 + k] = (uint8_t)(n >> (8 * k)); }
        sizes[m] = total;
    }
    if (hist) {
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < 316; k += blockDim.x) if (S.hist[k]) atomicAdd(&hist[k], S.hist[k]);
    }
}

/* exclusive prefix of sizes[0..n) into offs[0..n], one CTA */
__global__ void k_bgzf_offsets(const uint32_t *sizes, uint32_t n, unsigned long long *offs)
{
    __shared__ unsigned long long part[1024];
    const uint32_t t = threadIdx.x, per = (n + 1023u) / 1024u;
    unsigned long long s = 0;
    for (uint32_t k = t * per; k < n && k < (t + 1) * per; ++k) s += sizes[k];
    part[t] = s;
    __syncthreads();
    if (t == 0) { unsigned long long run = 0; for (int k = 0; k < 1024; ++k) { const unsigned long long v = part[k]; part[k] = run; run += v; } offs[n] = run; }
    __syncthreads();
    unsigned long long run = part[t];
    for (uint32_t k = t * per; k < n && k < (t + 1) * per; ++k) { offs[k] = run; run += sizes[k]; }
}

/* member m from its slot to dst + offs[m]: one warp per member */
__global__ void k_bgzf_pack(const uint8_t *slot, uint32_t slot_bytes, const uint32_t *sizes, const unsigned long long *offs, uint32_t n_members, uint8_t *dst)
{
    const uint32_t m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= n_members) return;
    const uint8_t *s = slot + (uint64_t)m * slot_bytes;
    uint8_t *d = dst + offs[m];
    const uint32_t n = sizes[m];
    /* destination-aligned 16-byte stores; the source words come by byte (L2-resident, written a moment ago) */
    const uint32_t head = (uint32_t)((16 - ((uintptr_t)d & 15)) & 15);
    for (uint32_t k = (uint32_t)lane; k < head && k < n; k += 32) d[k] = s[k];
    if (n > head) {
        const uint32_t body = (n - head) / 16;
        const uint32_t sh = head & 3u;                             /* the slot is 4-byte aligned: source offset of the body mod 4 */
        const uint32_t *SW = (const uint32_t *)(s + (head - sh));
        for (uint32_t k = (uint32_t)lane; k < body; k += 32) {
            uint32_t w[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) w[j] = SW[4 * k + j];
            uint4 v;
            v.x = __funnelshift_r(w[0], w[1], sh * 8u); v.y = __funnelshift_r(w[1], w[2], sh * 8u);
            v.z = __funnelshift_r(w[2], w[3], sh * 8u); v.w = __funnelshift_r(w[3], w[4], sh * 8u);
            *(uint4 *)(d + head + 16u * k) = v;
        }
        for (uint32_t k = head + 16u * body + (uint32_t)lane; k < n; k += 32) d[k] = s[k];
    }
}
#endif

}  // namespace xm
