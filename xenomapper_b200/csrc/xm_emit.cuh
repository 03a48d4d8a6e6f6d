/*
 * xm_emit.cuh -- the walk over rows (included by xm_kernels.cu).
 *
 * Both streams have been scanned by k_scan2 into compact rows (byte offset,
 * AS, XS, QNAME hash, emitted length; 28 bytes per record), so record i of the
 * lockstep reader (xm.py:95-118) is row i of both, and nothing below looks at
 * the text except to copy it:
 *
 *   k_size    joins row i of the two streams (QNAME assert xm.py:106: hashes,
 *             then the bytes), decides the category (xm.py:258-289, pair chains
 *             xm.py:423-448 / 521-550), adds up the bytes every tile of 512
 *             records sends to each of the six bins, counts the categories;
 *   k_prefix  turns the tile totals into the offset at which each tile starts
 *             in each bin (one CTA; the totals are a few MB);
 *   k_emit    decides again (forty bytes of rows per record), places every line
 *             with warp scans behind its tile's base and copies it, global to
 *             global, neighbouring lines of one bin as one run.
 *
 * No look-back chain, no wait between tiles: the order of the output is fixed
 * by k_prefix before a byte is copied, and a tile's copy depends on nothing
 * but its own rows.  The price is a second read of the emitted lines from HBM
 * (the fused k_classify2 finds them in L2).  Clean inputs only, like the scans
 * that feed it: any row with a flag, any QNAME mismatch raises Globals::pad and
 * the host walks the streams with the exact pair.
 */
#pragma once

namespace xm {

struct RowDec {
    uint32_t key;        /* histogram slot, 36 = none */
    uint32_t bin;        /* NO_BIN = nothing emitted */
    uint32_t plen, slen; /* bytes from the primary / secondary text */
    uint64_t psrc, ssrc; /* where they start in the streams */
};

/* the QNAMEs of the lines at P.p + ps and S.p + ss are equal byte for byte up to and including the separator
 * that ends the primary one.  Both lines' first 64 aligned bytes are requested at once (eight 16-byte loads in
 * flight per lane, one round trip to memory) and parked in the lane's slot of shared memory, from where the compare
 * takes its words at the lanes' own offsets; names that do not end inside those bytes take the loop over global
 * memory. */
constexpr int QN_SLOT = 144;            /* bytes per lane: 2 x 64 and a pad that spreads the lanes over the banks */
__device__ __noinline__ bool rows_names_equal_long(const StreamBuf &P, uint64_t ps, const StreamBuf &S, uint64_t ss)
{
    const uint32_t *pw = (const uint32_t *)(P.p + (ps & ~3ull)), *sw = (const uint32_t *)(S.p + (ss & ~3ull));
    const uint32_t shp = (uint32_t)(ps & 3u) * 8u, shs = (uint32_t)(ss & 3u) * 8u;
    uint32_t plo = __ldg(pw), slo = __ldg(sw);
#pragma unroll 1
    for (int j = 1; j <= 64; ++j) {
        const uint32_t phi = __ldg(pw + j), shi = __ldg(sw + j);
        const uint32_t a = __funnelshift_r(plo, phi, shp), b = __funnelshift_r(slo, shi, shs);
        const uint32_t t = ctrl_mask(a);
        if (t) {
            const uint32_t keep = 0xffffffffu >> (32 - __ffs((int)t));       /* through the separator byte */
            return ((a ^ b) & keep) == 0u;
        }
        if (a != b) return false;
        plo = phi; slo = shi;
    }
    return false;       /* a QNAME of 256 bytes or more: the exact kernel compares it */
}
__device__ __forceinline__ bool rows_names_equal(const StreamBuf &P, uint64_t ps, const StreamBuf &S, uint64_t ss, uint8_t *slot)
{
    const uint8_t *pb = P.p + (ps & ~15ull), *sb = S.p + (ss & ~15ull);
    /* the buffers are readable to the next multiple of 16 past their end; blocks beyond that are not touched */
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool okp = (ps & ~15ull) + 16u * (uint64_t)k < P.len, oks = (ss & ~15ull) + 16u * (uint64_t)k < S.len;
        v[k] = okp ? ld_src16(pb + 16 * k, false) : make_uint4(0, 0, 0, 0);
        v[4 + k] = oks ? ld_src16(sb + 16 * k, false) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) *(uint4 *)(slot + 16 * k) = v[k];
    __syncwarp();
    const uint32_t *pw = (const uint32_t *)slot + ((uint32_t)(ps & 15u) >> 2), *sw = (const uint32_t *)(slot + 64) + ((uint32_t)(ss & 15u) >> 2);
    const uint32_t shp = (uint32_t)(ps & 3u) * 8u, shs = (uint32_t)(ss & 3u) * 8u;
    uint32_t plo = pw[0], slo = sw[0];
    int verdict = -1;
#pragma unroll
    for (int j = 1; j <= 12; ++j) {
        const uint32_t phi = pw[j], shi = sw[j];
        const uint32_t a = __funnelshift_r(plo, phi, shp), b = __funnelshift_r(slo, shi, shs);
        const uint32_t t = ctrl_mask(a);
        if (verdict < 0) {
            if (t) verdict = (((a ^ b) & (0xffffffffu >> (32 - __ffs((int)t)))) == 0u) ? 1 : 0;
            else if (a != b) verdict = 0;
        }
        plo = phi; slo = shi;
    }
    if (verdict < 0) return rows_names_equal_long(P, ps, S, ss);
    return verdict == 1;
}

/* One batch of 32 consecutive records, one per lane: lane's record is `i` (valid = i < n).  `carry` hands the last
 * lane's state to the next batch of the same warp; the record before a warp's first is looked up by lane 0.
 * CHECK: also assert the QNAMEs and the row flags (k_size); otherwise trust them (k_emit runs only after k_size). */
struct RowCarry {
    int st;
    uint32_t pout, sout;
    uint64_t ps, ss;
    bool valid;
};

template <bool CHECK>
__device__ __forceinline__ RowDec rows_decide(const EmitArgs &a, uint64_t i, bool valid, bool first_batch, RowCarry &carry, bool &bad, uint8_t *slot = nullptr)
{
    const int lane = threadIdx.x & 31;
    const bool paired = a.mode != MODE_SE;
    RowDec d;
    d.key = 36; d.bin = NO_BIN; d.plen = 0; d.slen = 0; d.psrc = 0; d.ssrc = 0;
    int st = UA;
    uint32_t pout = 0, sout = 0, pmeta = 0;
    uint64_t ps = 0, ss = 0;
    if (valid) {
        const uint4 pr = a.rp.rec[i], sr = a.rs.rec[i];
        pmeta = a.rp.meta[i];
        const uint32_t smeta = a.rs.meta[i];
        ps = a.rp.start[i]; ss = a.rs.start[i];
        if (CHECK) {
            if (((pmeta | smeta) & META_FLAGS) || pr.z != sr.z || pr.w != sr.w) bad = true;       /* dirty or failing line, QNAME assert */
        }
        st = mapping_state((int32_t)pr.x, (int32_t)pr.y, (int32_t)sr.x, (int32_t)sr.y, a.thr);
        pout = pmeta & META_LEN_MASK; sout = smeta & META_LEN_MASK;
    }
    if (CHECK && a.exact_names) {
        /* the assert of xm.py:106 on the bytes (all lanes together: the slots are handed over with a warp barrier) */
        const bool eq = rows_names_equal(a.P, valid ? ps : 0, a.S, valid ? ss : 0, slot);
        if (valid && !eq) bad = true;
    }
    if (!paired) {
        if (valid && !(a.halo && i == 0)) {
            d.key = (uint32_t)st; d.bin = (uint32_t)st;
            const bool pside = st == PS || st == PM || st == UA || st == UR, sside = st == SS || st == SM || st == UR;
            d.plen = pside ? pout : 0u; d.slen = sside ? sout : 0u; d.psrc = ps; d.ssrc = ss;
        }
    } else {
        /* the record before: the lane below, the previous batch's last lane, or (first batch, lane 0) row i - 1 */
        int pst = __shfl_up_sync(0xffffffffu, st, 1);
        uint32_t ppout = __shfl_up_sync(0xffffffffu, pout, 1), psout = __shfl_up_sync(0xffffffffu, sout, 1);
        uint64_t pps = __shfl_up_sync(0xffffffffu, ps, 1), pss = __shfl_up_sync(0xffffffffu, ss, 1);
        bool pvalid = __shfl_up_sync(0xffffffffu, (int)valid, 1) != 0;
        if (lane == 0) {
            if (!first_batch) { pst = carry.st; ppout = carry.pout; psout = carry.sout; pps = carry.ps; pss = carry.ss; pvalid = carry.valid; }
            else {
                pvalid = valid && i > 0;
                if (pvalid) {
                    const uint4 pr = a.rp.rec[i - 1], sr = a.rs.rec[i - 1];
                    pst = mapping_state((int32_t)pr.x, (int32_t)pr.y, (int32_t)sr.x, (int32_t)sr.y, a.thr);
                    ppout = a.rp.meta[i - 1] & META_LEN_MASK; psout = a.rs.meta[i - 1] & META_LEN_MASK;
                    pps = a.rp.start[i - 1]; pss = a.rs.start[i - 1];
                }
            }
        }
        carry.st = __shfl_sync(0xffffffffu, st, 31); carry.pout = __shfl_sync(0xffffffffu, pout, 31); carry.sout = __shfl_sync(0xffffffffu, sout, 31);
        carry.ps = __shfl_sync(0xffffffffu, ps, 31); carry.ss = __shfl_sync(0xffffffffu, ss, 31);
        carry.valid = __shfl_sync(0xffffffffu, (int)valid, 31) != 0;
        /* a unit fires when this record carries the QNAME of the one before it (xm.py:402) */
        if (valid && !a.skip && (pmeta & META_SAME) && pvalid && i > 0) {
            d.key = (uint32_t)(pst * 6 + st);
            d.bin = (uint32_t)(a.mode == MODE_PE_CONSERVATIVE ? pair_bin_conservative(pst, st) : pair_bin_liberal(pst, st));
            const bool pside = d.bin == PS || d.bin == PM || d.bin == UA || d.bin == UR, sside = d.bin == SS || d.bin == SM || d.bin == UR;
            d.plen = pside ? ppout + pout : 0u; d.slen = sside ? psout + sout : 0u; d.psrc = pps; d.ssrc = pss;
            /* both lines of a side leave as one piece: they follow each other in the text */
            if (CHECK && (pps + ppout != ps || pss + psout != ss)) bad = true;
        }
    }
    if (d.bin != NO_BIN && !((a.enabled >> d.bin) & 1u)) { d.plen = 0; d.slen = 0; }
    return d;
}

/* ---- k_size ---------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(EM_WARPS * 32) k_size(const EmitArgs a)
{
    __shared__ unsigned long long s_tot[EM_WARPS][8];
    __shared__ __align__(16) uint8_t s_qn[EM_WARPS * 32][QN_SLOT];
    __shared__ uint32_t s_cnt[36];              /* the CTA's part of the histogram: one global atomic per category and CTA */
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t w0 = ((uint64_t)blockIdx.x * EM_WARPS + (uint64_t)warp) * EM_PER_WARP;
    if (threadIdx.x < 36) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    bool bad = false;
    RowCarry carry{UA, 0, 0, 0, 0, false};
    unsigned long long tot = 0;                 /* lane b < 6: this warp's bytes for bin b */
#pragma unroll
    for (int kb = 0; kb < EM_PER_WARP; kb += 32) {
        const uint64_t i = w0 + (uint64_t)(kb + lane);
        const bool valid = i < a.n;
        if (!__any_sync(0xffffffffu, valid)) break;
        const RowDec d = rows_decide<true>(a, i, valid, kb == 0, carry, bad, s_qn[threadIdx.x]);
        {
            const unsigned peers = __match_any_sync(0xffffffffu, d.key);
            if (d.key < 36u && lane == __ffs((int)peers) - 1) atomicAdd(&s_cnt[d.key], (uint32_t)__popc(peers));
        }
        const uint32_t bytes = d.plen + d.slen;
#pragma unroll
        for (uint32_t b = 0; b < 6; ++b) {
            const uint32_t v = d.bin == b ? bytes : 0u;
            if (__any_sync(0xffffffffu, v != 0u)) {
                const uint32_t sum = __reduce_add_sync(0xffffffffu, v);
                if (lane == (int)b) tot += sum;
            }
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) a.g->pad = 1u;
    if (lane < 8) s_tot[warp][lane] = lane < 6 ? tot : 0ull;
    __syncthreads();
    if (threadIdx.x < 8) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < EM_WARPS; ++w) t += s_tot[w][threadIdx.x];
        a.tile_tot[(size_t)blockIdx.x * C2_SLOTS + threadIdx.x] = t;
    }
    if (threadIdx.x >= 32 && threadIdx.x < 68) {
        const uint32_t v = s_cnt[threadIdx.x - 32];
        if (v) atomicAdd(&a.g->counts[threadIdx.x - 32], (unsigned long long)v);
    }
}

/* ---- k_prefix: tile totals -> tile bases, bin lengths (one CTA) --------------------------------------------- */
__global__ void __launch_bounds__(1024) k_prefix(unsigned long long *tile_tot, uint32_t ntiles, Globals *g)
{
    constexpr uint32_t NTHR = 1024u / 6u;                  /* 170 threads per bin; the last four threads only meet the barriers */
    __shared__ unsigned long long s_part[6][NTHR];
    const uint32_t b = threadIdx.x % 6u, t = threadIdx.x / 6u;
    const bool active = t < NTHR;
    const uint32_t per = (ntiles + NTHR - 1u) / NTHR;
    const uint32_t lo = t * per < ntiles ? t * per : ntiles, hi = lo + per < ntiles ? lo + per : ntiles;
    if (active) {
        unsigned long long sum = 0;
        for (uint32_t k = lo; k < hi; ++k) sum += tile_tot[(size_t)k * C2_SLOTS + b];
        s_part[b][t] = sum;
    }
    __syncthreads();
    if (t == 0) {
        unsigned long long run = 0;
        for (uint32_t k = 0; k < NTHR; ++k) { const unsigned long long v = s_part[b][k]; s_part[b][k] = run; run += v; }
        g->out_len[b] = run;
    }
    __syncthreads();
    if (active) {
        unsigned long long run = s_part[b][t];
        for (uint32_t k = lo; k < hi; ++k) {
            const unsigned long long v = tile_tot[(size_t)k * C2_SLOTS + b];
            tile_tot[(size_t)k * C2_SLOTS + b] = run;
            run += v;
        }
    }
}

/* ---- k_emit ---------------------------------------------------------------------------------------------- */
#ifndef XM_EMIT_OCC
#define XM_EMIT_OCC 3      /* 80 registers, no spills: 3.51 ms per 20 M records against 3.83 ms at 4 (64 registers) */
#endif
__global__ void __launch_bounds__(EM_WARPS * 32, XM_EMIT_OCC) k_emit(const EmitArgs a)
{
    __shared__ unsigned long long s_tot[EM_WARPS][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t w0 = ((uint64_t)blockIdx.x * EM_WARPS + (uint64_t)warp) * EM_PER_WARP;
    bool bad = false;
    RowCarry carry{UA, 0, 0, 0, 0, false};
    constexpr int NB = EM_PER_WARP / 32;
    RowDec D[NB];
    uint32_t off[NB];                           /* where the lane's bytes start inside the warp's part of its bin */
    uint32_t run_len[NB];                       /* bytes of the run of primary pieces that starts at this lane, 0 if none starts here */
    uint32_t wtot = 0;                          /* lane b < 6: the warp's running total of bin b */
#pragma unroll
    for (int q = 0; q < NB; ++q) {
        const uint64_t i = w0 + (uint64_t)(32 * q + lane);
        const bool valid = i < a.n;
        D[q] = rows_decide<false>(a, i, valid, q == 0, carry, bad);
        const RowDec &d = D[q];
        const uint32_t bytes = d.plen + d.slen;
        uint32_t o = 0;
#pragma unroll
        for (uint32_t b = 0; b < 6; ++b) {
            const uint32_t v = d.bin == b ? bytes : 0u;
            if (__any_sync(0xffffffffu, v != 0u)) {
                uint32_t x = v;
#pragma unroll
                for (int s = 1; s < 32; s <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, s); if (lane >= s) x += y; }
                const uint32_t before = __shfl_sync(0xffffffffu, wtot, (int)b);
                if (d.bin == b) o = before + x - v;
                const uint32_t add = __shfl_sync(0xffffffffu, x, 31);
                if (lane == (int)b) wtot += add;
            }
        }
        off[q] = o;
        /* primary pieces that continue each other in the text and in the bin leave as one run; lanes without a
         * primary piece do not break a run */
        const uint64_t pend_src = d.psrc + d.plen;
        const uint32_t pend_dst = o + d.plen;
        const uint32_t items = __ballot_sync(0xffffffffu, d.plen != 0u);
        const uint32_t below = items & ((1u << lane) - 1u);
        const int pl = below ? 31 - __clz((int)below) : 0;
        const uint64_t q_src = __shfl_sync(0xffffffffu, pend_src, pl);
        const uint32_t q_dst = __shfl_sync(0xffffffffu, pend_dst, pl), q_bin = __shfl_sync(0xffffffffu, d.bin, pl);
        /* (a lane with a secondary piece behind its primary one ends the run by itself: the bin's next piece starts
         * behind that secondary piece, not at q_dst) */
        const bool head = d.plen && !(below && q_bin == d.bin && q_src == d.psrc && q_dst == o);
        const uint32_t heads = __ballot_sync(0xffffffffu, head);
        const uint32_t above = heads & (lane == 31 ? 0u : (0xffffffffu << (lane + 1)));
        const uint32_t upto = above ? ((1u << (__ffs((int)above) - 1)) - 1u) : 0xffffffffu;
        const uint32_t mineq = items & upto;
        const int last = mineq ? 31 - __clz((int)mineq) : lane;
        const uint64_t run_end = __shfl_sync(0xffffffffu, pend_src, last);
        run_len[q] = head ? (uint32_t)(run_end - d.psrc) : 0u;
    }
    /* ---- where the warp's bytes start: the tile's base and the warps before this one ----------------------------- */
    if (lane < 8) s_tot[warp][lane] = lane < 6 ? (unsigned long long)wtot : 0ull;
    __syncthreads();
    unsigned long long wbase = 0;               /* lane b < 6 */
    if (lane < 6) {
        wbase = a.tile_tot[(size_t)blockIdx.x * C2_SLOTS + lane];
        for (int w = 0; w < warp; ++w) wbase += s_tot[w][lane];
    }
    /* ---- copy: the lanes that start a run or carry a secondary piece, one after the other ------------------------ */
#pragma unroll
    for (int q = 0; q < NB; ++q) {
        const RowDec &d = D[q];
        uint32_t todo = __ballot_sync(0xffffffffu, run_len[q] != 0u || d.slen != 0u);
#pragma unroll 1
        for (; todo; todo &= todo - 1) {
            const int L = __ffs((int)todo) - 1;
            const uint32_t bin = __shfl_sync(0xffffffffu, d.bin, L);
            const unsigned long long bb = __shfl_sync(0xffffffffu, wbase, (int)(bin < 6u ? bin : 0u));
            const uint32_t o = __shfl_sync(0xffffffffu, off[q], L);
            const uint32_t rl = __shfl_sync(0xffffffffu, run_len[q], L), sl = __shfl_sync(0xffffffffu, d.slen, L), pl = __shfl_sync(0xffffffffu, d.plen, L);
            if (rl) {
                const uint64_t src = __shfl_sync(0xffffffffu, d.psrc, L);
                const unsigned long long doff = bb + o;
                if (doff + rl <= a.out_cap[bin]) dev_copy_global(a.out[bin] + doff, a.P.p + src, rl);
            }
            if (sl) {
                const uint64_t src = __shfl_sync(0xffffffffu, d.ssrc, L);
                const unsigned long long doff = bb + o + pl;
                if (doff + sl <= a.out_cap[bin]) dev_copy_global(a.out[bin] + doff, a.S.p + src, sl);
            }
        }
    }
}

cudaError_t launch_size(const EmitArgs &a, cudaStream_t st)
{
    if (!a.ntiles) return cudaSuccess;
    k_size<<<a.ntiles, EM_WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_prefix(const EmitArgs &a, cudaStream_t st)
{
    k_prefix<<<1, 1024, 0, st>>>(a.tile_tot, a.ntiles, a.g);
    return cudaGetLastError();
}
cudaError_t launch_emit(const EmitArgs &a, cudaStream_t st)
{
    if (!a.ntiles) return cudaSuccess;
    k_emit<<<a.ntiles, EM_WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace xm
