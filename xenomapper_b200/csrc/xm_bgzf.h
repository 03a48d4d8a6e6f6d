/*
 * xm_bgzf.h -- BGZF output (host code): the six bins written as blocked gzip.
 *
 * The reference writes SAM text and leaves compression to a pipe into
 * `samtools view -bS` (README.md:138, xm.py:590-592).  After the classifier the
 * uncompressed bins are the next bottleneck, so the descriptor walk can
 * compress them itself -- an additive option, SAM text stays the default.
 * Format: the BGZF of the SAM/BAM specification, section 4.1: a series of gzip
 * members of at most 64 KiB of input, each with the "BC" extra field that holds
 * the member's size minus one, ended by the 28-byte empty member.  Any gzip
 * reader inflates the concatenation to exactly the SAM text; htslib can index
 * it.  This file is the host writer: headers, xm_bgzf_write, and the bins when XM_BGZF_DEFLATE=host (members deflated
 * side by side by a pool of host threads).  The bins of an XM_OUT_BGZF walk are normally deflated on the device:
 * xm_deflate.h.
 */
#pragma once
#include <stdint.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

namespace xm {

constexpr uint32_t BGZF_IN_MAX = 0xff00;           /* input bytes per member (htslib's choice: the member stays below 64 KiB) */
static const uint8_t BGZF_EOF[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

/* one member: header (18 bytes), raw deflate, CRC32, ISIZE.  dst holds at least 18 + deflateBound(n) + 8 bytes. */
inline bool bgzf_member(const uint8_t *src, uint32_t n, int level, uint8_t *dst, uint32_t dst_cap, uint32_t &out_n)
{
    static const uint8_t head[12] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0};
    memcpy(dst, head, 12);
    dst[12] = 'B'; dst[13] = 'C'; dst[14] = 2; dst[15] = 0;
    z_stream z;
    memset(&z, 0, sizeof z);
    if (deflateInit2(&z, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    z.next_in = const_cast<uint8_t *>(src); z.avail_in = n;
    z.next_out = dst + 18; z.avail_out = dst_cap - 18 - 8;
    const int r = deflate(&z, Z_FINISH);
    const uint32_t clen = (uint32_t)z.total_out;
    deflateEnd(&z);
    if (r != Z_STREAM_END) return false;
    const uint32_t total = 18 + clen + 8;
    if (total > 0x10000) return false;             /* cannot happen for n <= BGZF_IN_MAX: deflate never grows that much */
    dst[16] = (uint8_t)((total - 1) & 0xff); dst[17] = (uint8_t)((total - 1) >> 8);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, n);
    uint8_t *t = dst + 18 + clen;
    for (int k = 0; k < 4; ++k) { t[k] = (uint8_t)(crc >> (8 * k)); t[4 + k] = (uint8_t)(n >> (8 * k)); }
    out_n = total;
    return true;
}

/* [src, src + n) as BGZF members appended to `out`, deflated by `threads` threads */
inline bool bgzf_compress(const uint8_t *src, uint64_t n, int level, int threads, std::vector<uint8_t> &out)
{
    const uint64_t members = (n + BGZF_IN_MAX - 1) / BGZF_IN_MAX;
    if (!members) return true;
    const uint32_t slot = 18 + (uint32_t)compressBound(BGZF_IN_MAX) + 8 + 64;
    std::vector<uint8_t> tmp((size_t)(members * slot));
    std::vector<uint32_t> len((size_t)members, 0);
    bool ok = true;
    const int nt = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)threads, members));
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&, t] {
            for (uint64_t m = (uint64_t)t; m < members; m += (uint64_t)nt) {
                const uint64_t lo = m * BGZF_IN_MAX;
                const uint32_t k = (uint32_t)std::min<uint64_t>(BGZF_IN_MAX, n - lo);
                if (!bgzf_member(src + lo, k, level, tmp.data() + m * slot, slot, len[(size_t)m])) ok = false;
            }
        });
    for (auto &t : th) t.join();
    if (!ok) return false;
    uint64_t total = 0;
    for (auto v : len) total += v;
    const size_t at0 = out.size();
    out.resize(at0 + (size_t)total);
    uint64_t at = at0;
    for (uint64_t m = 0; m < members; ++m) { memcpy(out.data() + at, tmp.data() + m * slot, len[(size_t)m]); at += len[(size_t)m]; }
    return true;
}

inline int bgzf_level()
{
    const char *e = getenv("XM_BGZF_LEVEL");
    if (e && *e) { const int v = atoi(e); if (v >= 0 && v <= 9) return v; }
    return 1;              /* fast: the writer has to keep up with a GPU */
}

}  // namespace xm
