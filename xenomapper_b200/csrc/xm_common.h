/*
 * xm_common.h -- types shared by the CUDA kernels, the C-ABI runtime and the
 * CPU emulation harness used by the tests (tests/emu).
 */
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define XM_HD __host__ __device__ __forceinline__
#define XM_COLD __host__ __device__ __noinline__      /* rare paths: kept out of line so the hot loop fits the instruction cache */
#else
#define XM_HD inline
#define XM_COLD inline
#endif

#if defined(__CUDA_ARCH__)
#define XM_DEVICE_PASS 1
#else
#define XM_DEVICE_PASS 0
#endif

#if !defined(__CUDACC__)
/* host-only builds (CPU emulation harness): the CUDA vector type the tile code uses */
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
#endif

namespace xm {

/* states / bins: reference output argument order, xm.py:291-297 */
enum { PS = 0, SS = 1, PM = 2, SM = 3, UA = 4, UR = 5, NO_BIN = 7 };
enum { MODE_SE = 0, MODE_PE_LIBERAL = 1, MODE_PE_CONSERVATIVE = 2 };
enum { SCORE_AS_XS = 0, SCORE_AS_ZS = 1, SCORE_CIGAR_NM = 2 };

/* -inf of the reference (missing tag, xm.py:188) */
constexpr int32_t SCORE_ABSENT = INT32_MIN;

/* per-line flags.  The low six travel in SCompact::meta. */
enum : uint32_t {
    F_DIRTY = 1,    /* output differs from the raw bytes: whitespace other than single tabs, CR, missing final newline */
    F_AS_DUP = 2,   /* two tokens match the AS lookup: ValueError xm.py:190 */
    F_AS_NUM = 4,   /* AS value is not a plain 31-bit decimal integer: host diagnoses ValueError vs unsupported */
    F_XS_DUP = 8,
    F_XS_NUM = 16,
    F_TEXT = 32,    /* non-ASCII byte, lone CR, or a line too long for the meta word */
    F_BLANK = 64,   /* no tokens: stops the walk, xm.py:105 */
    F_FAR = 128     /* line not fully inside the staged window */
};
constexpr uint32_t META_LEN_BITS = 24;
constexpr uint32_t META_LEN_MASK = (1u << META_LEN_BITS) - 1;
/* meta bit 30 (row kernels, xm_emit.cuh): this record's QNAME equals the QNAME of the line before it -- the pair
 * predicate of xm.py:402, decided by the scan where both lines are at hand (hash, then bytes) */
constexpr uint32_t META_SAME = 1u << 30;
constexpr uint32_t META_FLAGS = 0x3fu << META_LEN_BITS;

/* error word: (record index << 8) | (code << 2) | (previous-record << 1) | stream; atomicMin keeps the first */
enum { EC_ASSERT = 1, EC_TEXT = 2, EC_DUP = 3, EC_NUM_AS = 4, EC_NUM_XS = 5 };
constexpr unsigned long long NO_ERROR = ~0ull;

struct StreamBuf {
    const uint8_t *p;
    uint64_t len;
};

/* what the secondary-stream scan leaves behind, one entry per yielded record */
struct SCompact {
    uint64_t *start;     /* [n+1] byte offset of the record; [n] = offset just past the last one */
    uint4 *rec;          /* [n] x = AS, y = XS (SCORE_ABSENT when missing), z,w = 64-bit QNAME hash */
    uint32_t *meta;      /* [n] emitted length (24 bits) | flags << 24 */
};

struct Globals {
    unsigned long long counts[36];
    unsigned long long n_stream[2];   /* records before EOF / first blank line: [0] primary, [1] secondary */
    unsigned long long err;           /* first error word, NO_ERROR if none */
    unsigned long long out_len[6];
    unsigned long long bytes_in[2];
    unsigned long long end_off[2];    /* byte offset just past the last counted record of each stream */
    unsigned long long limit_off;     /* byte offset of primary record `limit` (error re-run: locates the failing line) */
    unsigned int overflow;            /* a tile held more lines than the geometry allows: relaunch smaller */
    unsigned int reserved_[2];        /* (was a tile ticket: tiles are blockIdx.x, see DESIGN section 3 on dispatch order) */
    unsigned int pad;
    unsigned long long phase[48];     /* XM_PHASE_TIMING builds: thread-0 clock cycles per tile phase, summed over tiles */
};

/* decoupled look-back descriptors */
constexpr unsigned long long C1_AGG = 1ull << 62, C1_INC = 2ull << 62, C1_STOP = 1ull << 61;
constexpr unsigned long long C1_COUNT = (1ull << 61) - 1;
constexpr int C2_SLOTS = 8;           /* six bins + raw primary bytes + spare (per-tile sums; the chain carries the six bins) */
constexpr unsigned long long C2_AGG = 1ull << 62, C2_INC = 2ull << 62, C2_VAL = (1ull << 62) - 1;

struct ScanArgs {
    StreamBuf S;
    SCompact sc;
    uint64_t sc_cap;
    unsigned long long *chain1;       /* [ntiles] */
    Globals *g;
    uint32_t ntiles;
    int32_t score_src, skip;
    int32_t stream_id;                /* which n_stream slot this scan owns */
    uint32_t debug;
    int32_t want_same;                /* mark rows whose QNAME repeats the previous line's (META_SAME): paired walks over rows */
    int32_t count_only;               /* record counts and row offsets only: the score tokens are not parsed */
    uint64_t start_bias;              /* added to every row's byte offset (sharded walks: offsets count from the shard allocation's first byte) */
    int32_t short_lines;              /* lines of less than ~300 bytes on average: the span kernels take spans of half the size (64 lines per span still fit) */
    int32_t pad_;
};

/* The row walk (xm_emit.cuh): both streams have been scanned into compact rows; record i of the walk is row i of
 * both.  k_size sizes the six bins tile by tile, k_prefix turns the tile totals into bases, k_emit copies. */
constexpr int EM_WARPS = 8, EM_PER_WARP = 64, EM_TILE = EM_WARPS * EM_PER_WARP;
struct EmitArgs {
    StreamBuf P, S;
    SCompact rp, rs;                  /* rows of the primary / secondary stream, already offset to the walk's record 0 */
    uint64_t n;                       /* records to walk */
    int32_t mode, skip, halo;
    int32_t exact_names;              /* behind equal QNAME hashes compare the bytes too (XM_DEBUG_EXACT_NAMES) */
    long long thr;
    uint32_t enabled;
    Globals *g;
    unsigned long long *tile_tot;     /* [ntiles][C2_SLOTS]: bytes per bin of each tile, then (k_prefix) where each tile starts */
    uint32_t ntiles;
    uint8_t *out[6];
    uint64_t out_cap[6];
};

struct ClassifyArgs {
    StreamBuf P, S;
    SCompact sc;
    uint64_t sc_cap;                  /* rows the compact arrays hold: records at or beyond it are not looked up (the host grows the arrays and walks again) */
    unsigned long long *chain1;       /* [ntiles] */
    unsigned long long *chain2;       /* [ntiles][C2_SLOTS] status << 62 | bytes, one look-back chain per bin */
    Globals *g;
    uint32_t ntiles;
    int32_t mode, score_src, skip;
    long long thr;                    /* AS > min_score  <=>  AS >= thr */
    uint32_t enabled;
    uint64_t limit;                   /* records >= limit are dropped (error re-run) */
    uint64_t *p_start;                /* [sc_cap] byte offset of every yielded primary record, or null (chunked walks carry the tail over) */
    uint64_t p_start_cap;
    int32_t halo;                     /* record 0 was the last record of the previous chunk: context only, not classified again */
    uint8_t *out[6];
    uint64_t out_cap[6];
    uint32_t debug;
    int32_t short_lines;              /* as in ScanArgs */
};

constexpr uint32_t DBG_FORCE_GENERIC = 1, DBG_SMALL_TILES = 2, DBG_ROWS = 4, DBG_EXACT_NAMES = 8;

/* tile geometry: one line per thread */
template <int TILE_, int HALO_, int THREADS_>
struct Cfg {
    static constexpr int TILE = TILE_;        /* bytes of the stream one CTA owns (lines are owned by their first byte) */
    static constexpr int HALO = HALO_;        /* bytes staged before and after the tile */
    static constexpr int THREADS = THREADS_;
    static constexpr int WIN = TILE + 2 * HALO;
    static constexpr int NW = WIN / 32 + 2;             /* 32-bit mask words: the window, the virtual newline bit, one spare */
    static constexpr int WPT = (NW + THREADS - 1) / THREADS;   /* mask words per thread in the index pass */
    static constexpr int LCAP = THREADS - 1;            /* owned lines a tile can hold; the last thread serves the halo line */
    static_assert(TILE % 32 == 0 && HALO % 32 == 0 && HALO > 0, "geometry");
    static_assert(WIN + 1 < 65535, "line starts are kept as 16-bit window offsets");
    static_assert(THREADS <= 512 && THREADS % 32 == 0, "per-warp scratch holds 16 warps");
};
/* Production geometry, from the sweeps in profiles/r01_geometry_sweep.md: the time a tile takes is dominated by
 * latencies that do not depend on its size (look-back waits, one line per thread in the parse), so throughput grows
 * with the bytes a SM holds in flight: the largest window three CTAs fit into 227 KB of shared memory, and enough
 * threads that a tile of ~125 lines still has two threads per line. */
#ifndef XM_BIG_THREADS
#define XM_BIG_THREADS 320
#endif
#ifndef XM_BIG_TILE
#define XM_BIG_TILE 47104
#endif
#ifndef XM_BIG_HALO
#define XM_BIG_HALO 1024
#endif
#ifndef XM_BIG_OCC
#define XM_BIG_OCC 3
#endif
using CfgBig = Cfg<XM_BIG_TILE, XM_BIG_HALO, XM_BIG_THREADS>;
using CfgSmall = Cfg<480, 512, 256>;          /* LCAP >= TILE/2: cannot lose the stop (non-blank lines need 2 bytes) */

}  // namespace xm
