"""ctypes face of the CPU emulation of the tile programs (tests/emu/xm_emu.cpp).  Test scaffolding only."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emu", "xm_emu.cpp")
SO = os.path.join(HERE, "emu", "libxm_emu.so")
DEPS = [SRC] + [os.path.join(ROOT, "xenomapper_b200", "csrc", f) for f in
                ("xm_common.h", "xm_parse.h", "xm_tile.h", "xm_walk.h", "xm_stream.h", "xm_shard.h", "xm_inflate.h", "xm_bamchain.h", "xm_deflate.h")] + [os.path.join(ROOT, "include", "xenomapper_b200.h")]


class Opts(C.Structure):
    _fields_ = [("mode", C.c_int32), ("score_src", C.c_int32), ("skip_repeated", C.c_int32),
                ("enabled_bins", C.c_uint32), ("min_score", C.c_double)]


class Result(C.Structure):
    _fields_ = [("counts", C.c_uint64 * 36), ("n_records", C.c_uint64), ("out_len", C.c_uint64 * 6),
                ("bytes_in", C.c_uint64 * 2), ("status", C.c_int32), ("err_stream", C.c_int32),
                ("err_record", C.c_uint64), ("ms_scan", C.c_float), ("ms_classify", C.c_float),
                ("ms_total", C.c_float), ("n_launches", C.c_uint32),
                ("ms_kernel", C.c_float * 6), ("reserved", C.c_uint32 * 2)]


class ShardInfo(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("first_start", C.c_uint64), ("stop_at", C.c_uint64), ("end_off", C.c_uint64)]


class ShardStats(C.Structure):
    _fields_ = [("rec_lo", C.c_uint64), ("rec_hi", C.c_uint64), ("n_records_total", C.c_uint64),
                ("out_offset", C.c_uint64 * 6), ("out_total", C.c_uint64 * 6), ("sliver_bytes", C.c_uint64),
                ("sent_bytes", C.c_uint64), ("align_ms", C.c_float), ("index_ms", C.c_float), ("sliver_ms", C.c_float),
                ("walk_ms", C.c_float), ("comm_ms", C.c_float), ("total_ms", C.c_float), ("n_collectives", C.c_uint32),
                ("first_bad_rank", C.c_int32)]


class Xfer(C.Structure):
    _fields_ = [("peer", C.c_int), ("ptr", C.c_void_p), ("bytes", C.c_uint64)]


ALL_GATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64)
EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.POINTER(Xfer), C.c_int, C.POINTER(Xfer), C.c_int)

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
            subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                                   "-pthread", "-o", SO, SRC])
        _lib = C.CDLL(SO)
        _lib.xm_emu_classify.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.POINTER(Opts), C.c_uint32,
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(Result),
                                         C.c_char_p, C.c_size_t]
        _lib.xm_emu_classify.restype = C.c_int
        _lib.xm_emu_classify_stream.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.POINTER(Opts), C.c_uint32, C.c_uint64,
                                                C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(Result),
                                                C.c_char_p, C.c_size_t]
        _lib.xm_emu_classify_stream.restype = C.c_int
        _lib.xm_emu_index.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_uint64), C.POINTER(ShardInfo), C.c_char_p, C.c_size_t]
        _lib.xm_emu_index.restype = C.c_int
        _lib.xm_emu_classify_sharded.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.POINTER(Opts), C.c_int, C.c_int,
                                                 ALL_GATHER_FN, EXCHANGE_FN, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                                                 C.POINTER(Result), C.POINTER(ShardStats), C.c_char_p, C.c_size_t]
        _lib.xm_emu_classify_sharded.restype = C.c_int
        _lib.xm_emu_inflate.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        _lib.xm_emu_inflate.restype = C.c_int
        _lib.xm_emu_crc32.argtypes = [C.c_void_p, C.c_uint32]
        _lib.xm_emu_crc32.restype = C.c_uint32
        _lib.xm_emu_bam_chain.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64), C.c_uint64,
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_uint64, C.POINTER(C.c_uint64)]
        _lib.xm_emu_bam_chain.restype = C.c_int64
        _lib.xm_emu_deflate_literals.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_uint32), C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                                 C.POINTER(C.c_uint8), C.POINTER(C.c_uint8)]
        _lib.xm_emu_deflate_literals.restype = C.c_uint64
    return _lib


def index(buf, record_indices=(), skip_repeated=False, debug=0):
    """index pass of the sharded walk: (ShardInfo, byte offsets of the queried records)"""
    L = lib()
    buf = bytes(buf)
    n = len(record_indices)
    q = (C.c_uint64 * max(1, n))(*record_indices)
    off = (C.c_uint64 * max(1, n))()
    info = ShardInfo()
    err = C.create_string_buffer(512)
    rc = L.xm_emu_index(buf, len(buf), int(bool(skip_repeated)), debug, n, q, off, C.byref(info), err, 512)
    if rc:
        raise RuntimeError("xm_emu_index failed (%d): %s" % (rc, err.value.decode()))
    return info, list(off)[:n]


def classify(prim, sec, mode=0, score_src=0, skip_repeated=False, min_score=float("-inf"), enabled_bins=0x3F,
             debug=0, cap=None, chunk=None, first_is_context=False):
    """chunk: run the chunked walk (xm_stream.h) with that many new bytes per stream and step"""
    L = lib()
    prim, sec = bytes(prim), bytes(sec)
    if cap is None:
        cap = 4 * (len(prim) + len(sec)) + 64     # a pair walk can emit a line twice (overlapping units)
    bufs = [C.create_string_buffer(cap) for _ in range(6)]
    outp = (C.c_void_p * 6)(*[C.cast(b, C.c_void_p) for b in bufs])
    caps = (C.c_uint64 * 6)(*([cap] * 6))
    o = Opts(mode, score_src, int(bool(skip_repeated)) | (2 if first_is_context else 0), enabled_bins, min_score)
    r = Result()
    err = C.create_string_buffer(512)
    if chunk:
        rc = L.xm_emu_classify_stream(prim, len(prim), sec, len(sec), C.byref(o), debug, chunk, outp, caps, C.byref(r), err, 512)
    else:
        rc = L.xm_emu_classify(prim, len(prim), sec, len(sec), C.byref(o), debug, outp, caps, C.byref(r), err, 512)
    outs = [bufs[b].raw[:r.out_len[b]] for b in range(6)]
    return dict(status=rc, outputs=outs, counts=list(r.counts), n_records=r.n_records, err_record=r.err_record,
                bytes_in=list(r.bytes_in), message=err.value.decode())


def classify_sharded(prim, sec, rank, world, all_gather, exchange, mode=0, score_src=0, skip_repeated=False,
                     min_score=float("-inf"), enabled_bins=0x3F, room=1 << 16):
    """one rank of the walk across ranks (csrc/xm_shard.h) on the CPU.  prim / sec: this rank's BYTE shards;
    all_gather(send_addr, recv_addr, nbytes) and exchange(sends, ns, recvs, nr) are the test's collectives.
    status -2: every rank declined together (input needs the exact kernels)."""
    L = lib()
    prim, sec = bytes(prim), bytes(sec)
    cap = 4 * (len(prim) + len(sec)) + 2 * room + 4096
    bufs = [C.create_string_buffer(cap) for _ in range(6)]
    outp = (C.c_void_p * 6)(*[C.cast(b, C.c_void_p) for b in bufs])
    caps = (C.c_uint64 * 6)(*([cap] * 6))
    o = Opts(mode, score_src, int(bool(skip_repeated)), enabled_bins, min_score)
    r, st = Result(), ShardStats()
    err = C.create_string_buffer(512)
    ag, ex = ALL_GATHER_FN(all_gather), EXCHANGE_FN(exchange)
    rc = L.xm_emu_classify_sharded(prim, len(prim), sec, len(sec), C.byref(o), rank, world, ag, ex, room, outp, caps,
                                   C.byref(r), C.byref(st), err, 512)
    outs = [bufs[b].raw[:r.out_len[b]] for b in range(6)]
    return dict(status=rc, outputs=outs, counts=list(r.counts), n_records=int(r.n_records), err_record=int(r.err_record),
                out_offset=list(st.out_offset), out_total=list(st.out_total), records=(int(st.rec_lo), int(st.rec_hi)),
                sliver_bytes=int(st.sliver_bytes), message=err.value.decode())


def inflate(deflate_stream, out_len, misalign=0):
    """csrc/xm_inflate.h inflate_raw on the CPU: (status, bytes).  The stream is placed `misalign` bytes behind a 16-byte
    boundary, with the 12 readable bytes behind it the decoder's word fetches may touch."""
    L = lib()
    raw = bytes(deflate_stream)
    buf = C.create_string_buffer(len(raw) + 64)
    base = C.addressof(buf)
    at = base + ((16 - base % 16) % 16) + misalign
    C.memmove(at, raw, len(raw))
    out = C.create_string_buffer(max(1, out_len))
    rc = L.xm_emu_inflate(at, len(raw), out, out_len)
    return rc, out.raw[:out_len]


def crc32(data):
    data = bytes(data)
    buf = C.create_string_buffer(data, max(1, len(data)))
    return int(lib().xm_emu_crc32(buf, len(data)))


def bam_chain(inflated, first, n_ref, seg_bytes=16384, stop=0, want_guess=False):
    """record offsets of an inflated BAM stream by the segment-parallel chain: (offsets, end, repaired segments); None: corrupt.
    first=None: no record start is known (a part of a file): the chain starts at the first guess.  stop: only records that
    start before that byte."""
    data = bytes(inflated)
    buf = C.create_string_buffer(data, len(data) + 64)
    cap = len(data) // 36 + 16
    rec = (C.c_uint64 * cap)()
    end, rep, guess = C.c_uint64(), C.c_uint32(), C.c_uint64()
    n = lib().xm_emu_bam_chain(buf, len(data), seg_bytes, (1 << 64) - 1 if first is None else first, n_ref, rec, cap, C.byref(end), C.byref(rep),
                               stop, C.byref(guess))
    if n < 0:
        return None
    out = (list(rec[:n]), int(end.value), int(rep.value))
    return out + (int(guess.value),) if want_guess else out


def deflate_literals(sample, data, hist=None):
    """csrc/xm_deflate.h's plan (from a byte sample, or from 316 token counts) used to code `data` as one block of literals:
    (raw DEFLATE bytes, literal/length code lengths, distance code lengths)"""
    sample, data = bytes(sample), bytes(data)
    cap = 2 * len(data) + 4096
    out = C.create_string_buffer(cap)
    ll, dl = (C.c_uint8 * 286)(), (C.c_uint8 * 30)()
    h = (C.c_uint32 * 316)(*hist) if hist is not None else None
    n = lib().xm_emu_deflate_literals(sample, len(sample), h, data, len(data), out, cap, ll, dl)
    return out.raw[:n], list(ll), list(dl)
