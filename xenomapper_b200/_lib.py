"""ctypes binding of libxenomapper_b200.so (include/xenomapper_b200.h).

The library is the only implementation of the read-binning walk in this
package: if it is missing or no B200 is usable, loading / Context() raises.
There is no Python or CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("XM_LIB_PATH") or os.path.join(_HERE, "libxenomapper_b200.so")

BINS = ("primary_specific", "secondary_specific", "primary_multi",
        "secondary_multi", "unassigned", "unresolved")
MODE_SE, MODE_PE_LIBERAL, MODE_PE_CONSERVATIVE = 0, 1, 2
SCORE_AS_XS, SCORE_AS_ZS, SCORE_CIGAR_NM = 0, 1, 2
(XM_OK, XM_ERR_ASSERT, XM_ERR_VALUE, XM_ERR_RUNTIME, XM_ERR_UNICODE, XM_ERR_UNSUPPORTED,
 XM_ERR_NOMEM, XM_ERR_CUDA, XM_ERR_ARG, XM_ERR_IO, XM_ERR_INDEX) = range(11)
DEBUG_FORCE_GENERIC, DEBUG_SMALL_TILES, DEBUG_ROWS, DEBUG_EXACT_NAMES = 1, 2, 4, 8

EXPORTS = ("xm_abi_version", "xm_create", "xm_destroy", "xm_last_error", "xm_classify_device",
           "xm_classify_host", "xm_get_output", "xm_classify_fds", "xm_count_device", "xm_dev_alloc",
           "xm_dev_free", "xm_host_alloc_pinned", "xm_host_free_pinned", "xm_memcpy_h2d", "xm_memcpy_d2h",
           "xm_memcpy_d2d", "xm_dev_mem_info", "xm_set_debug", "xm_locate_device", "xm_classify_bam_host",
           "xm_bam_header_text", "xm_bam_render_host", "xm_bam_get_stats", "xm_get_walk_kernels",
           "xm_comm_unique_id", "xm_comm_init_rank", "xm_comm_destroy", "xm_comm_barrier", "xm_comm_allreduce_f64",
           "xm_classify_sharded_device", "xm_classify_sharded_host", "xm_copy_ceiling",
           "xm_process_headers_fds", "xm_process_headers_mem", "xm_classify_fds_ex", "xm_bgzf_write", "xm_classify_streams",
           "xm_bgzf_deflate_host", "xm_bgzf_get_stats", "xm_classify_bam_fds",
           "xm_bam_shard_open", "xm_bam_shard_chain", "xm_bam_shard_text")
OUT_BGZF = 1


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
    _fields_ = [("n_records", C.c_uint64), ("first_start", C.c_uint64), ("stop_at", C.c_uint64),
                ("end_off", C.c_uint64)]


class ShardStats(C.Structure):
    _fields_ = [("rec_lo", C.c_uint64), ("rec_hi", C.c_uint64), ("n_records_total", C.c_uint64),
                ("out_offset", C.c_uint64 * 6), ("out_total", C.c_uint64 * 6), ("sliver_bytes", C.c_uint64),
                ("sent_bytes", C.c_uint64), ("align_ms", C.c_float), ("index_ms", C.c_float), ("sliver_ms", C.c_float),
                ("walk_ms", C.c_float), ("comm_ms", C.c_float), ("total_ms", C.c_float), ("n_collectives", C.c_uint32),
                ("first_bad_rank", C.c_int32)]


class Headers(C.Structure):
    _fields_ = [("record_offset", C.c_uint64 * 2), ("text", C.c_void_p * 6), ("text_len", C.c_uint64 * 6),
                ("status", C.c_int32 * 6), ("failed_input", C.c_int32)]


class BamStats(C.Structure):
    _fields_ = [("inflate_s", C.c_double), ("render_ms", C.c_float), ("n_launches", C.c_uint32), ("bam_bytes", C.c_uint64),
                ("inflated_bytes", C.c_uint64), ("text_bytes", C.c_uint64), ("records", C.c_uint64),
                ("upload_s", C.c_double), ("inflate_ms", C.c_float), ("chain_repairs", C.c_uint32)]


class BgzfStats(C.Structure):
    _fields_ = [("in_bytes", C.c_uint64), ("out_bytes", C.c_uint64), ("members", C.c_uint64), ("kernel_ms", C.c_float), ("n_launches", C.c_uint32)]


class XenomapperLibraryError(RuntimeError):
    """CUDA / allocation / argument failures of the native library."""


class UnsupportedInput(ValueError):
    """Input the reference accepts but the device grammar does not (XM_ERR_UNSUPPORTED)."""


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XenomapperLibraryError(
            "%s is missing: build it with `python -m xenomapper_b200.build` (needs nvcc). "
            "xenomapper_b200 has no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u64, i = C.c_void_p, C.c_uint64, C.c_int
    L.xm_abi_version.restype = i
    L.xm_create.argtypes = [i, C.c_uint32]
    L.xm_create.restype = vp
    L.xm_destroy.argtypes = [vp]
    L.xm_destroy.restype = None
    L.xm_last_error.argtypes = [vp]
    L.xm_last_error.restype = C.c_char_p
    L.xm_classify_device.argtypes = [vp, vp, u64, vp, u64, C.POINTER(Opts), C.POINTER(vp), C.POINTER(u64), C.POINTER(Result)]
    L.xm_classify_host.argtypes = [vp, vp, u64, vp, u64, C.POINTER(Opts), C.POINTER(Result)]
    L.xm_get_output.argtypes = [vp, i, C.POINTER(vp), C.POINTER(u64)]
    L.xm_classify_fds.argtypes = [vp, i, C.c_int64, i, C.c_int64, C.POINTER(i), C.POINTER(Opts), C.POINTER(Result)]
    L.xm_count_device.argtypes = [vp, vp, u64, i, C.POINTER(ShardInfo)]
    L.xm_locate_device.argtypes = [vp, vp, u64, i, C.c_uint32, C.POINTER(u64), C.POINTER(u64)]
    L.xm_dev_alloc.argtypes = [vp, u64, C.POINTER(vp)]
    L.xm_dev_free.argtypes = [vp, vp]
    L.xm_host_alloc_pinned.argtypes = [vp, u64, C.POINTER(vp)]
    L.xm_host_free_pinned.argtypes = [vp, vp]
    L.xm_memcpy_h2d.argtypes = [vp, vp, vp, u64]
    L.xm_memcpy_d2h.argtypes = [vp, vp, vp, u64]
    L.xm_memcpy_d2d.argtypes = [vp, vp, vp, u64]
    L.xm_dev_mem_info.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.xm_set_debug.argtypes = [vp, C.c_uint32]
    L.xm_classify_bam_host.argtypes = [vp, vp, u64, vp, u64, C.POINTER(Opts), C.POINTER(Result)]
    L.xm_bam_header_text.argtypes = [vp, u64, vp, u64, C.POINTER(u64)]
    L.xm_bam_render_host.argtypes = [vp, vp, u64, C.POINTER(vp), C.POINTER(u64)]
    L.xm_bam_get_stats.argtypes = [vp, C.POINTER(BamStats), i]
    L.xm_get_walk_kernels.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.xm_copy_ceiling.argtypes = [vp, u64, u64, i, C.POINTER(C.c_float)]
    L.xm_classify_fds_ex.argtypes = [vp, i, C.c_int64, i, C.c_int64, C.POINTER(i), C.POINTER(Opts), C.c_uint32, C.POINTER(Result)]
    L.xm_bgzf_write.argtypes = [i, vp, u64, i]
    L.xm_classify_streams.argtypes = [vp, i, i, C.POINTER(i), C.POINTER(Opts), C.c_uint32, C.c_char_p, C.POINTER(Result)]
    L.xm_process_headers_fds.argtypes = [i, i, C.c_char_p, C.POINTER(Headers)]
    L.xm_process_headers_mem.argtypes = [vp, u64, vp, u64, C.c_char_p, C.POINTER(Headers)]
    L.xm_comm_unique_id.argtypes = [vp]
    L.xm_comm_init_rank.argtypes = [vp, i, i, vp]
    L.xm_comm_destroy.argtypes = [vp]
    L.xm_comm_barrier.argtypes = [vp]
    L.xm_comm_allreduce_f64.argtypes = [vp, C.POINTER(C.c_double), i, i]
    L.xm_classify_sharded_device.argtypes = [vp, vp, u64, vp, u64, u64, u64, C.POINTER(Opts), C.POINTER(vp), C.POINTER(u64),
                                             C.POINTER(Result), C.POINTER(ShardStats)]
    L.xm_classify_sharded_host.argtypes = [vp, vp, u64, vp, u64, C.POINTER(Opts), C.POINTER(Result), C.POINTER(ShardStats)]
    for name in EXPORTS:
        if name not in ("xm_create", "xm_destroy", "xm_last_error", "xm_abi_version"):
            getattr(L, name).restype = i
    _lib = L
    return L


def _host_ptr(buf):
    """(address, length, keep-alive) for bytes / bytearray / memoryview / numpy uint8 arrays"""
    if isinstance(buf, bytes):
        return C.cast(C.c_char_p(buf), C.c_void_p), len(buf), buf
    if isinstance(buf, (bytearray, memoryview)):
        mv = memoryview(buf).cast("B")
        if mv.readonly:
            b = mv.tobytes()
            return C.cast(C.c_char_p(b), C.c_void_p), len(b), b
        arr = (C.c_char * len(mv)).from_buffer(mv)
        return C.cast(arr, C.c_void_p), len(mv), arr
    return C.c_void_p(buf.ctypes.data), buf.nbytes, buf


def bgzf_write(fd, data=b"", eof=False):
    """data as BGZF members appended to fd; eof: the empty end-of-file member behind them"""
    L = load()
    a, n, keep = _host_ptr(data) if data else (None, 0, None)
    rc = L.xm_bgzf_write(fd, a, n, int(bool(eof)))
    if rc != XM_OK:
        raise XenomapperLibraryError("xm_bgzf_write failed (%d): %s" % (rc, L.xm_last_error(None).decode()))


def process_headers(prim, sec, version):
    """(rc, failed input, record offsets, six header texts, six statuses) of two SAM files given as descriptors
    (ints) or bytes-like objects: xm_process_headers_fds / _mem (xm.py:36-46, 120-174 on raw bytes; no GPU needed)"""
    L = load()
    h = Headers()
    if isinstance(prim, int):
        rc = L.xm_process_headers_fds(prim, sec, version.encode(), C.byref(h))
    else:
        pa, pn, pk = _host_ptr(prim)
        sa, sn, sk = _host_ptr(sec)
        rc = L.xm_process_headers_mem(pa, pn, sa, sn, version.encode(), C.byref(h))
        del pk, sk
    texts = [C.string_at(h.text[b], h.text_len[b]) if rc == XM_OK and h.text_len[b] else b"" for b in range(6)]
    return rc, h.failed_input, list(h.record_offset), texts, list(h.status)


class Context:
    """One xm_ctx: a CUDA device, its streams and scratch."""

    def __init__(self, device=0):
        self.lib = load()
        self.h = self.lib.xm_create(device, 0)
        if not self.h:
            raise XenomapperLibraryError(self.lib.xm_last_error(None).decode() or "xm_create failed")

    def close(self):
        if getattr(self, "h", None):
            self.lib.xm_destroy(self.h)
            self.h = None

    __del__ = close

    def error(self):
        return self.lib.xm_last_error(self.h).decode()

    def _check(self, rc, what):
        if rc in (XM_ERR_NOMEM, XM_ERR_CUDA, XM_ERR_ARG, XM_ERR_IO):
            raise XenomapperLibraryError("%s failed (%d): %s" % (what, rc, self.error()))
        return rc

    def set_debug(self, flags):
        self.lib.xm_set_debug(self.h, flags)

    # ---- memory -------------------------------------------------------------
    def dev_alloc(self, n):
        p = C.c_void_p()
        self._check(self.lib.xm_dev_alloc(self.h, n, C.byref(p)), "xm_dev_alloc")
        return p.value

    def dev_free(self, p):
        self.lib.xm_dev_free(self.h, p)

    def pinned_alloc(self, n):
        p = C.c_void_p()
        self._check(self.lib.xm_host_alloc_pinned(self.h, n, C.byref(p)), "xm_host_alloc_pinned")
        return p.value

    def pinned_free(self, p):
        self.lib.xm_host_free_pinned(self.h, p)

    def h2d(self, d, host, n=None):
        a, ln, keep = _host_ptr(host) if not isinstance(host, int) else (C.c_void_p(host), n, None)
        self._check(self.lib.xm_memcpy_h2d(self.h, d, a, ln if n is None else n), "xm_memcpy_h2d")

    def d2h(self, d, n):
        buf = C.create_string_buffer(n) if n else C.create_string_buffer(1)
        if n:
            self._check(self.lib.xm_memcpy_d2h(self.h, buf, d, n), "xm_memcpy_d2h")
        return buf.raw[:n]

    def d2d(self, dst, src, n):
        self._check(self.lib.xm_memcpy_d2d(self.h, dst, src, n), "xm_memcpy_d2d")

    def mem_info(self):
        f, t = C.c_uint64(), C.c_uint64()
        self.lib.xm_dev_mem_info(self.h, C.byref(f), C.byref(t))
        return f.value, t.value

    # ---- walks --------------------------------------------------------------
    @staticmethod
    def opts(mode=MODE_SE, score_src=SCORE_AS_XS, skip_repeated=False, min_score=float("-inf"), enabled_bins=0x3F,
             first_is_context=False):
        return Opts(mode, score_src, int(bool(skip_repeated)) | (2 if first_is_context else 0), enabled_bins, min_score)

    def classify_device(self, d_prim, prim_len, d_sec, sec_len, opts, d_out, out_cap):
        res = Result()
        outp = (C.c_void_p * 6)(*d_out)
        caps = (C.c_uint64 * 6)(*out_cap)
        rc = self.lib.xm_classify_device(self.h, d_prim, prim_len, d_sec, sec_len, C.byref(opts), outp, caps, C.byref(res))
        self._check(rc, "xm_classify_device")
        return rc, res

    def classify_host(self, prim, sec, opts, want_outputs=True):
        """prim / sec: bytes-like record regions (or (address, length) tuples of pinned memory)."""
        pa, pn, pk = prim + (None,) if isinstance(prim, tuple) else _host_ptr(prim)
        sa, sn, sk = sec + (None,) if isinstance(sec, tuple) else _host_ptr(sec)
        res = Result()
        rc = self.lib.xm_classify_host(self.h, pa, pn, sa, sn, C.byref(opts), C.byref(res))
        self._check(rc, "xm_classify_host")
        outs = None
        if want_outputs:
            outs = []
            for b in range(6):
                p, n = C.c_void_p(), C.c_uint64()
                self.lib.xm_get_output(self.h, b, C.byref(p), C.byref(n))
                outs.append(C.string_at(p.value, n.value) if n.value else b"")
        del pk, sk
        return rc, res, outs

    def copy_ceiling(self, h2d_bytes, d2h_bytes, reps=3):
        """ms one round of bare pinned copies takes: h2d_bytes up and d2h_bytes down at the same time, no kernels"""
        ms = C.c_float()
        self._check(self.lib.xm_copy_ceiling(self.h, h2d_bytes, d2h_bytes, reps, C.byref(ms)), "xm_copy_ceiling")
        return ms.value

    # ---- the walk across GPUs: one process per GPU, NCCL inside the library ------
    @staticmethod
    def comm_unique_id():
        """128 opaque bytes made by rank 0; the launcher hands them to every rank (comm_init_rank)"""
        buf = C.create_string_buffer(128)
        L = load()
        if L.xm_comm_unique_id(buf) != XM_OK:
            raise XenomapperLibraryError("xm_comm_unique_id failed: " + L.xm_last_error(None).decode())
        return buf.raw

    def comm_init_rank(self, nranks, rank, unique_id=None):
        buf = C.create_string_buffer(bytes(unique_id), 128) if unique_id is not None else None
        self._check(self.lib.xm_comm_init_rank(self.h, nranks, rank, buf), "xm_comm_init_rank")

    def comm_barrier(self):
        self._check(self.lib.xm_comm_barrier(self.h), "xm_comm_barrier")

    def comm_allreduce(self, values, op="sum"):
        arr = (C.c_double * len(values))(*values)
        self._check(self.lib.xm_comm_allreduce_f64(self.h, arr, len(values), 1 if op == "max" else 0), "xm_comm_allreduce_f64")
        return list(arr)

    def classify_sharded_host(self, prim, sec, opts, want_outputs=True):
        """prim / sec: this rank's BYTE shards of the two record regions (bytes-like).  Returns (rc, Result with the
        whole job's counts, ShardStats with this rank's place in the six bins, this rank's six outputs)."""
        pa, pn, pk = _host_ptr(prim)
        sa, sn, sk = _host_ptr(sec)
        res, st = Result(), ShardStats()
        rc = self.lib.xm_classify_sharded_host(self.h, pa, pn, sa, sn, C.byref(opts), C.byref(res), C.byref(st))
        self._check(rc, "xm_classify_sharded_host")
        outs = None
        if want_outputs:
            outs = []
            for b in range(6):
                p, n = C.c_void_p(), C.c_uint64()
                self.lib.xm_get_output(self.h, b, C.byref(p), C.byref(n))
                outs.append(C.string_at(p.value, n.value) if n.value else b"")
        del pk, sk
        return rc, res, st, outs

    def classify_sharded_device(self, d_prim, prim_len, d_sec, sec_len, front_room, back_room, opts, d_out, out_cap):
        res, st = Result(), ShardStats()
        outp = (C.c_void_p * 6)(*d_out)
        caps = (C.c_uint64 * 6)(*out_cap)
        rc = self.lib.xm_classify_sharded_device(self.h, d_prim, prim_len, d_sec, sec_len, front_room, back_room, C.byref(opts),
                                                 outp, caps, C.byref(res), C.byref(st))
        self._check(rc, "xm_classify_sharded_device")
        return rc, res, st

    def classify_bam_host(self, prim_bam, sec_bam, opts, want_outputs=True):
        """prim_bam / sec_bam: bytes-like BAM files (BGZF): inflated on the host, rendered and walked on the device"""
        pa, pn, pk = _host_ptr(prim_bam)
        sa, sn, sk = _host_ptr(sec_bam)
        res = Result()
        rc = self.lib.xm_classify_bam_host(self.h, pa, pn, sa, sn, C.byref(opts), C.byref(res))
        self._check(rc, "xm_classify_bam_host")
        outs = None
        if want_outputs:
            outs = []
            for b in range(6):
                p, n = C.c_void_p(), C.c_uint64()
                self.lib.xm_get_output(self.h, b, C.byref(p), C.byref(n))
                outs.append(C.string_at(p.value, n.value) if n.value else b"")
        del pk, sk
        return rc, res, outs

    def classify_bam_fds(self, prim_bam, sec_bam, out_fds, opts, out_flags=0):
        """BAM files (bytes-like, or numpy arrays over mapped files) in, the six bins appended to descriptors"""
        pa, pn, pk = _host_ptr(prim_bam)
        sa, sn, sk = _host_ptr(sec_bam)
        res = Result()
        fds = (C.c_int * 6)(*out_fds)
        rc = self.lib.xm_classify_bam_fds(self.h, pa, pn, sa, sn, fds, C.byref(opts), out_flags, C.byref(res))
        self._check(rc, "xm_classify_bam_fds")
        del pk, sk
        return rc, res

    NONE64 = (1 << 64) - 1

    def bam_shard_open(self, stream, bam, rank, world):
        """this rank's part of a BAM file: (guess, exit) offsets in the inflated stream (NONE64: no record start of its own);
        the buffer must stay alive until bam_shard_text has been called"""
        a, n, keep = _host_ptr(bam)
        g, e = C.c_uint64(), C.c_uint64()
        rc = self.lib.xm_bam_shard_open(self.h, stream, a, C.c_uint64(n), rank, world, C.byref(g), C.byref(e))
        self._check(rc, "xm_bam_shard_open")
        if rc:
            raise XenomapperLibraryError("xm_bam_shard_open failed (%d): %s" % (rc, self.error()))
        self._shard_keep = getattr(self, "_shard_keep", {})
        self._shard_keep[stream] = keep
        return int(g.value), int(e.value)

    def bam_shard_chain(self, stream, entry):
        e = C.c_uint64()
        rc = self.lib.xm_bam_shard_chain(self.h, stream, C.c_uint64(entry), C.byref(e))
        self._check(rc, "xm_bam_shard_chain")
        if rc:
            raise XenomapperLibraryError("xm_bam_shard_chain failed (%d): %s" % (rc, self.error()))
        return int(e.value)

    def bam_shard_text(self, stream, front_room, back_room):
        """(device pointer of the part's SAM text, its length)"""
        p, n = C.c_void_p(), C.c_uint64()
        rc = self.lib.xm_bam_shard_text(self.h, stream, C.c_uint64(front_room), C.c_uint64(back_room), C.byref(p), C.byref(n))
        self._check(rc, "xm_bam_shard_text")
        if rc:
            raise XenomapperLibraryError("xm_bam_shard_text failed (%d): %s" % (rc, self.error()))
        return p.value, int(n.value)

    def bam_render_host(self, bam):
        """all records of a BAM file as SAM text (what `samtools view` prints), rendered on the device"""
        a, n, keep = _host_ptr(bam)
        p, ln = C.c_void_p(), C.c_uint64()
        rc = self.lib.xm_bam_render_host(self.h, a, n, C.byref(p), C.byref(ln))
        self._check(rc, "xm_bam_render_host")
        if rc == XM_ERR_UNSUPPORTED:
            raise UnsupportedInput(self.error())
        del keep
        return C.string_at(p.value, ln.value) if ln.value else b""

    def bgzf_deflate_host(self, data):
        """data as BGZF members (no end-of-file member), deflated on the device"""
        a, n, keep = _host_ptr(data)
        p, ln = C.c_void_p(), C.c_uint64()
        rc = self.lib.xm_bgzf_deflate_host(self.h, a, n, C.byref(p), C.byref(ln))
        self._check(rc, "xm_bgzf_deflate_host")
        del keep
        return C.string_at(p.value, ln.value) if ln.value else b""

    def bgzf_stats(self, reset=True):
        st = BgzfStats()
        self.lib.xm_bgzf_get_stats(self.h, C.byref(st), int(reset))
        return st

    def walk_kernels(self):
        """names of the kernels the last resident walk ran"""
        m = C.c_uint32()
        self.lib.xm_get_walk_kernels(self.h, C.byref(m))
        return [n for b, n in enumerate(("k_scan2", "k_classify2", "k_scan", "k_classify", "k_size+k_prefix+k_emit")) if (m.value >> b) & 1]

    def bam_stats(self, reset=True):
        st = BamStats()
        self.lib.xm_bam_get_stats(self.h, C.byref(st), int(reset))
        return st

    def classify_fds(self, fd_prim, off_prim, fd_sec, off_sec, out_fds, opts, out_flags=0):
        """out_flags=OUT_BGZF: the bins leave as BGZF members (bgzf_write puts the header in front and the end-of-file
        member behind them)"""
        res = Result()
        fds = (C.c_int * 6)(*out_fds)
        rc = self.lib.xm_classify_fds_ex(self.h, fd_prim, off_prim, fd_sec, off_sec, fds, C.byref(opts), out_flags, C.byref(res))
        self._check(rc, "xm_classify_fds")
        return rc, res

    def classify_streams(self, fd_prim, fd_sec, out_fds, opts, version, out_flags=0):
        """headers + walk in one call for inputs that cannot seek (pipes): xm_classify_streams"""
        res = Result()
        fds = (C.c_int * 6)(*out_fds)
        rc = self.lib.xm_classify_streams(self.h, fd_prim, fd_sec, fds, C.byref(opts), out_flags, version.encode(), C.byref(res))
        if rc != XM_ERR_INDEX:
            self._check(rc, "xm_classify_streams")
        return rc, res

    def count_device(self, d_buf, n, skip_repeated=False):
        info = ShardInfo()
        self._check(self.lib.xm_count_device(self.h, d_buf, n, int(bool(skip_repeated)), C.byref(info)), "xm_count_device")
        return info


    def locate_device(self, d_buf, n, record_indices, skip_repeated=False):
        q = (C.c_uint64 * len(record_indices))(*record_indices)
        out = (C.c_uint64 * len(record_indices))()
        self._check(self.lib.xm_locate_device(self.h, d_buf, n, int(bool(skip_repeated)), len(record_indices), q, out), "xm_locate_device")
        return list(out)


_default_ctx = None


def default_context():
    """Process-wide context on the device named by XENOMAPPER_DEVICE (default: LOCAL_RANK or 0)."""
    global _default_ctx
    if _default_ctx is None:
        dev = int(os.environ.get("XENOMAPPER_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        _default_ctx = Context(dev)
    return _default_ctx


def bam_header_text(bam):
    """header text stored in a BAM file (no device needed: host inflate of the leading BGZF blocks)"""
    L = load()
    a, n, keep = _host_ptr(bam)
    need = C.c_uint64()
    rc = L.xm_bam_header_text(a, n, None, 0, C.byref(need))
    if rc:
        raise XenomapperLibraryError("xm_bam_header_text failed (%d): %s" % (rc, L.xm_last_error(None).decode()))
    buf = C.create_string_buffer(max(1, need.value))
    L.xm_bam_header_text(a, n, buf, need.value, C.byref(need))
    del keep
    return buf.raw[:need.value]
