"""ctypes face of the CPU oracle (oracle/xm_oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported from tests/, __graft_entry__.smoke() and bench.py's CPU legs;
never from xenomapper_b200/.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
STATES = ("primary_specific", "secondary_specific", "primary_multi",
          "secondary_multi", "unassigned", "unresolved")
MODE_SE, MODE_PE_LIBERAL, MODE_PE_CONSERVATIVE = 0, 1, 2
SCORE_AS_XS, SCORE_AS_ZS, SCORE_CIGAR_NM = 0, 1, 2
ERR_NAMES = {0: None, 1: "AssertionError", 2: "ValueError", 3: "RuntimeError",
             4: "UnicodeDecodeError", 5: "Unsupported", 6: "MemoryError"}


class _Opts(C.Structure):
    _fields_ = [("mode", C.c_int32), ("score_src", C.c_int32), ("skip_repeated", C.c_int32),
                ("enabled_bins", C.c_uint32), ("min_score", C.c_double)]


class _Result(C.Structure):
    _fields_ = [("data", C.c_void_p * 6), ("len", C.c_size_t * 6), ("counts", C.c_uint64 * 36),
                ("n_yielded", C.c_uint64), ("err", C.c_int32), ("err_record", C.c_uint64),
                ("errmsg", C.c_char * 128)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "xm_oracle.c")):
            build()
        _lib = C.CDLL(path)
        _lib.xmo_classify.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                      C.POINTER(_Opts), C.POINTER(_Result)]
        _lib.xmo_classify.restype = C.c_int
        _lib.xmo_free.argtypes = [C.POINTER(_Result)]
        _lib.xmo_mapping_state.argtypes = [C.c_double] * 5
        _lib.xmo_mapping_state.restype = C.c_int
        _lib.xmo_pair_bin.argtypes = [C.c_int] * 3
        _lib.xmo_pair_bin.restype = C.c_int
        _lib.xmo_line_scores.argtypes = [C.c_char_p, C.c_size_t, C.c_int,
                                         C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib.xmo_line_scores.restype = C.c_int
    return _lib


def _addr(buf):
    """address + keep-alive object for bytes / bytearray / numpy uint8 arrays"""
    if isinstance(buf, bytes):
        return C.cast(C.c_char_p(buf), C.c_void_p), len(buf), buf
    if isinstance(buf, (bytearray, memoryview)):
        arr = (C.c_char * len(buf)).from_buffer(buf)
        return C.cast(arr, C.c_void_p), len(buf), arr
    return C.c_void_p(buf.ctypes.data), buf.nbytes, buf      # numpy


def classify(prim, sec, mode=MODE_SE, score_src=SCORE_AS_XS, skip_repeated=False,
             min_score=float("-inf"), enabled_bins=0x3F, want_outputs=True):
    """Run the walk on two record regions (headers removed).

    Returns dict(outputs=[6 x bytes], counts=[36], n_yielded, err, err_record).
    """
    L = lib()
    o = _Opts(mode, score_src, int(bool(skip_repeated)), enabled_bins, min_score)
    r = _Result()
    pa, pn, pk = _addr(prim)
    sa, sn, sk = _addr(sec)
    L.xmo_classify(pa, pn, sa, sn, C.byref(o), C.byref(r))
    outs = None
    if want_outputs:
        outs = [C.string_at(r.data[b], r.len[b]) if r.len[b] else b"" for b in range(6)]
    res = dict(outputs=outs, out_len=[r.len[b] for b in range(6)], counts=list(r.counts),
               n_yielded=r.n_yielded, err=r.err, err_record=r.err_record)
    L.xmo_free(C.byref(r))
    del pk, sk
    return res


def mapping_state(AS1, XS1, AS2, XS2, min_score=float("-inf")):
    s = lib().xmo_mapping_state(AS1, XS1, AS2, XS2, min_score)
    return STATES[s] if s >= 0 else None


def line_scores(line, score_src=SCORE_AS_XS):
    a, x = C.c_double(), C.c_double()
    rc = lib().xmo_line_scores(line, len(line), score_src, C.byref(a), C.byref(x))
    return rc, a.value, x.value
