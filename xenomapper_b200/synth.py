"""Deterministic synthetic SAM pair generator (ctypes over csrc/xm_synth.c).

Shapes and distributions: SURVEY.md section 8(d).  Every record is a pure
function of (seed, style, record index), so shards generate independently.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
STYLE_SE_BOWTIE2, STYLE_PE_BOWTIE2, STYLE_PE_HISAT = 0, 1, 2
_lib = None

HEADER_PRIMARY = ("@HD\tVN:1.0\tSO:unsorted\n@SQ\tSN:chr1\tLN:248956422\n@SQ\tSN:chr2\tLN:242193529\n"
                  "@PG\tID:bowtie2\tPN:bowtie2\tVN:2.2.6\tCL:\"bowtie2-align-s --local -x hg38\"\n")
HEADER_SECONDARY = ("@HD\tVN:1.0\tSO:unsorted\n@SQ\tSN:1\tLN:195471971\n@SQ\tSN:2\tLN:182113224\n"
                    "@PG\tID:bowtie2\tPN:bowtie2\tVN:2.2.6\tCL:\"bowtie2-align-s --local -x mm10\"\n")


def build():
    src = os.path.join(_HERE, "csrc", "xm_synth.c")
    out = os.path.join(_HERE, "_xm_synth.so")
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", out, src])
    return out


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "_xm_synth.so")
        src = os.path.join(_HERE, "csrc", "xm_synth.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        _lib = C.CDLL(path)
        _lib.xm_synth_generate.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_uint64,
                                           C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64),
                                           C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        _lib.xm_synth_generate.restype = C.c_int
        _lib.xm_synth_max_record.restype = C.c_uint64
    return _lib


def generate(n_records, seed=1, style=STYLE_SE_BOWTIE2, first=0, out_p=None, out_s=None):
    """Return (primary, secondary) uint8 arrays holding records [first, first+n)."""
    L = _load()
    cap = int(n_records) * int(L.xm_synth_max_record())
    if out_p is None:
        out_p = np.empty(cap, dtype=np.uint8)
    if out_s is None:
        out_s = np.empty(cap, dtype=np.uint8)
    lp, ls = C.c_uint64(), C.c_uint64()
    rc = L.xm_synth_generate(seed, style, first, n_records,
                             out_p.ctypes.data, out_p.nbytes, C.byref(lp),
                             out_s.ctypes.data, out_s.nbytes, C.byref(ls))
    if rc:
        raise ValueError("synthetic generator: buffers too small")
    return out_p[:lp.value], out_s[:ls.value]
