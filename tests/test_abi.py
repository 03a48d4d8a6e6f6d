"""The C-ABI library: loads, exports every symbol include/xenomapper_b200.h declares, and fails loudly without a GPU."""
import ctypes
import io
import os
import re

import pytest

from tests.conftest import HAS_GPU, ROOT


def declared_functions():
    text = open(os.path.join(ROOT, "include", "xenomapper_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_entry_points():
    names = declared_functions()
    for must in ("xm_create", "xm_destroy", "xm_classify_device", "xm_classify_host", "xm_classify_fds", "xm_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from xenomapper_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(L, name), "libxenomapper_b200.so does not export " + name
    assert _lib.load().xm_abi_version() == 6
    assert sorted(_lib.EXPORTS) == declared_functions()


def test_struct_layouts_match_the_header():
    from xenomapper_b200 import _lib
    assert ctypes.sizeof(_lib.Opts) == 24
    assert ctypes.sizeof(_lib.Result) == 36 * 8 + 8 + 48 + 16 + 8 + 8 + 16 + 24 + 8
    assert ctypes.sizeof(_lib.ShardInfo) == 32


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback():
    from xenomapper_b200 import _lib, xenomapper as xm
    L = _lib.load()
    assert not L.xm_create(0, 0)
    assert b"no usable CUDA device" in L.xm_last_error(None)
    with pytest.raises(_lib.XenomapperLibraryError):
        _lib.Context(0)
    rec = "r1\t0\tchr1\t1\t42\t5M\t*\t0\t0\tACGTA\tFFFFF\tAS:i:10\n"
    pairs = xm.getReadPairs(io.StringIO(rec), io.StringIO(rec))
    with pytest.raises(_lib.XenomapperLibraryError):
        xm.main_single_end(pairs, primary_specific=io.StringIO())


def test_product_package_does_not_touch_the_oracle():
    """only tests/, smoke() and bench.py's CPU legs may use oracle/ (or the CPU emulation harness)"""
    pkg = os.path.join(ROOT, "xenomapper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".c", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|tests)\b", src, flags=re.M), f
                assert "liboracle" not in src and "xm_emu" not in src.replace("xm_emu.cpp", ""), f
