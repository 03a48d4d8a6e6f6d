"""BGZF writer of the library (csrc/xm_bgzf.h, xm_bgzf_write): structure per the SAM/BAM specification section 4.1,
and gunzip(output) == input.  Host code only: runs without a GPU."""
import gzip
import os
import struct

import pytest

from xenomapper_b200 import _lib

EOF_MEMBER = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def members(raw):
    """split a BGZF file into its gzip members by the BC extra field's BSIZE"""
    out, at = [], 0
    while at < len(raw):
        assert raw[at:at + 4] == b"\x1f\x8b\x08\x04"
        xlen = struct.unpack_from("<H", raw, at + 10)[0]
        assert xlen == 6 and raw[at + 12:at + 16] == b"BC\x02\x00"
        bsize = struct.unpack_from("<H", raw, at + 16)[0] + 1
        out.append(raw[at:at + bsize])
        at += bsize
    assert at == len(raw)
    return out


@pytest.mark.parametrize("n", [0, 1, 65279, 65280, 65281, 1_000_003])
def test_bgzf_members_and_round_trip(tmp_path, n):
    data = bytes((i * 131 + (i >> 9)) & 0x7f for i in range(n))
    path = tmp_path / "x.gz"
    fd = os.open(path, os.O_WRONLY | os.O_CREAT, 0o600)
    _lib.bgzf_write(fd, data)
    _lib.bgzf_write(fd, eof=True)
    os.close(fd)
    raw = path.read_bytes()
    assert gzip.decompress(raw) == data
    ms = members(raw)
    assert ms[-1] == EOF_MEMBER
    assert len(ms) == (n + 0xff00 - 1) // 0xff00 + 1
    for m in ms[:-1]:
        isize = struct.unpack_from("<I", m, len(m) - 4)[0]
        assert 0 < isize <= 0xff00 and len(m) <= 0x10000
