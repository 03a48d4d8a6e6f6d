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


# ---- the device compressor (csrc/xm_deflate.h) ------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx():
    c = _lib.Context(0)
    yield c
    c.close()


def _device_cases():
    import random
    from xenomapper_b200 import synth
    rnd = random.Random(9)
    p, s = synth.generate(9000, seed=3, style=1)
    yield "one byte", b"x", None
    yield "65279", bytes(p)[:65279], None
    yield "65280", bytes(p)[:65280], None
    yield "65281", bytes(p)[:65281], None
    yield "sam text", bytes(p), 0.5
    yield "secondary sam text", bytes(s), 0.5
    yield "zeros", bytes(300000), 0.02
    yield "one run and a tail", b"A" * 70000 + b"xyz", 0.02
    yield "period 3", b"abc" * 50000, 0.05
    yield "random bytes (stored members)", bytes(rnd.getrandbits(8) for _ in range(200000)), 1.01
    yield "odd length", bytes(p)[:131073], None


DEVICE_CASES = list(_device_cases())


@pytest.mark.gpu
@pytest.mark.parametrize("name,data,max_ratio", DEVICE_CASES, ids=[c[0] for c in DEVICE_CASES])
def test_device_deflate_is_read_by_any_inflater(ctx, name, data, max_ratio):
    z = ctx.bgzf_deflate_host(data)
    assert gzip.decompress(z + EOF_MEMBER) == data
    ms = members(z)
    assert (len(data) + 0xff00 - 1) // 0xff00 <= len(ms) <= (len(data) + 0x4000 - 1) // 0x4000      # small inputs are cut finer
    at = 0
    for m in ms:
        isize = struct.unpack_from("<I", m, len(m) - 4)[0]
        assert 0 < isize <= 0xff00 and len(m) <= 0x10000
        import zlib
        assert struct.unpack_from("<I", m, len(m) - 8)[0] == zlib.crc32(data[at:at + isize])
        at += isize
    if max_ratio is not None:
        assert len(z) <= max_ratio * len(data) + 64, (len(z), len(data))


@pytest.mark.gpu
def test_device_deflate_of_nothing(ctx):
    assert ctx.bgzf_deflate_host(b"") == b""


@pytest.mark.gpu
def test_device_deflate_output_is_read_by_the_device_inflater(ctx):
    """a BAM whose BGZF members were made by the device compressor goes through k_bgzf_inflate"""
    from tests import _bamwriter
    from tests.test_bam import FULL_HEADER
    from xenomapper_b200 import synth
    p, _ = synth.generate(5000, seed=12, style=1)
    raw = gzip.decompress(_bamwriter.sam_to_bam(FULL_HEADER, bytes(p)))
    bam = ctx.bgzf_deflate_host(raw) + EOF_MEMBER
    assert ctx.bam_render_host(bam) == bytes(p)
