"""The device inflate (csrc/xm_inflate.h) and the parallel BAM record chain (csrc/xm_bamchain.h), run on the CPU from the
same source the GPU compiles, against zlib and the serial chain."""
import gzip
import os
import random
import struct
import zlib

import pytest

from tests import _bamwriter, _emu


def deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return c.compress(data) + c.flush()


def samples():
    rnd = random.Random(11)
    text = ("read%07d\t99\tchr1\t%d\t42\t150M\t=\t%d\t300\t" % (7, 12345, 12645)).encode() + bytes(rnd.choice(b"ACGT") for _ in range(150))
    yield "empty", b""
    yield "one byte", b"x"
    yield "short text", text
    yield "zeros", bytes(65280)
    yield "run of one byte", b"A" * 5000
    yield "period 3", b"abc" * 9000
    yield "random bytes", bytes(rnd.getrandbits(8) for _ in range(40000))
    yield "dna", bytes(rnd.choice(b"ACGT") for _ in range(65280))
    yield "sam-like", b"".join(("r%d\t%d\tchr%d\t%d\t%d\t100M\t*\t0\t0\t" % (k, rnd.choice((0, 16, 99, 147)), rnd.randrange(1, 23), rnd.randrange(1, 10 ** 8), rnd.randrange(60))).encode()
                               + bytes(rnd.choice(b"ACGT") for _ in range(100)) + b"\t" + bytes(rnd.randrange(35, 74) for _ in range(100)) + b"\tAS:i:-%d\tXS:i:-%d\n" % (rnd.randrange(40), rnd.randrange(60))
                               for k in range(200))
    yield "skewed alphabet (long codes)", bytes(min(255, int(rnd.expovariate(0.08))) for _ in range(60000))


SAMPLES = list(samples())


@pytest.mark.parametrize("name,data", SAMPLES, ids=[s[0] for s in SAMPLES])
@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_inflate_equals_zlib(name, data, level):
    stream = deflate(data, level)
    for mis in (0, 1, 2, 3, 7):
        rc, out = _emu.inflate(stream, len(data), misalign=mis)
        assert rc == 0, (name, level, mis, rc)
        assert out == data


@pytest.mark.parametrize("strategy", [zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED])
def test_inflate_other_strategies(strategy):
    for name, data in SAMPLES:
        stream = deflate(data, 6, strategy)
        rc, out = _emu.inflate(stream, len(data))
        assert rc == 0 and out == data, (name, strategy, rc)


def test_inflate_stream_of_several_blocks_of_every_type():
    rnd = random.Random(5)
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    parts, stream = [], b""
    for k in range(12):
        piece = bytes(rnd.choice(b"ACGTN\t\n0123") for _ in range(rnd.randrange(1, 9000)))
        parts.append(piece)
        stream += c.compress(piece) + c.flush(zlib.Z_FULL_FLUSH if k % 3 else zlib.Z_SYNC_FLUSH)      # empty stored blocks in between
    stream += c.flush()
    data = b"".join(parts)
    rc, out = _emu.inflate(stream, len(data))
    assert rc == 0 and out == data


def test_inflate_refuses_wrong_size_and_garbage_without_reading_astray():
    data = SAMPLES[8][1]
    stream = deflate(data)
    assert _emu.inflate(stream, len(data) - 1)[0] != 0            # more data than ISIZE says
    assert _emu.inflate(stream, len(data) + 1)[0] != 0            # less
    assert _emu.inflate(stream[: len(stream) // 2], len(data))[0] != 0
    rnd = random.Random(3)
    bad = 0
    for k in range(300):
        junk = bytearray(stream)
        for _ in range(rnd.randrange(1, 6)):
            junk[rnd.randrange(len(junk))] ^= 1 << rnd.randrange(8)
        rc, out = _emu.inflate(bytes(junk), len(data))
        bad += rc != 0 or out != data
    assert bad > 250                                              # the rest flips bits the CRC would catch
    for k in range(200):
        junk = bytes(rnd.getrandbits(8) for _ in range(rnd.randrange(1, 400)))
        _emu.inflate(junk, rnd.randrange(0, 70000))               # must come back


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 63, 64, 65, 1000, 65279, 65280, 65536])
def test_crc_by_lanes_equals_zlib(n):
    rnd = random.Random(n)
    data = bytes(rnd.getrandbits(8) for _ in range(n))
    assert _emu.crc32(data) == zlib.crc32(data)


# ---- the record chain -------------------------------------------------------------------------------------------------
def _inflated_bam(n, seed=1, long_every=0):
    from tests.test_bam import FULL_HEADER
    from xenomapper_b200 import synth
    p, _ = synth.generate(n, seed=seed, style=synth.STYLE_PE_BOWTIE2)
    lines = bytes(p).decode().split("\n")
    if long_every:
        for k in range(0, len(lines) - 1, long_every):
            f = lines[k].split("\t")
            f[5], f[9], f[10] = "%dM" % 40000, "ACGT" * 10000, "I" * 40000       # a record longer than two segments
            lines[k] = "\t".join(f)
    raw = gzip.decompress(_bamwriter.sam_to_bam(FULL_HEADER, "\n".join(lines).encode(), level=1))
    l_text = struct.unpack_from("<i", raw, 4)[0]
    o = 8 + l_text
    n_ref = struct.unpack_from("<i", raw, o)[0]
    o += 4
    for _ in range(n_ref):
        o += 8 + struct.unpack_from("<i", raw, o)[0]
    return raw, o, n_ref


def _serial_chain(raw, first):
    rec, o = [], first
    while o + 4 <= len(raw):
        bs = struct.unpack_from("<I", raw, o)[0]
        if o + 4 + bs > len(raw):
            break
        rec.append(o)
        o += 4 + bs
    return rec, o


@pytest.mark.parametrize("seg", [512, 4096, 16384, 1 << 20])
@pytest.mark.parametrize("long_every", [0, 97])
def test_parallel_chain_equals_the_serial_one(seg, long_every):
    raw, first, n_ref = _inflated_bam(1500, long_every=long_every)
    want, want_end = _serial_chain(raw, first)
    got = _emu.bam_chain(raw, first, n_ref, seg)
    assert got is not None
    rec, end, repaired = got
    assert rec == want and end == want_end == len(raw)
    if seg >= 4096 and not long_every:
        assert repaired <= len(raw) // seg // 10 + 1           # the guesses are nearly always right


def test_parallel_chain_with_a_cut_record_at_the_end():
    raw, first, n_ref = _inflated_bam(400)
    for cut in (1, 3, 4, 5, 35, 36, 37, 100, 333):
        part = raw[:-cut]
        want, want_end = _serial_chain(part, first)
        rec, end, _ = _emu.bam_chain(part, first, n_ref, 4096)
        assert rec == want and end == want_end


def test_parallel_chain_reports_a_corrupt_length():
    raw, first, n_ref = _inflated_bam(400)
    want, _ = _serial_chain(raw, first)
    bad = bytearray(raw)
    bad[want[200]:want[200] + 4] = struct.pack("<I", 7)
    assert _emu.bam_chain(bytes(bad), first, n_ref, 4096) is None


# ---- the plan of the device BGZF compressor (csrc/xm_deflate.h, host half) ------------------------------------------------
def _kraft(lengths):
    from fractions import Fraction
    return sum(Fraction(1, 2 ** l) for l in lengths if l)


@pytest.mark.parametrize("name,data", SAMPLES[1:], ids=[s[0] for s in SAMPLES[1:]])
def test_deflate_plan_is_a_valid_dynamic_block(name, data):
    """the header bits the kernel copies in front of every member and the codes it looks up: zlib reads a block coded with them"""
    stream, ll, dl = _emu.deflate_literals(data[:4096], data)
    assert all(1 <= l <= 15 for l in ll) and all(1 <= l <= 15 for l in dl)          # every symbol stays encodable
    assert _kraft(ll) == 1 and _kraft(dl) == 1                                       # complete codes: no inflater objects
    assert zlib.decompress(stream, -15) == data
    rc, out = _emu.inflate(stream, len(data))                                        # and the device inflater
    assert rc == 0 and out == data


def test_deflate_plan_from_token_counts():
    rnd = random.Random(2)
    data = SAMPLES[8][1]
    for hist in ([0] * 316, [rnd.randrange(0, 5) ** 6 for _ in range(316)], [10 ** 9] * 316, [1 << k % 31 for k in range(316)]):
        stream, ll, dl = _emu.deflate_literals(b"", data, hist=hist)
        assert max(ll) <= 15 and max(dl) <= 15 and min(ll) >= 1 and min(dl) >= 1
        assert _kraft(ll) == 1 and _kraft(dl) == 1
        assert zlib.decompress(stream, -15) == data


def test_deflate_plan_fits_the_sample():
    data = SAMPLES[7][1]                                   # DNA: four letters
    stream, ll, dl = _emu.deflate_literals(data[:8192], data)
    assert max(ll[c] for c in b"ACGT") <= 3
    assert len(stream) < 0.3 * len(data)


@pytest.mark.parametrize("seg", [512, 4096])
def test_parallel_chain_of_a_part_of_the_stream(seg):
    """a rank's part of a file shared between GPUs: no record start is known, the records that start before `stop` are its own"""
    raw, first, n_ref = _inflated_bam(1500, long_every=211)
    want, _ = _serial_chain(raw, first)
    rnd = random.Random(seg)
    for _ in range(12):
        lo = rnd.randrange(first, len(raw) - 20000)
        hi = lo + rnd.randrange(1000, 150000)
        part = raw[lo:min(len(raw), hi + 120000)]                        # the part and what follows it (the last record's tail)
        stop = min(hi, len(raw)) - lo
        rec, end, _, guess = _emu.bam_chain(part, None, n_ref, seg, stop=stop, want_guess=True)
        own = [w - lo for w in want if lo <= w < lo + stop]
        if not own:                                                      # the part lies inside one long record
            continue
        if guess == own[0]:                                              # the guess is the true first record: everything follows
            assert rec == own
            nxt = [w - lo for w in want if w >= lo + stop]
            assert end == (nxt[0] if nxt and nxt[0] <= len(part) else end)
        # given the true entry (what the rank before reports), the part is exact whatever the guess was
        rec2, end2, _ = _emu.bam_chain(part, own[0], n_ref, seg, stop=stop)
        assert rec2 == own
