"""BAM input (BASELINE configs[4]): host BGZF inflate + GPU record rendering + the walk.

The reference cannot run its own BAM path here (samtools is absent, xm.py:48-93 are `# pragma: no cover`); the pin
is the fixture twins: paired_end_testdata_{human,mouse}.bam decode to exactly the .sam files next to them, so the
expected outputs are the SAM goldens (which the unmodified reference produced)."""
import io
import os

import pytest

from tests import _bamwriter
from tests import _golden as G


def test_bam_header_text_needs_no_gpu():
    from xenomapper_b200 import _lib
    for which in ("primary", "secondary"):
        bam = G.fixture_bytes("pe", which, "bam")
        header, _ = G.split_header(G.fixture_bytes("pe", which, "sam"))
        assert _lib.bam_header_text(bam) == header


def test_bam_writer_round_trips_the_fixture_through_the_header_reader():
    from xenomapper_b200 import _lib
    header, records = G.split_header(G.fixture_bytes("pe", "primary", "sam"))
    bam = _bamwriter.sam_to_bam(header, records + b"\n", block=3000)
    assert _lib.bam_header_text(bam) == header


def test_corrupt_bam_is_refused():
    from xenomapper_b200 import _lib
    bam = bytearray(G.fixture_bytes("pe", "primary", "bam"))
    with pytest.raises(_lib.XenomapperLibraryError):
        _lib.bam_header_text(bytes(bam[:100]))
    bam[40] ^= 0x55
    with pytest.raises(_lib.XenomapperLibraryError):
        _lib.bam_header_text(bytes(bam))


@pytest.fixture(scope="module")
def ctx():
    from xenomapper_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["primary", "secondary"])
def test_gpu_renders_fixture_bam_as_its_sam_twin(ctx, which):
    _, records = G.split_header(G.fixture_bytes("pe", which, "sam"))
    if not records.endswith(b"\n"):
        records += b"\n"                       # the fixture's last line is unterminated; samtools view prints whole lines
    assert ctx.bam_render_host(G.fixture_bytes("pe", which, "bam")) == records


@pytest.mark.gpu
@pytest.mark.parametrize("style", [0, 1, 2])
@pytest.mark.parametrize("block", [0xff00, 777])
def test_gpu_renders_synthetic_bam_back_to_its_sam(ctx, style, block):
    """records that straddle BGZF blocks (777-byte blocks cut every record), every aux type the writer emits"""
    from xenomapper_b200 import synth
    p, _ = synth.generate(4000 if block > 1000 else 400, seed=31 + style, style=style)
    extra = b"rX\t4\t*\t0\t0\t*\t*\t0\t0\t*\t*\tXA:A:q\tXB:B:c,-1,2,3\tXC:B:S,65535,0\tXH:H:1AE301\tXI:i:-70000\tXJ:i:4000000000\tXS:i:-200\tXZ:Z:a b\n"
    text = bytes(p) + extra + b"*\t4\t*\t0\t0\t*\t*\t0\t0\tACGTN\t*\n"
    bam = _bamwriter.sam_to_bam(FULL_HEADER, text, block=block)
    ctx.bam_stats(reset=True)
    assert ctx.bam_render_host(bam) == text
    st = ctx.bam_stats()
    assert st.records == text.count(b"\n") and st.text_bytes == len(text)


@pytest.mark.gpu
def test_float_aux_prints_like_the_c_library(ctx):
    """f and B:f values through the renderer against Python's "%g" of the same single-precision value (the C library's
    conversion; tests/test_fmtg.py checks the formatter itself on a million bit patterns, scripts/check_fmtg_all.sh on all)"""
    import random
    import struct
    import numpy as np
    rnd = random.Random(8)
    rec = _bamwriter.record("r1\t4\t*\t0\t0\t*\t*\t0\t0\tAC\tII", {})
    patterns = [rnd.getrandbits(32) for _ in range(400)] + [0x3f800000, 0x7f800000, 0xff800000, 0x00000001, 0x80000001, 0x7f7fffff, 0x49742400, 0x49742408, 0x3a83126f]
    vals = [np.frombuffer(struct.pack("<I", b), dtype=np.float32)[0] for b in patterns]
    vals = [v for v in vals if v == v]                      # "nan" carries a sign in some C libraries
    body = rec[4:] + b"XFf" + struct.pack("<f", vals[0]) + b"XGBf" + struct.pack("<I", len(vals)) + b"".join(struct.pack("<f", v) for v in vals)
    raw = b"BAM\1" + struct.pack("<i", 0) + struct.pack("<i", 0) + struct.pack("<i", len(body)) + body
    want = "r1\t4\t*\t0\t0\t*\t*\t0\t0\tAC\tII\tXF:f:%g\tXG:B:f,%s\n" % (float(vals[0]), ",".join("%g" % float(v) for v in vals))
    assert ctx.bam_render_host(_bamwriter.bgzf(raw)).decode() == want


# every reference name the synthetic generator uses (its own header lists two)
FULL_HEADER = "@HD\tVN:1.0\tSO:unsorted\n" + "".join("@SQ\tSN:%s\tLN:1000000000\n" % n for n in
                                                     ("chr1", "chr2", "chr7", "chr11", "chr17", "chr19", "chrX", "chrM")) + "@PG\tID:bowtie2\tPN:bowtie2\n"

FIXTURE_CASES = [c for c in G.CASES if c["input"]["kind"] == "fixture" and c["input"]["key"] == "pe" and c["header"]]


@pytest.mark.gpu
@pytest.mark.parametrize("case", FIXTURE_CASES, ids=[c["name"] for c in FIXTURE_CASES])
def test_bam_walk_equals_the_sam_goldens(case):
    """process_headers(bam=True) + getBamReadPairs + the walks on the fixture BAMs: the six outputs, headers
    included, and the counts are those the reference produced from the SAM twins"""
    from xenomapper_b200 import xenomapper as xm
    o = case["opts"]
    walk = (xm.main_single_end, xm.main_paired_end, xm.conservative_main_paired_end)[o["mode"]]
    tag_func = (xm.get_tag, xm.get_tag_with_ZS_as_XS, xm.get_cigarbased_AS_tag)[o["score_src"]]
    f1, f2 = io.BytesIO(G.fixture_bytes("pe", "primary", "bam")), io.BytesIO(G.fixture_bytes("pe", "secondary", "bam"))
    outs = [io.StringIO() if (o["enabled_bins"] >> b) & 1 else None for b in range(6)]
    names = ("primary_specific", "secondary_specific", "primary_multi", "secondary_multi", "unassigned", "unresolved")
    kw = dict(zip(names, outs))
    if kw["primary_specific"] is None:
        pytest.skip("the reference always writes the primary_specific header")
    xm.process_headers(f1, f2, bam=True, **kw)
    counts = walk(xm.getBamReadPairs(f1, f2, skip_repeated_reads=o["skip_repeated"]), min_score=o["min_score"], tag_func=tag_func, **kw)
    e = case["expect"]
    got = [x.getvalue().encode() if x is not None else b"" for x in outs]
    assert [len(x) for x in got] == e["full_len"]
    assert [G.sha(x) for x in got] == e["full_sha256"]
    key = (lambda k: k) if o["mode"] == 0 else (lambda k: k[0] + "|" + k[1])
    assert {key(k): v for k, v in counts.items()} == e["counts"]


@pytest.mark.gpu
@pytest.mark.parametrize("style,mode,score,skip,min_score", [(0, 0, 0, True, float("-inf")), (1, 1, 2, False, -40.0), (2, 2, 1, False, -18.0)],
                         ids=["se_skip", "pe_cigar", "pe_conservative_zs"])
def test_bam_walk_on_synthetic_pairs_equals_the_oracle(ctx, style, mode, score, skip, min_score):
    """configs[4]-shaped inputs (--cigar_scores on BAM among them): the walk on BAM input against the oracle on the SAM text"""
    from oracle import oracle
    from xenomapper_b200 import _lib, synth
    p, s = synth.generate(20000, seed=41, style=style)
    hdr2 = FULL_HEADER.replace("SN:chr", "SN:").replace("SN:M\t", "SN:MT\t")
    bam_p, bam_s = _bamwriter.sam_to_bam(FULL_HEADER, bytes(p)), _bamwriter.sam_to_bam(hdr2, bytes(s))
    ref = oracle.classify(p, s, mode=mode, score_src=score, skip_repeated=skip, min_score=min_score)
    rc, res, outs = ctx.classify_bam_host(bam_p, bam_s, _lib.Context.opts(mode, score, skip, min_score))
    assert rc == 0, ctx.error()
    assert list(res.counts) == ref["counts"]
    assert outs == ref["outputs"]


def _literal_bam(aux_bytes):
    """A one-record BAM built byte by byte from the layout tables of the SAM/BAM specification (section 4.2: the
    alignment record, 4.2.4: auxiliary data) -- literal hex for every aux value, no encoder shared with
    tests/_bamwriter.py.  Record: read `r1`, unmapped (refID -1, pos -1, FLAG 4), l_seq 4 = ACGT, quals 30 31 32 33."""
    import struct
    import zlib
    read_name = b"r1\0"
    core = struct.pack("<iiBBHHHIiii", -1, -1, len(read_name), 0, 4680, 0, 4, 4, -1, -1, 0)   # refID pos l_read_name mapq bin n_cigar flag l_seq next_refID next_pos tlen
    seq = bytes([0x12, 0x48])                   # A=1 C=2 | G=4 T=8 (section 4.2.3: "=ACMGRSVTWYHKDBN")
    qual = bytes([30, 31, 32, 33])
    body = core + read_name + seq + qual + aux_bytes
    rec = struct.pack("<I", len(body)) + body
    text = b"@HD\tVN:1.6\n"
    payload = b"BAM\1" + struct.pack("<I", len(text)) + text + struct.pack("<I", 0) + rec
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    data = co.compress(payload) + co.flush()
    block = (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(data) + 25) + data +
             struct.pack("<II", zlib.crc32(payload), len(payload)))
    return block + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


# aux bytes (tag, type, little-endian value) -> the SAM text the specification prescribes (section 1.5: every
# integer type prints as `i`; section 4.2.4: A one character, Z / H NUL-terminated, B: sub-type, count, values)
AUX_VECTORS = [
    (bytes.fromhex("5841 63 fe"), "XA:i:-2"),                                   # c  int8
    (bytes.fromhex("5842 43 ff"), "XB:i:255"),                                  # C  uint8
    (bytes.fromhex("5843 73 0080"), "XC:i:-32768"),                             # s  int16
    (bytes.fromhex("5844 53 ffff"), "XD:i:65535"),                              # S  uint16
    (bytes.fromhex("5845 69 ffffffff"), "XE:i:-1"),                             # i  int32
    (bytes.fromhex("5846 69 00000080"), "XF:i:-2147483648"),
    (bytes.fromhex("5847 49 ffffffff"), "XG:i:4294967295"),                     # I  uint32
    (bytes.fromhex("5848 41 7a"), "XH:A:z"),                                    # A  printable character
    (bytes.fromhex("5849 5a 68656c6c6f20776f726c64 00"), "XI:Z:hello world"),   # Z  string with a space
    (bytes.fromhex("584a 48 314145333031 00"), "XJ:H:1AE301"),                  # H  hex string (the specification's own example value)
    (bytes.fromhex("584b 42 63 03000000 01ff7f"), "XK:B:c,1,-1,127"),           # B:c
    (bytes.fromhex("584c 42 43 02000000 00ff"), "XL:B:C,0,255"),                # B:C
    (bytes.fromhex("584d 42 73 02000000 0080ff7f"), "XM:B:s,-32768,32767"),     # B:s
    (bytes.fromhex("584e 42 53 01000000 ffff"), "XN:B:S,65535"),                # B:S
    (bytes.fromhex("584f 42 69 02000000 ffffffff00000080"), "XO:B:i,-1,-2147483648"),   # B:i
    (bytes.fromhex("5850 42 49 01000000 ffffffff"), "XP:B:I,4294967295"),       # B:I
    (bytes.fromhex("5851 42 43 00000000"), "XQ:B:C"),                           # an empty array
    # floats print as samtools prints them, "%g" (htslib sam_format1): IEEE-754 single precision, little endian
    (bytes.fromhex("5852 66 0000c03f"), "XR:f:1.5"),                            # f  1.5
    (bytes.fromhex("5853 66 cdcccc3d"), "XS:f:0.1"),                            #    0.1f is 0.100000001490116: six digits
    (bytes.fromhex("5854 66 00247449"), "XT:f:1e+06"),                          #    1000000 takes the exponent form
    (bytes.fromhex("5855 66 17b7d1b8"), "XU:f:-0.0001"),                        #    -1e-4f: the last fixed-notation exponent
    (bytes.fromhex("5856 66 acc52737"), "XV:f:1e-05"),                          #    1e-5f
    (bytes.fromhex("5857 66 ffff7f7f"), "XW:f:3.40282e+38"),                    #    FLT_MAX
    (bytes.fromhex("5858 66 01000000"), "XX:f:1.4013e-45"),                     #    the smallest denormal
    (bytes.fromhex("5859 66 00000080"), "XY:f:-0"),                             #    minus zero
    (bytes.fromhex("585a 42 66 03000000 0000803f 0000807f 79e9f642"), "XZ:B:f,1,inf,123.456"),      # B:f
]


@pytest.mark.gpu
def test_gpu_renders_literal_aux_vectors(ctx):
    """every aux type the renderer prints, from hand-written bytes: pins c s S i I A H B:* (the fixture twins only
    hold C and Z) against the specification's tables rather than against the test-side writer"""
    for raw, text in AUX_VECTORS:
        line = ctx.bam_render_host(_literal_bam(raw))
        assert line == b"r1\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\t?@AB\t" + text.encode() + b"\n", (raw.hex(), line)
    allraw = b"".join(r for r, _ in AUX_VECTORS)
    alltext = "\t".join(t for _, t in AUX_VECTORS)
    assert ctx.bam_render_host(_literal_bam(allraw)) == b"r1\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\t?@AB\t" + alltext.encode() + b"\n"
    assert ctx.bam_render_host(_literal_bam(b"")) == b"r1\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\t?@AB\n"


def test_literal_bam_is_well_formed():
    """the hand-built container inflates with Python's gzip and carries the record it claims (no GPU)"""
    import gzip
    import struct
    raw = gzip.decompress(_literal_bam(AUX_VECTORS[0][0]))
    assert raw[:4] == b"BAM\1"
    l_text = struct.unpack_from("<I", raw, 4)[0]
    assert raw[8:8 + l_text] == b"@HD\tVN:1.6\n"
    block_size = struct.unpack_from("<I", raw, 8 + l_text + 4)[0]
    assert 8 + l_text + 4 + 4 + block_size == len(raw)


# ---- BGZF inflate and record chain on the device (csrc/xm_inflate.h, csrc/xm_bamchain.h) ---------------------------------
def _long_text(n=6000, seed=77):
    from xenomapper_b200 import synth
    p, _ = synth.generate(n, seed=seed, style=1)
    lines = bytes(p).split(b"\n")
    f = lines[n // 2].split(b"\t")
    f[5], f[9], f[10] = b"70000M", b"ACGT" * 17500, b"I" * 70000          # one record longer than a BGZF block and a small window
    lines[n // 2] = b"\t".join(f)
    return b"\n".join(lines)


@pytest.mark.gpu
@pytest.mark.parametrize("window,seg", [(70000, 512), (300000, 4096), (1 << 30, 16384)])
@pytest.mark.parametrize("level", [0, 1, 9])
def test_device_inflate_in_windows_equals_the_text(ctx, monkeypatch, window, seg, level):
    """windows smaller than the file (records, and one giant record, straddle them), segments smaller than a record,
    stored / fast / best DEFLATE"""
    text = _long_text()
    bam = _bamwriter.sam_to_bam(FULL_HEADER, text, level=level)
    monkeypatch.setenv("XM_BAM_WINDOW", str(window))
    monkeypatch.setenv("XM_BAM_SEG", str(seg))
    ctx.bam_stats(reset=True)
    assert ctx.bam_render_host(bam) == text
    st = ctx.bam_stats()
    assert st.records == text.count(b"\n") and st.inflated_bytes > len(text) // 3
    assert st.inflate_ms > 0 or os.environ.get("XM_BAM_INFLATE") == "host"
    monkeypatch.setenv("XM_BAM_INFLATE", "host")
    assert ctx.bam_render_host(bam) == text                            # zlib on the host: the same text


@pytest.mark.gpu
def test_device_inflate_with_a_header_longer_than_the_window(ctx, monkeypatch):
    header = "@HD\tVN:1.0\n" + "".join("@SQ\tSN:contig_%06d\tLN:%d\n" % (k, 1000 + k) for k in range(9000))
    text = b"".join(b"r%d\t0\tcontig_%06d\t%d\t30\t4M\t*\t0\t0\tACGT\tIIII\tAS:i:-%d\n" % (k, k % 9000, k + 1, k % 7 + 1) for k in range(3000))
    bam = _bamwriter.sam_to_bam(header, text)
    monkeypatch.setenv("XM_BAM_WINDOW", "65536")
    assert ctx.bam_render_host(bam) == text


@pytest.mark.gpu
def test_device_inflate_refuses_damaged_blocks(ctx):
    from xenomapper_b200 import _lib
    text = _long_text(2000)
    bam = bytearray(_bamwriter.sam_to_bam(FULL_HEADER, text, level=6))
    assert ctx.bam_render_host(bytes(bam)) == text
    second = int.from_bytes(bam[16:18], "little") + 1                    # BSIZE of the first block: the second starts behind it
    for at in (second + 18 + 40, second + 18 + 400):                     # inside the second block's DEFLATE data
        bad = bytearray(bam)
        bad[at] ^= 0x10
        with pytest.raises(_lib.XenomapperLibraryError, match="does not inflate"):
            ctx.bam_render_host(bytes(bad))
    bsize2 = int.from_bytes(bam[second + 16:second + 18], "little") + 1
    bad = bytearray(bam)
    bad[second + bsize2 - 8] ^= 1                                        # the block's CRC-32
    with pytest.raises(_lib.XenomapperLibraryError, match="does not inflate"):
        ctx.bam_render_host(bytes(bad))
    assert ctx.bam_render_host(bytes(bam)) == text                       # the context is still usable


@pytest.mark.gpu
def test_bam_walk_with_small_windows_equals_the_big_one(ctx, monkeypatch):
    from xenomapper_b200 import _lib, synth
    p, s = synth.generate(30000, seed=43, style=1)
    hdr2 = FULL_HEADER.replace("SN:chr", "SN:").replace("SN:M\t", "SN:MT\t")
    bam_p, bam_s = _bamwriter.sam_to_bam(FULL_HEADER, bytes(p), level=6), _bamwriter.sam_to_bam(hdr2, bytes(s), level=6)
    opts = _lib.Context.opts(_lib.MODE_PE_LIBERAL, _lib.SCORE_CIGAR_NM, False, -40.0)
    rc, res, outs = ctx.classify_bam_host(bam_p, bam_s, opts)
    assert rc == 0, ctx.error()
    rc2, res2, outs2 = ctx.classify_host(p, s, opts)
    assert outs == outs2 and list(res.counts) == list(res2.counts)
    monkeypatch.setenv("XM_BAM_WINDOW", "200000")
    monkeypatch.setenv("XM_CHUNK_BYTES", "300000")
    rc3, res3, outs3 = ctx.classify_bam_host(bam_p, bam_s, opts)
    assert rc3 == 0, ctx.error()
    assert outs3 == outs and list(res3.counts) == list(res.counts)


@pytest.mark.gpu
def test_device_inflate_edge_shapes(ctx, monkeypatch):
    """no records at all; empty BGZF members between the data; no end-of-file member; a record cut off by the end of the file"""
    from xenomapper_b200 import _lib, synth
    empty = _bamwriter.sam_to_bam(FULL_HEADER, b"")
    assert ctx.bam_render_host(empty) == b""
    p, _ = synth.generate(3000, seed=5, style=0)
    text = bytes(p)
    bam = _bamwriter.sam_to_bam(FULL_HEADER, text, block=5000)
    eof = bam[-28:]
    assert eof == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    # split the members, put empty ones in between
    ms, at = [], 0
    while at < len(bam):
        bs = int.from_bytes(bam[at + 16:at + 18], "little") + 1
        ms.append(bam[at:at + bs]); at += bs
    holes = b"".join(m + (eof if k % 7 == 3 else b"") for k, m in enumerate(ms[:-1]))
    for window in ("20000", str(1 << 30)):
        monkeypatch.setenv("XM_BAM_WINDOW", window)
        assert ctx.bam_render_host(holes + eof) == text
        assert ctx.bam_render_host(holes) == text                       # samtools warns about the missing EOF marker and goes on
        with pytest.raises(_lib.XenomapperLibraryError, match="truncated"):
            ctx.bam_render_host(b"".join(ms[:-3]) + eof)                # the last record is cut by the missing members


@pytest.mark.gpu
def test_bam_walk_survives_an_error_and_goes_on(ctx):
    """a failed call leaves the context usable and the next BAM walk correct"""
    from xenomapper_b200 import _lib, synth
    p, s = synth.generate(4000, seed=6, style=0)
    hdr2 = FULL_HEADER.replace("SN:chr", "SN:").replace("SN:M\t", "SN:MT\t")
    bam_p, bam_s = _bamwriter.sam_to_bam(FULL_HEADER, bytes(p)), _bamwriter.sam_to_bam(hdr2, bytes(s))
    opts = _lib.Context.opts(_lib.MODE_SE, _lib.SCORE_AS_XS, True, float("-inf"))
    bad = bytearray(bam_s)
    bad[len(bad) // 2] ^= 0x40
    with pytest.raises(_lib.XenomapperLibraryError):
        ctx.classify_bam_host(bam_p, bytes(bad), opts)
    rc, res, outs = ctx.classify_bam_host(bam_p, bam_s, opts)
    rc2, res2, outs2 = ctx.classify_host(p, s, opts)
    assert rc == 0 and outs == outs2 and list(res.counts) == list(res2.counts)


# ---- BAM input of the walk across GPUs: a rank's part of a file (xm_bam_shard_*) -------------------------------------------
def _parts_text(ctx, bam, world, stream=0):
    """every rank's part of `bam` rendered one after the other on this GPU, the chains joined the way sharded.py does"""
    none = ctx.NONE64
    cur, texts, repaired = None, [], 0
    for r in range(world):
        guess, exit_off = ctx.bam_shard_open(stream, bam, r, world)
        if cur is not None and guess != cur:
            exit_off = ctx.bam_shard_chain(stream, cur)
            repaired += guess != none
            guess = cur
        if guess != none:
            cur = exit_off
        d, n = ctx.bam_shard_text(stream, 4096, 4096)
        texts.append(ctx.d2h(d, n))
    return texts, repaired


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 3, 8, 37])
@pytest.mark.parametrize("block", [0xff00, 3000])
def test_bam_parts_concatenate_to_the_whole_text(ctx, world, block):
    """records straddle the parts' block ranges, one record is longer than a whole part, many parts hold no record start"""
    text = _long_text(3000)
    bam = _bamwriter.sam_to_bam(FULL_HEADER, text, block=block, level=1)
    parts, repaired = _parts_text(ctx, bam, world)
    assert b"".join(parts) == text
    assert all(p == b"" or p.endswith(b"\n") for p in parts)
    if world <= 8 and block == 0xff00:
        assert sum(1 for p in parts if p) >= min(world, 3)


@pytest.mark.gpu
def test_bam_parts_with_a_header_longer_than_a_part(ctx):
    header = "@HD\tVN:1.0\n" + "".join("@SQ\tSN:contig_%06d\tLN:%d\n" % (k, 1000 + k) for k in range(9000))
    text = b"".join(b"r%d\t0\tcontig_%06d\t%d\t30\t4M\t*\t0\t0\tACGT\tIIII\tAS:i:-%d\n" % (k, k % 9000, k + 1, k % 7 + 1) for k in range(3000))
    bam = _bamwriter.sam_to_bam(header, text)
    for world in (2, 5, 16):
        parts, _ = _parts_text(ctx, bam, world)
        assert b"".join(parts) == text


@pytest.mark.gpu
def test_sharded_bam_walk_on_one_rank_equals_the_sam_walk(ctx):
    from xenomapper_b200 import _lib, sharded, synth
    p, s = synth.generate(20000, seed=44, style=1)
    hdr2 = FULL_HEADER.replace("SN:chr", "SN:").replace("SN:M\t", "SN:MT\t")
    bam_p, bam_s = _bamwriter.sam_to_bam(FULL_HEADER, bytes(p)), _bamwriter.sam_to_bam(hdr2, bytes(s))
    sharded.init_comm(ctx, 0, 1)
    res = sharded.sharded_bam_walk(ctx, 0, 1, None, bam_p, bam_s, mode=_lib.MODE_PE_LIBERAL, score_src=_lib.SCORE_CIGAR_NM, min_score=-40.0)
    assert res["status"] == 0, res["message"]
    rc, r2, outs2 = ctx.classify_host(p, s, _lib.Context.opts(_lib.MODE_PE_LIBERAL, _lib.SCORE_CIGAR_NM, False, -40.0))
    assert res["outputs"] == outs2 and res["counts"] == list(r2.counts)
