"""Runs the kernels' tile programs on the CPU (tests/emu) against the reference goldens.

The same source (xenomapper_b200/csrc/xm_tile.h, xm_parse.h, xm_walk.h) is what
nvcc compiles into the CUDA kernels; here each CTA is emulated thread by thread
so the parsing, line indexing, rank/offset bookkeeping and host orchestration
are checked on every golden case without a GPU.  Four variants: the production
32 KiB tiles and 1 KiB tiles (many tile boundaries even on tiny inputs), each
with the mask-driven fast parse and with the exact byte-wise parse forced.
"""
import pytest

from tests import _emu
from tests import _golden as G

VARIANTS = {"big": 0, "big_generic": 1, "small": 2, "small_generic": 3}


def check_case(case, debug):
    p, s = G.case_records(case)
    o = case["opts"]
    e = case["expect"]
    r = _emu.classify(p, s, mode=o["mode"], score_src=o["score_src"], skip_repeated=o["skip_repeated"],
                      min_score=o["min_score"], enabled_bins=o["enabled_bins"], debug=debug)
    if case["gpu"] == "unsupported":
        assert r["status"] == 5, r["message"]           # XM_ERR_UNSUPPORTED: detected, never mis-scored
        return
    assert r["status"] == G.ERR_CODE[e["error"]], r["message"]
    assert [len(x) for x in r["outputs"]] == e["records_len"]
    assert [G.sha(x) for x in r["outputs"]] == e["records_sha256"]
    if e["error"] is None:
        assert G.counts_dict(r["counts"], o["mode"]) == e["counts"]


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("case", G.CASES, ids=[c["name"] for c in G.CASES])
def test_emulated_tiles_match_reference(case, variant):
    if variant.startswith("small") and case["input"]["kind"] == "synth" and case["input"]["seed"] != 1:
        pytest.skip("1 KiB-tile emulation of the large synthetic cases is covered by seed 1")
    check_case(case, VARIANTS[variant])


@pytest.mark.parametrize("chunk", [600, 1000, 4096, 100000])
@pytest.mark.parametrize("case", [c for c in G.CASES if not (c["input"]["kind"] == "synth" and c["input"]["n"] > 3000)],
                         ids=lambda c: c["name"])
def test_emulated_chunked_walk_matches_reference(case, chunk):
    """the chunked walk (xm_stream.h: carry-over between chunks, halo record, end of walk) over the emulated kernels:
    chunks of 600 bytes cut almost every record pair apart, 100000 bytes is a single step for most cases"""
    p, s = G.case_records(case)
    o = case["opts"]
    e = case["expect"]
    r = _emu.classify(p, s, mode=o["mode"], score_src=o["score_src"], skip_repeated=o["skip_repeated"],
                      min_score=o["min_score"], enabled_bins=o["enabled_bins"], chunk=chunk)
    if case["gpu"] == "unsupported":
        assert r["status"] == 5, r["message"]
        return
    longest = max((len(x) for x in p.split(b"\n")), default=0)
    if longest + 1 > chunk:
        assert r["status"] in (5, G.ERR_CODE[e["error"]])      # a line that does not fit the staging buffer is refused, not mangled
        if r["status"] == 5:
            return
    assert r["status"] == G.ERR_CODE[e["error"]], r["message"]
    assert [len(x) for x in r["outputs"]] == e["records_len"]
    assert [G.sha(x) for x in r["outputs"]] == e["records_sha256"]
    if e["error"] is None:
        assert G.counts_dict(r["counts"], o["mode"]) == e["counts"]
        assert r["n_records"] == sum(e["counts"].values()) if o["mode"] == 0 else True


def _fixed_width_pair(n, width, first=0):
    """n records per stream, every primary line exactly `width` bytes (secondary width-8): tiles of 32 KiB then
    hold exactly 32768/width lines"""
    P, S = [], []
    for i in range(first, first + n):
        q = "r%07d" % (i // 2 if i % 5 == 0 else i)
        a1, a2 = 40 + (i * 7) % 60, 40 + (i * 11) % 60
        for out, a, w in ((P, a1, width), (S, a2, width - 8)):
            head = "%s\t0\tchr1\t%d\t42\t50M\t*\t0\t0\t" % (q, 1000 + i)
            tail = "\tAS:i:%d\tXS:i:%d\tNM:i:%d\n" % (a, a - (i % 3), i % 4)
            fill = w - len(head) - len(tail) - 1
            out.append(head + "A" * (fill // 2) + "\t" + "I" * (fill - fill // 2) + tail)
            assert len(out[-1]) == w
    return "".join(P).encode(), "".join(S).encode()


@pytest.mark.parametrize("width", [256, 128, 264, 248, 292, 294, 296])
@pytest.mark.parametrize("skip", [False, True])
def test_tiles_with_exactly_half_the_threads_in_lines(width, skip):
    """tiles that own exactly THREADS/2 lines and one more or less (128 lines of 256 bytes in 32 KiB tiles, 160 lines of
    294 bytes in 46 KiB tiles): the two-threads-per-line
    parse must leave the last thread to the halo line"""
    from oracle import oracle
    p, s = _fixed_width_pair(700, width)
    ref = oracle.classify(p, s, skip_repeated=skip)
    r = _emu.classify(p, s, skip_repeated=skip)
    assert r["status"] == 0, r["message"]
    assert r["counts"] == ref["counts"]
    assert r["outputs"] == ref["outputs"]


def _pair_with_widths(widths, first=0):
    """one record per entry of `widths` in both streams (primary `w` bytes, secondary `w - 8`)"""
    P, S = [], []
    for k, w in enumerate(widths):
        p, s = _fixed_width_pair(1, w, first=first + k)
        P.append(p); S.append(s)
    return b"".join(P), b"".join(S)


def test_more_secondary_records_than_the_sampled_estimate():
    """The compact per-record arrays are sized from the mean line length of the stream's first 256 KiB.  Long lines
    up front and many short lines behind them yield more records than that estimate plus its margin: the classify
    kernels must not look past the arrays (out-of-bounds read at classify_tile) and the host must grow
    them and walk again."""
    from oracle import oracle
    p, s = _pair_with_widths([1000] * 400 + [90] * 30000)
    ref = oracle.classify(p, s)
    r = _emu.classify(p, s)
    assert r["status"] == 0, r["message"]
    assert r["n_records"] == 30400
    assert r["counts"] == ref["counts"]
    assert r["outputs"] == ref["outputs"]


@pytest.mark.timeout(120)
@pytest.mark.parametrize("where", ["middle", "chunk_boundary"])
def test_chunked_walk_refuses_a_line_longer_than_the_chunk(where):
    """A line that fits no staging step must end the walk with XM_ERR_UNSUPPORTED after the records before it,
    not spin (from step 1 on the carried halo record hid the 'longer than the buffer' test)."""
    from oracle import oracle
    chunk = 256
    head = 40 if where == "middle" else 5       # 100-byte lines: the long one starts mid-chunk or right at a cut
    p, s = _pair_with_widths([100] * head + [1300] + [100] * 40)
    r = _emu.classify(p, s, chunk=chunk)
    assert r["status"] == 5, (r["status"], r["message"])
    assert "longer than the staging buffer" in r["message"]
    assert r["n_records"] <= head
    # the control: without the long line the same walk finishes and matches the oracle
    p2, s2 = _pair_with_widths([100] * (head + 40))
    r2 = _emu.classify(p2, s2, chunk=chunk)
    ref = oracle.classify(p2, s2)
    assert r2["status"] == 0 and r2["outputs"] == ref["outputs"] and r2["counts"] == ref["counts"]


def test_qname_assert_compares_bytes_not_only_hashes(tmp_path):
    """The cross-stream QNAME assert (xm.py:106) joins the streams by a 64-bit hash and then compares the bytes.  A build
    whose hash depends on the name's length only (-DXM_WEAK_HASH: every pair of equally long names collides) must
    still raise AssertionError at the first record whose names differ, and walk clean inputs unchanged."""
    import subprocess
    from oracle import oracle
    so = str(tmp_path / "libxm_emu_weak.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-DXM_WEAK_HASH", "-o", so, _emu.SRC])
    saved_so, saved_lib = _emu.SO, _emu._lib
    try:
        _emu.SO, _emu._lib = so, None
        os_utime = __import__("os").utime
        os_utime(so)                                  # newer than the sources: _emu.lib() loads it as it is
        p, s = _fixed_width_pair(400, 256)
        ref = oracle.classify(p, s)
        r = _emu.classify(p, s)
        assert r["status"] == 0 and r["outputs"] == ref["outputs"]
        bad = bytearray(s)
        at = bad.index(b"\n", len(bad) // 2) + 1
        bad[at + 3] = ord("X")                        # same length, different name: the weak hashes still agree
        ref = oracle.classify(p, bytes(bad))
        r = _emu.classify(p, bytes(bad))
        assert ref["err"] == 1 and r["status"] == 1, r["message"]
        assert r["outputs"] == ref["outputs"]
    finally:
        _emu.SO, _emu._lib = saved_so, saved_lib
