"""GPU parity: the CUDA path, called through the C ABI, against the reference goldens and the oracle."""
import pytest

from tests import _golden as G

pytestmark = pytest.mark.gpu

ERR = {0: None, 1: "AssertionError", 2: "ValueError", 3: "RuntimeError", 4: "UnicodeDecodeError", 5: "Unsupported"}


@pytest.fixture(scope="module")
def ctx():
    from xenomapper_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def run_case(ctx, case, debug):
    from xenomapper_b200 import _lib
    ctx.set_debug(debug)
    p, s = G.case_records(case)
    o = case["opts"]
    opts = _lib.Context.opts(o["mode"], o["score_src"], o["skip_repeated"], o["min_score"], o["enabled_bins"])
    rc, res, outs = ctx.classify_host(p, s, opts)
    e = case["expect"]
    if case["gpu"] == "unsupported":
        assert rc == _lib.XM_ERR_UNSUPPORTED, ctx.error()
        return
    assert ERR[rc] == e["error"], ctx.error()
    assert [len(x) for x in outs] == e["records_len"]
    assert [G.sha(x) for x in outs] == e["records_sha256"]
    if e["error"] is None:
        assert G.counts_dict(list(res.counts), o["mode"]) == e["counts"]


@pytest.mark.parametrize("debug", [0, 1, 2, 3, 4], ids=["big", "big_generic", "small", "small_generic", "rows"])
@pytest.mark.parametrize("case", G.CASES, ids=[c["name"] for c in G.CASES])
def test_cuda_matches_reference_golden(ctx, case, debug):
    """debug 4 (XM_DEBUG_ROWS): clean cases walk over rows (both streams through k_scan2, then k_size / k_prefix /
    k_emit -- the kernels of the sharded walk); every other case must be handed to the exact pair"""
    if debug in (2, 3) and case["input"]["kind"] == "synth" and case["input"]["seed"] != 1:
        pytest.skip("covered by seed 1")
    run_case(ctx, case, debug)


@pytest.mark.parametrize("chunk", [700, 5000])
@pytest.mark.parametrize("case", [c for c in G.CASES if c["name"] != "adv_very_long_lines"], ids=lambda c: c["name"])
def test_cuda_chunked_walk_matches_reference_golden(ctx, case, chunk, monkeypatch):
    """xm_classify_host in many small steps (XM_CHUNK_BYTES): staging, carry-over, halo records and ordered bins on the device"""
    if case["input"]["kind"] == "synth" and (case["input"]["seed"] != 1 or chunk < 5000):
        pytest.skip("covered by seed 1 at the larger chunk")
    monkeypatch.setenv("XM_CHUNK_BYTES", str(chunk))
    run_case(ctx, case, 0)


@pytest.mark.parametrize("width", [256, 128, 264, 248, 292, 294, 296])
@pytest.mark.parametrize("mode,skip", [(0, False), (0, True), (1, False)])
def test_cuda_fixed_width_lines_match_oracle(ctx, width, mode, skip):
    """32 KiB tiles that own exactly 128, 127, 129 or 256 lines (tests/test_emu_tiles.py has the CPU twin)"""
    from oracle import oracle
    from tests.test_emu_tiles import _fixed_width_pair
    from xenomapper_b200 import _lib
    ctx.set_debug(0)
    p, s = _fixed_width_pair(3000, width)
    ref = oracle.classify(p, s, mode=mode, skip_repeated=skip)
    rc, res, outs = ctx.classify_host(p, s, _lib.Context.opts(mode, 0, skip))
    assert rc == 0, ctx.error()
    assert list(res.counts) == ref["counts"]
    assert outs == ref["outputs"]


@pytest.mark.parametrize("rows", [0, 4], ids=["fused", "rows"])
@pytest.mark.parametrize("width", [600, 632, 648, 700, 900, 1000])
@pytest.mark.parametrize("mode,skip", [(0, False), (0, True), (1, False)])
def test_cuda_span_kernels_find_the_line_before_the_span(ctx, width, mode, skip, rows):
    """walks that look at the line before a span scan 640 bytes before it first and the whole 1 KiB only when that
    line starts earlier (xm_scan2.cuh span_front): lines just below and above that length, on the barrier-free pair"""
    from oracle import oracle
    from tests.test_emu_tiles import _fixed_width_pair
    from xenomapper_b200 import _lib
    ctx.set_debug(rows)
    p, s = _fixed_width_pair(4000, width)
    ref = oracle.classify(p, s, mode=mode, skip_repeated=skip)
    rc, res, outs = ctx.classify_host(p, s, _lib.Context.opts(mode, 0, skip))
    assert rc == 0, ctx.error()
    assert ctx.walk_kernels() == (["k_scan2", "k_size+k_prefix+k_emit"] if rows else ["k_scan2", "k_classify2"])
    assert list(res.counts) == ref["counts"]
    assert outs == ref["outputs"]


@pytest.mark.parametrize("rows", [0, 4], ids=["fused", "rows"])
@pytest.mark.parametrize("style,mode,skip", [(0, 0, True), (1, 1, False), (2, 2, False)])
def test_cuda_large_synthetic_matches_oracle(ctx, style, mode, skip, rows):
    """300 k records (about 130 MB per stream: thousands of tiles, look-back across several waves of CTAs)"""
    from oracle import oracle
    from xenomapper_b200 import _lib, synth
    ctx.set_debug(rows)
    p, s = synth.generate(300000, seed=77, style=style)
    score = 1 if style == 2 else 0
    ref = oracle.classify(p, s, mode=mode, score_src=score, skip_repeated=skip, min_score=-18.0 if style == 2 else float("-inf"))
    rc, res, outs = ctx.classify_host(p, s, _lib.Context.opts(mode, score, skip, -18.0 if style == 2 else float("-inf")))
    assert rc == 0, ctx.error()
    assert list(res.counts) == ref["counts"]
    assert outs == ref["outputs"]


@pytest.mark.parametrize("style,mode", [(0, 0), (1, 1)], ids=["single_end", "paired"])
def test_cuda_size_independent_properties(ctx, style, mode):
    """1.5 M records per stream (1.2 GB, beyond what the oracle is asked to chew in a test): properties that hold
    at any size.  (1) the walk of a concatenation is the concatenation of the walks (cut at a pair boundary);
    (2) every bin holds as many lines as its categories counted (xm.py:330-350; unresolved emits both streams);
    (3) the bytes read are the bytes given."""
    from xenomapper_b200 import _lib, synth
    import numpy as np
    ctx.set_debug(0)
    n = 1_500_000
    p, s = synth.generate(n, seed=5, style=style)
    half = n // 2
    pa, sa = synth.generate(half, seed=5, style=style)
    o = _lib.Context.opts(mode, 0, False)
    rc, res, outs = ctx.classify_host(p, s, o)
    assert rc == 0, ctx.error()
    assert int(res.bytes_in[0]) == p.nbytes and int(res.bytes_in[1]) == s.nbytes and int(res.n_records) == n
    rc1, r1, o1 = ctx.classify_host(pa, sa, o)
    rc2, r2, o2 = ctx.classify_host(p[pa.nbytes:], s[sa.nbytes:], o)
    assert rc1 == 0 and rc2 == 0
    assert [a + b for a, b in zip(o1, o2)] == outs
    assert [a + b for a, b in zip(r1.counts, r2.counts)] == list(res.counts)
    lines = [np.count_nonzero(np.frombuffer(x, dtype=np.uint8) == 10) for x in outs]
    c = list(res.counts)
    if mode == 0:
        expect = [c[0], c[1], c[2], c[3], c[4], 2 * c[5]]
        assert sum(c) == n
    else:
        # liberal chain xm.py:423-448: the better ranked of (fwd, rev) in PS > SS > PM > SM > UR > UA; 2 lines per unit, 4 for UR
        rank = {0: 0, 1: 1, 2: 2, 3: 3, 5: 4, 4: 5}
        expect = [0] * 6
        for f in range(6):
            for r in range(6):
                b = f if rank[f] <= rank[r] else r
                expect[b] += c[f * 6 + r] * (4 if b == 5 else 2)
        assert sum(c) == n // 2                     # every record pair is one unit in the interlaced synthetic data
    assert lines == expect


def test_cuda_kernel_selection_and_fallback(ctx):
    """clean, error-free input runs the barrier-free pair (k_scan2 / k_classify2); anything they do not handle --
    here a space inside a record, then a duplicated AS tag -- sends the walk through the exact pair, with the
    reference's result either way"""
    from oracle import oracle
    from xenomapper_b200 import _lib, synth
    ctx.set_debug(0)
    p, s = synth.generate(50000, seed=9, style=0)
    o = _lib.Context.opts(0, 0, True)
    rc, res, outs = ctx.classify_host(p, s, o)
    assert rc == 0 and ctx.walk_kernels() == ["k_scan2", "k_classify2"]
    ref = oracle.classify(p, s, skip_repeated=True)
    assert outs == ref["outputs"] and list(res.counts) == ref["counts"]
    # a space in the secondary stream: k_scan2 declines, k_scan runs; the primary stream is still clean
    sd = bytes(s).replace(b"\tXM:i:", b" XM:i:", 1)
    rc, res, outs = ctx.classify_host(p, sd, o)
    ref = oracle.classify(p, sd, skip_repeated=True)
    assert rc == 0 and "k_scan" in ctx.walk_kernels()
    assert outs == ref["outputs"] and list(res.counts) == ref["counts"]
    # a space in the primary stream: k_classify2 declines
    pd = bytes(p).replace(b"\tXM:i:", b" XM:i:", 1)
    rc, res, outs = ctx.classify_host(pd, s, o)
    ref = oracle.classify(pd, s, skip_repeated=True)
    assert rc == 0 and "k_classify" in ctx.walk_kernels()
    assert outs == ref["outputs"] and list(res.counts) == ref["counts"]
    # a backspace right after a tab: the one byte pair the span kernels' three-operation tab mask misreads
    # (xm_parse.h tab_mask_loose); two touching W bytes, so the span declines and the exact kernel runs
    for prim, sec, kern in ((bytes(p).replace(b"\tXM:i:", b"\t\x08XM:i:", 1), s, "k_classify"),
                            (p, bytes(s).replace(b"\tXM:i:", b"\t\x08XM:i:", 1), "k_scan")):
        rc, res, outs = ctx.classify_host(prim, sec, o)
        ref = oracle.classify(prim, sec, skip_repeated=True)
        assert rc == 0 and kern in ctx.walk_kernels()
        assert outs == ref["outputs"] and list(res.counts) == ref["counts"]
    # an input error (two AS tags): reported by the exact kernels, outputs up to the failing record
    pe = bytes(p)
    cut = pe.index(b"\n", len(pe) // 2) + 1
    line_end = pe.index(b"\n", cut)
    pe = pe[:line_end] + b"\tAS:i:1" + pe[line_end:]
    rc, res, outs = ctx.classify_host(pe, s, o)
    ref = oracle.classify(pe, s, skip_repeated=True)
    assert ERR[rc] == "ValueError" and ref["err"] == 2
    assert outs == ref["outputs"]


def test_cuda_row_walk_selection_and_fallback(ctx):
    """XM_DEBUG_ROWS: clean input walks over rows; a dirty line in either stream, a QNAME that differs in one byte
    (same length, so only the hash and the byte compare of k_size can tell) or an input error hands the whole walk
    to the exact pair, with the reference's result either way"""
    from oracle import oracle
    from xenomapper_b200 import _lib, synth
    ctx.set_debug(_lib.DEBUG_ROWS)
    for style, mode, skip in ((0, 0, True), (1, 1, False), (1, 2, False)):
        p, s = synth.generate(40000, seed=21, style=style)
        o = _lib.Context.opts(mode, 0, skip)
        rc, res, outs = ctx.classify_host(p, s, o)
        assert rc == 0 and ctx.walk_kernels() == ["k_scan2", "k_size+k_prefix+k_emit"]
        ref = oracle.classify(p, s, mode=mode, skip_repeated=skip)
        assert outs == ref["outputs"] and list(res.counts) == ref["counts"]
        for prim, sec in ((bytes(p).replace(b"\tXM:i:", b" XM:i:", 1), s), (p, bytes(s).replace(b"\tXM:i:", b" XM:i:", 1))):
            rc, res, outs = ctx.classify_host(prim, sec, o)
            ref = oracle.classify(prim, sec, mode=mode, skip_repeated=skip)
            assert rc == 0 and "k_classify" in ctx.walk_kernels()
            assert outs == ref["outputs"] and list(res.counts) == ref["counts"]
        # one QNAME byte changed in the middle of the secondary stream: AssertionError at that record, the prefix before it
        sb = bytearray(bytes(s))
        at = sb.index(b"\n", len(sb) // 2) + 1
        sb[at + 5] = ord("Z") if sb[at + 5] != ord("Z") else ord("Y")
        rc, res, outs = ctx.classify_host(p, bytes(sb), o)
        ref = oracle.classify(p, bytes(sb), mode=mode, skip_repeated=skip)
        assert ERR[rc] == "AssertionError" and ref["err"] == 1
        assert outs == ref["outputs"]
    ctx.set_debug(0)


def test_cuda_qname_assert_compares_bytes(tmp_path):
    """the weak-hash twin of the library (every pair of equally long QNAMEs collides): the exact kernels, and the two
    barrier-free walks with XM_DEBUG_EXACT_NAMES, must still raise AssertionError where two names differ, and walk
    clean input unchanged"""
    import json
    import os
    import subprocess
    import sys
    from xenomapper_b200 import _lib
    weak = os.path.join(os.path.dirname(_lib.LIB_PATH), "libxenomapper_b200_weakhash.so")
    assert os.path.exists(weak), "python -m xenomapper_b200.build makes it"
    script = r'''
import json, sys
from oracle import oracle
from xenomapper_b200 import _lib, synth
ctx = _lib.Context(0)
out = {}
for name, debug in (("fused", 8), ("rows", 4 | 8), ("exact", 1)):          # 8 = XM_DEBUG_EXACT_NAMES
    ctx.set_debug(debug)
    p, s = synth.generate(30000, seed=31, style=1)
    o = ctx.opts(1, 0, False)
    rc, res, outs = ctx.classify_host(p, s, o)
    ref = oracle.classify(p, s, mode=1)
    clean_ok = rc == 0 and outs == ref["outputs"] and list(res.counts) == ref["counts"]
    sb = bytearray(bytes(s))
    at = sb.index(b"\n", len(sb) // 2) + 1
    sb[at + 5] = ord("Z") if sb[at + 5] != ord("Z") else ord("Y")
    rc, res, outs = ctx.classify_host(p, bytes(sb), o)
    ref = oracle.classify(p, bytes(sb), mode=1)
    out[name] = dict(clean_ok=clean_ok, rc=rc, ref_err=ref["err"], prefix_ok=outs == ref["outputs"], kernels=ctx.walk_kernels())
print(json.dumps(out))
'''
    r = subprocess.run([sys.executable, "-c", script], cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       env=dict(os.environ, XM_LIB_PATH=weak), capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    res = json.loads(r.stdout.decode().strip().splitlines()[-1])
    for name, v in res.items():
        assert v["clean_ok"], (name, v)
        assert v["rc"] == 1 and v["ref_err"] == 1 and v["prefix_ok"], (name, v)


@pytest.mark.parametrize("rows", [0, 4], ids=["fused", "rows"])
@pytest.mark.parametrize("mode,skip", [(0, True), (1, False)])
def test_cuda_short_read_lines_stay_on_the_span_kernels(mode, skip, rows):
    """lines under ~190 bytes overflow the 64 lines a 12 KiB span holds and used to send the whole call to the exact
    kernels.  Streams of short lines take spans of half the size -- chosen from a sample on a context's first walk, and switched
    to when a span overflows on a later one."""
    from oracle import oracle
    from tests.test_emu_tiles import _fixed_width_pair
    from xenomapper_b200 import _lib
    fast = ["k_scan2", "k_size+k_prefix+k_emit"] if rows else ["k_scan2", "k_classify2"]
    opts = _lib.Context.opts(mode, 0, skip)

    def walk(c, width, n=6000):
        p, s = _fixed_width_pair(n, width)
        ref = oracle.classify(p, s, mode=mode, skip_repeated=skip)
        rc, res, outs = c.classify_host(p, s, opts)
        assert rc == 0, c.error()
        assert list(res.counts) == ref["counts"] and outs == ref["outputs"]
        return c.walk_kernels()

    c = _lib.Context(0)
    c.set_debug(rows)
    for width in (128, 104, 160, 190, 700):                 # short lines from the start: the sample picks the small spans; long lines fit them too
        assert walk(c, width) == fast, width
    assert walk(c, 80) != fast                              # more than 64 lines per small span: the exact kernels
    c.close()
    c = _lib.Context(0)
    c.set_debug(rows)
    assert walk(c, 440) == fast                             # the usual geometry
    assert walk(c, 128) == fast                             # a span overflows: the call goes on with the small spans
    assert walk(c, 440) == fast
    c.close()
