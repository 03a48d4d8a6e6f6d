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


@pytest.mark.parametrize("debug", [0, 1, 2, 3], ids=["big", "big_generic", "small", "small_generic"])
@pytest.mark.parametrize("case", G.CASES, ids=[c["name"] for c in G.CASES])
def test_cuda_matches_reference_golden(ctx, case, debug):
    if debug >= 2 and case["input"]["kind"] == "synth" and case["input"]["seed"] != 1:
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


@pytest.mark.parametrize("style,mode,skip", [(0, 0, True), (1, 1, False), (2, 2, False)])
def test_cuda_large_synthetic_matches_oracle(ctx, style, mode, skip):
    """300 k records (about 130 MB per stream: thousands of tiles, look-back across several waves of CTAs)"""
    from oracle import oracle
    from xenomapper_b200 import _lib, synth
    ctx.set_debug(0)
    p, s = synth.generate(300000, seed=77, style=style)
    score = 1 if style == 2 else 0
    ref = oracle.classify(p, s, mode=mode, score_src=score, skip_repeated=skip, min_score=-18.0 if style == 2 else float("-inf"))
    rc, res, outs = ctx.classify_host(p, s, _lib.Context.opts(mode, score, skip, -18.0 if style == 2 else float("-inf")))
    assert rc == 0, ctx.error()
    assert list(res.counts) == ref["counts"]
    assert outs == ref["outputs"]
