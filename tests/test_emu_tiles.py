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
