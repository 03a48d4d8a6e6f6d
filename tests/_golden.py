"""Loader for tests/golden/golden.json (made by tests/golden/make_golden.py)."""
import base64
import gzip
import hashlib
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
BINS = ("primary_specific", "secondary_specific", "primary_multi",
        "secondary_multi", "unassigned", "unresolved")
ERR_CODE = {None: 0, "AssertionError": 1, "ValueError": 2, "RuntimeError": 3, "UnicodeDecodeError": 4}

with open(os.path.join(GOLD, "golden.json")) as _f:
    _DOC = json.load(_f)
CASES = _DOC["cases"]
for _c in CASES:
    _m = _c["opts"]["min_score"]
    _c["opts"]["min_score"] = float(_m)
BY_NAME = {c["name"]: c for c in CASES}
_cache = {}


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def fixture_bytes(key, which, ext="sam"):
    with gzip.open(os.path.join(GOLD, "inputs", "fixture_%s_%s.%s.gz" % (key, which, ext)), "rb") as g:
        return g.read()


def split_header(raw):
    pos = 0
    while pos < len(raw) and raw[pos:pos + 1] == b"@":
        nl = raw.find(b"\n", pos)
        pos = len(raw) if nl < 0 else nl + 1
    return raw[:pos], raw[pos:]


def case_files(case):
    """full file bytes (header included for fixture cases) of (primary, secondary)"""
    inp = case["input"]
    k = json.dumps(inp, sort_keys=True)
    if k in _cache:
        return _cache[k]
    if inp["kind"] == "fixture":
        r = fixture_bytes(inp["key"], "primary"), fixture_bytes(inp["key"], "secondary")
    elif inp["kind"] == "inline":
        r = gzip.decompress(base64.b64decode(inp["prim"])), gzip.decompress(base64.b64decode(inp["sec"]))
    else:
        from xenomapper_b200 import synth
        p, s = synth.generate(inp["n"], seed=inp["seed"], style=inp["style"])
        r = p.tobytes(), s.tobytes()
        assert sha(r[0]) == inp["prim_sha256"] and sha(r[1]) == inp["sec_sha256"], \
            "synthetic generator drifted: regenerate goldens (tests/golden/make_golden.py)"
    if len(_cache) > 8:
        _cache.clear()
    _cache[k] = r
    return r


def case_records(case):
    """record regions (headers removed) of (primary, secondary)"""
    p, s = case_files(case)
    if case["header"]:
        return split_header(p)[1], split_header(s)[1]
    return p, s


def counts_dict(counts36, mode):
    """36-slot histogram -> the reference's Counter rendering used in golden.json"""
    out = {}
    if mode == 0:
        for i in range(6):
            if counts36[i]:
                out[BINS[i]] = counts36[i]
    else:
        for f in range(6):
            for r in range(6):
                if counts36[f * 6 + r]:
                    out[BINS[f] + "|" + BINS[r]] = counts36[f * 6 + r]
    return out
