#!/usr/bin/env python3
"""Generate tests/golden/golden.json by running the UNMODIFIED reference.

Run in the build container (needs /root/reference, read-only):

    python tests/golden/make_golden.py

For every case the script feeds the reference's own Python API
(process_headers / getReadPairs / main_single_end / main_paired_end /
conservative_main_paired_end, xenomapper/xenomapper.py) with in-memory text
streams and records what it produced: the six outputs (sha256 + length, with
and without the header part), the category Counter, the summary text and the
exception class if it raised.  Inputs are stored next to the answers
(fixture files gzip'd under inputs/, small adversarial cases inline, synthetic
cases by generator parameters plus an input sha256) so the tests can run where
/root/reference does not exist.
"""
import base64
import gzip
import hashlib
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("XM_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from xenomapper import xenomapper as ref      # noqa: E402  (the reference itself)
from xenomapper_b200 import synth             # noqa: E402

BINS = ("primary_specific", "secondary_specific", "primary_multi",
        "secondary_multi", "unassigned", "unresolved")
TAG_FUNCS = {0: ref.get_tag, 1: ref.get_tag_with_ZS_as_XS, 2: ref.get_cigarbased_AS_tag}
WALKS = {0: ref.main_single_end, 1: ref.main_paired_end, 2: ref.conservative_main_paired_end}


def sha(b):
    return hashlib.sha256(b).hexdigest()


def split_header(raw):
    """bytes of the leading '@' lines (what get_sam_header consumes, xm.py:36-46)"""
    pos = 0
    while pos < len(raw) and raw[pos:pos + 1] == b"@":
        nl = raw.find(b"\n", pos)
        pos = len(raw) if nl < 0 else nl + 1
    return raw[:pos], raw[pos:]


def run_reference(prim, sec, mode=0, score_src=0, skip_repeated=False, min_score=float("-inf"),
                  enabled_bins=0x3F, with_header=False):
    """prim/sec: full file bytes (header + records when with_header)."""
    f1 = io.TextIOWrapper(io.BytesIO(prim))
    f2 = io.TextIOWrapper(io.BytesIO(sec))
    outs = [io.StringIO() if enabled_bins >> b & 1 else None for b in range(6)]
    kw = dict(zip(BINS, outs))
    error = None
    counts = {}
    hdr_len = [0] * 6
    try:
        if with_header:
            # process_headers always prints to primary_specific (xm.py:151); give it one
            if kw["primary_specific"] is None:
                raise RuntimeError("golden cases with headers keep primary_specific enabled")
            ref.process_headers(f1, f2, **kw)
            hdr_len = [len(o.getvalue().encode()) if o is not None else 0 for o in outs]
        pairs = ref.getReadPairs(f1, f2, skip_repeated_reads=skip_repeated)
        if kw["primary_specific"] is None:
            kw["primary_specific"] = None
        c = WALKS[mode](pairs, min_score=min_score, tag_func=TAG_FUNCS[score_src], **kw)
        counts = dict(c)
    except Exception as e:          # noqa: BLE001 - the class name is the datum
        error = type(e).__name__
    full = [o.getvalue().encode() if o is not None else b"" for o in outs]
    recs = [full[b][hdr_len[b]:] for b in range(6)]
    summary = io.StringIO()
    if error is None:
        ref.output_summary(counts, outfile=summary)
    def key(k):
        return k if isinstance(k, str) else "|".join(k)
    return dict(error=error,
                counts={key(k): v for k, v in sorted(counts.items(), key=lambda kv: str(kv[0]))},
                records_sha256=[sha(r) for r in recs], records_len=[len(r) for r in recs],
                full_sha256=[sha(f) for f in full], full_len=[len(f) for f in full],
                summary_sha256=sha(summary.getvalue().encode()))


# --------------------------------------------------------------------------
# case construction

SEQ = "ACGTTGCAAGGCTTAACCGGTTAAGCTAGCTAGGATCCATGCATGCAAGT"
QUAL = "CCCFFFFFHHHHHJJJJJJJJJJJJJJJIJJJJJJJJIIJJJJJJJJHHF"


def rec(q, flag=0, rname="chr1", pos=100, mapq=42, cigar="50M", tags=("AS:i:100", "XN:i:0", "NM:i:0", "YT:Z:UU"),
        seq=SEQ, qual=QUAL, sep="\t"):
    f = [q, str(flag), rname, str(pos), str(mapq), cigar, "*", "0", "0", seq, qual] + list(tags)
    return sep.join(f)


def unmapped(q):
    return rec(q, 4, "*", 0, 0, "*", tags=("YT:Z:UU",))


def lines(*ls, end="\n", last=True):
    s = end.join(ls)
    return (s + (end if last else "")).encode()


def adversarial():
    """(name, prim_bytes, sec_bytes, opts) -- SURVEY.md section 9 checklist."""
    A = []

    def add(name, p, s, **o):
        A.append((name, p, s, o))

    base_p = [rec("r%d" % i, tags=("AS:i:%d" % (100 - i), "XS:i:%d" % x, "NM:i:0")) for i, x in enumerate((50, 100, 0, 99, 98))]
    base_s = [rec("r%d" % i, tags=("AS:i:%d" % a, "NM:i:1")) for i, a in enumerate((90, 99, 98, 97, 120))]
    add("basic_se", lines(*base_p), lines(*base_s))
    add("basic_se_no_trailing_newline", lines(*base_p, last=False), lines(*base_s, last=False))
    add("crlf", lines(*base_p, end="\r\n"), lines(*base_s, end="\r\n"))
    add("lone_cr", lines(*base_p, end="\r"), lines(*base_s, end="\r"), gpu="unsupported")
    add("blank_line_mid_primary", lines(base_p[0], base_p[1], "", base_p[2]), lines(*base_s))
    add("blank_line_mid_secondary", lines(*base_p), lines(base_s[0], "", base_s[1]))
    add("whitespace_only_line", lines(base_p[0], " \t ", base_p[1]), lines(*base_s))
    add("first_line_blank", lines("", *base_p), lines(*base_s))
    add("secondary_shorter", lines(*base_p), lines(*base_s[:3]))
    add("primary_shorter", lines(*base_p[:2]), lines(*base_s))
    add("empty_primary", b"", lines(*base_s))
    add("empty_both", b"", b"")
    add("only_newline", b"\n", b"\n")
    add("qname_mismatch_at_2", lines(*base_p), lines(base_s[0], base_s[1], rec("zz", tags=("AS:i:5",)), base_s[3]))
    add("qname_mismatch_at_0", lines(rec("a")), lines(rec("b")))
    add("qname_prefix_mismatch", lines(rec("read1")), lines(rec("read10")))
    # tag grammar
    add("dup_tag_substring", lines(rec("a", tags=("AS:i:5", "RG:Z:BASS"))), lines(rec("a")))
    add("tag_only_substring", lines(rec("a", tags=("RG:Z:BASS",))), lines(rec("a")))
    add("tag_substring_numeric_value", lines(rec("a", tags=("RG:Z:BASS:7",))), lines(rec("a", tags=("AS:i:6",))))
    add("as_twice_in_one_token", lines(rec("a", tags=("AS:Z:AS:9",))), lines(rec("a", tags=("AS:i:8",))))
    add("xs_token_contains_as", lines(rec("a", tags=("AS:i:9", "XS:i:AS"))), lines(rec("a")))
    add("xas_token", lines(rec("a", tags=("XAS:i:9",))), lines(rec("a", tags=("AS:i:8", "XS:i:8"))))
    add("as_in_qual_not_a_tag", lines(rec("a", qual="AS" * 25, tags=("YT:Z:UU",))), lines(rec("a", tags=("AS:i:8",))))
    add("as_in_token_10_and_11_tokens_only", lines("\t".join(["a"] + ["x"] * 9 + ["AS:i:5"])), lines(rec("a", tags=("AS:i:3",))))
    add("exactly_12_tokens", lines("\t".join(["a"] + ["x"] * 10 + ["AS:i:5"])), lines(rec("a", tags=("AS:i:3",))))
    add("few_tokens", lines("a", "b\tc"), lines("a\tzz", "b"))
    for nm, val in (("exp", "1e2"), ("decimal", "1.5"), ("inf", "inf"), ("neg_inf", "-inf"), ("nan", "nan"),
                    ("underscore", "1_0"), ("bad_underscore", "1__0"), ("plus", "+5"), ("empty", ""),
                    ("hex", "0x10"), ("neg_zero", "-0"), ("leading_zeros", "007"), ("trailing_dot", "5."),
                    ("big", "2147483648"), ("huge", "123456789012345678901234567890"), ("neg_big", "-2147483649"),
                    ("int32_min", "-2147483648"), ("int32_max", "2147483647"), ("infinity_word", "Infinity")):
        add("as_value_" + nm, lines(rec("a", tags=("AS:i:" + val, "XS:i:3"))), lines(rec("a", tags=("AS:i:50",))),
            gpu="unsupported" if nm in ("exp", "decimal", "inf", "neg_inf", "nan", "underscore", "big", "huge",
                                        "neg_big", "int32_min", "trailing_dot", "infinity_word") else None)
    add("no_colon_value", lines(rec("a", tags=("AS77",))), lines(rec("a")))
    add("xs_zero_is_absent", lines(rec("a", tags=("AS:i:0", "XS:i:0")), rec("b", tags=("AS:i:5", "XS:i:0")), rec("c", tags=("AS:i:-3", "XS:i:0"))),
        lines(rec("a", tags=("AS:i:-1",)), rec("b", tags=("AS:i:6", "XS:i:0")), rec("c", tags=("AS:i:-5",))))
    add("xs_greater_than_as", lines(rec("a", tags=("AS:i:5", "XS:i:9"))), lines(rec("a")))
    add("min_score_equal", lines(rec("a", tags=("AS:i:10",)), rec("b", tags=("AS:i:11",)), rec("c", tags=("AS:i:10",))),
        lines(rec("a", tags=("AS:i:10",)), rec("b", tags=("AS:i:10",)), rec("c", tags=("AS:i:12", "XS:i:12"))), min_score=10.0)
    add("min_score_fraction", lines(rec("a", tags=("AS:i:10",)), rec("b", tags=("AS:i:11",))),
        lines(rec("a", tags=("AS:i:11",)), rec("b", tags=("AS:i:10",))), min_score=10.5)
    add("min_score_pos_inf", lines(*base_p), lines(*base_s), min_score=float("inf"))
    # whitespace normalisation
    add("double_tab", lines(rec("a").replace("\t50M\t", "\t\t50M\t")), lines(rec("a", tags=("AS:i:1",))))
    add("leading_and_trailing_ws", lines("  " + rec("a") + "\t "), lines("\t" + rec("a", tags=("AS:i:1",))))
    add("space_separated", lines(rec("a", sep=" ")), lines(rec("a", tags=("AS:i:1",), sep="  ")))
    add("vt_ff_fs_separators", lines(rec("a").replace("\t", "\x0b", 1).replace("\t", "\x0c", 1).replace("\t", "\x1c", 1).replace("\t", "\x1f", 1)),
        lines(rec("a", tags=("AS:i:1",))))
    add("ctrl_char_in_token", lines(rec("a\x01b", tags=("AS:i:5", "X\x02:i:1"))), lines(rec("a\x01b", tags=("AS:i:4",))))
    add("nul_in_token", lines(rec("a\x00b")), lines(rec("a\x00b", tags=("AS:i:4",))))
    add("space_in_tags_changes_tokens", lines(rec("a", tags=("AS:i:66 XN:i:0", "MD:Z:50 YT:Z:UU"))), lines(rec("a", tags=("AS:i:66  XS:i:66",))))
    add("utf8_in_qname", lines(rec("réad")), lines(rec("réad", tags=("AS:i:1",))), gpu="unsupported")
    add("nbsp_separator", lines(rec("a").replace("\t", " ", 2)), lines(rec("a", tags=("AS:i:1",))), gpu="unsupported")
    add("invalid_utf8", lines(rec("a")).replace(b"chr1", b"ch\xff1"), lines(rec("a")))
    # single-end duplicates handling (CLI default for SE, xm.py:691)
    runs_p = [rec(q, tags=("AS:i:%d" % a,)) for q, a in (("a", 9), ("a", 8), ("b", 7), ("c", 6), ("c", 5), ("c", 4), ("d", 3))]
    runs_s = [rec(q, tags=("AS:i:%d" % a,)) for q, a in (("a", 1), ("b", 9), ("b", 8), ("b", 8), ("c", 2), ("d", 9), ("d", 1))]
    add("skip_repeated_runs", lines(*runs_p), lines(*runs_s), skip_repeated=True)
    add("skip_repeated_blank_inside_run", lines(runs_p[0], runs_p[1], "", runs_p[2]), lines(*runs_s), skip_repeated=True)
    add("skip_repeated_uneven_end", lines(*runs_p[:5]), lines(*runs_s), skip_repeated=True)
    add("skip_repeated_mismatch", lines(*runs_p), lines(runs_s[0], runs_s[4], runs_s[5]), skip_repeated=True)
    add("skip_repeated_paired_mode", lines(*runs_p), lines(*runs_s), skip_repeated=True, mode=1)
    # paired walks
    def pe(names, ps, ss):
        p = [rec(n, tags=("AS:i:%d" % a, "XS:i:%d" % x)) if a is not None else unmapped(n) for n, (a, x) in zip(names, ps)]
        s = [rec(n, tags=("AS:i:%d" % a, "XS:i:%d" % x)) if a is not None else unmapped(n) for n, (a, x) in zip(names, ss)]
        return lines(*p), lines(*s)
    names = ["p1", "p1", "p2", "p2", "p3", "p3", "p4", "p4", "single", "p5", "p5", "p5", "p6/1", "p6/2"]
    ps = [(100, 50), (100, 100), (90, 0), (None, 0), (80, 10), (80, 10), (None, 0), (None, 0), (70, 1), (60, 60), (60, 1), (50, 1), (40, 1), (40, 1)]
    ss = [(90, 50), (90, 10), (95, 5), (95, 95), (80, 10), (70, 70), (None, 0), (30, 30), (70, 1), (10, 1), (70, 70), (None, 0), (40, 1), (40, 1)]
    p, s = pe(names, ps, ss)
    for mode in (1, 2):
        add("paired_mixed_mode%d" % mode, p, s, mode=mode)
        add("paired_mixed_mode%d_min50" % mode, p, s, mode=mode, min_score=50.0)
        add("paired_mixed_mode%d_few_bins" % mode, p, s, mode=mode, enabled_bins=0b100101)
    add("paired_run_of_three", *pe(["x"] * 3 + ["y"] * 4, [(9, 1)] * 7, [(8, 1), (9, 1), (10, 1), (9, 9), (9, 1), (None, 0), (8, 1)]), mode=1)
    add("paired_single_record", *pe(["x"], [(9, 1)], [(8, 1)]), mode=1)
    add("se_disabled_bins_still_count", lines(*base_p), lines(*base_s), enabled_bins=0b000001)
    add("se_only_unresolved_enabled", lines(*base_p), lines(*base_s), enabled_bins=0b100000)
    # every (fwd, rev) state combination, both chains
    st = {"PS": ((100, 1), (50, 1)), "SS": ((50, 1), (100, 1)), "PM": ((100, 100), (50, 1)), "SM": ((50, 1), (100, 100)),
          "UR": ((70, 1), (70, 1)), "UA": ((None, 0), (None, 0))}
    nm, pp, sq = [], [], []
    for i, f in enumerate(st):
        for j, r in enumerate(st):
            nm += ["u%d_%d" % (i, j)] * 2
            pp += [st[f][0], st[r][0]]
            sq += [st[f][1], st[r][1]]
    p, s = pe(nm, pp, sq)
    add("paired_all_36_liberal", p, s, mode=1)
    add("paired_all_36_conservative", p, s, mode=2)
    # CIGAR + NM scoring
    cg = [("50M", "NM:i:0"), ("1S49M", "NM:i:0"), ("50M", "NM:i:2"), ("10M1I39M", "NM:i:0"), ("10M1D39M", "NM:i:0"),
          ("10M2D38M", "NM:i:0"), ("10M1I10M1D28M", "NM:i:0"), ("10M1234N40M", "NM:i:0"), ("*", "NM:i:0"), ("5H45M3S", "NM:i:1"),
          ("12Q3S4", "NM:i:0"), ("3S", "NM:i:-2"), ("0010S40M", "NM:i:+1")]
    p = [rec("c%d" % i, cigar=c, tags=("AS:i:7", n, "XS:i:3")) for i, (c, n) in enumerate(cg)]
    s = [rec("c%d" % i, cigar="48M2S", tags=("NM:i:0",)) for i in range(len(cg))]
    add("cigar_scores", lines(*p), lines(*s), score_src=2)
    add("cigar_scores_min", lines(*p), lines(*s), score_src=2, min_score=-7.0)
    add("cigar_nm_duplicate_takes_first", lines(rec("a", tags=("NM:i:1", "XNM:i:9"))), lines(rec("a", tags=("NM:i:2",))), score_src=2)
    add("cigar_nm_decimal", lines(rec("a", tags=("NM:i:1.5",))), lines(rec("a", tags=("NM:i:2",))), score_src=2)
    add("cigar_nm_missing", lines(rec("a", tags=("AS:i:5",))), lines(rec("a", tags=("NM:i:2",))), score_src=2)
    add("cigar_xs_strand_tag", lines(rec("a", tags=("NM:i:1", "XS:A:+"))), lines(rec("a", tags=("NM:i:2",))), score_src=2)
    add("cigar_paired", *[lines(*[rec(n, cigar=c, tags=("NM:i:%d" % m,)) for n, c, m in rows]) for rows in (
        [("a", "50M", 0), ("a", "45M5S", 1), ("b", "50M", 3), ("b", "50M", 0)],
        [("a", "40M10S", 0), ("a", "50M", 0), ("b", "50M", 0), ("b", "20M2I28M", 0)])], score_src=2, mode=2)
    # HISAT-style ZS
    hz_p = [rec("h%d" % i, tags=("AS:i:%d" % a,) + (("ZS:i:%d" % z,) if z is not None else ()) + ("XS:A:+", "NH:i:1"))
            for i, (a, z) in enumerate(((0, 0), (0, -5), (-5, -5), (-18, None), (-19, None), (-3, 0)))]
    hz_s = [rec("h%d" % i, tags=("AS:i:%d" % a, "XS:A:-")) for i, a in enumerate((-1, 0, -6, -18, -18, -3))]
    add("hisat_zs", lines(*hz_p), lines(*hz_s), score_src=1)
    add("hisat_zs_min18", lines(*hz_p), lines(*hz_s), score_src=1, min_score=-18.0)
    add("hisat_without_use_zs", lines(*hz_p), lines(*hz_s), score_src=0)
    # geometry: long and short lines
    long_seq = "ACGT" * 30000
    add("very_long_lines", lines(rec("L1", seq=long_seq, qual="J" * len(long_seq)), rec("L2"), rec("L3", seq=long_seq[:70000], qual="F" * 70000, tags=("AS:i:9", "XS:i:9"))),
        lines(rec("L1", tags=("AS:i:150",)), rec("L2", seq=long_seq, qual="#" * len(long_seq), tags=("AS:i:200", "XS:i:1")), rec("L3")))
    add("many_short_lines", lines(*["q%d" % i for i in range(5000)]), lines(*["q%d\tx" % i for i in range(5000)]))
    add("many_short_lines_paired", lines(*["q%d" % (i // 2) for i in range(5000)]), lines(*["q%d\tx" % (i // 2) for i in range(5000)]), mode=1)
    return A


def main():
    cases = []
    os.makedirs(os.path.join(HERE, "inputs"), exist_ok=True)
    data = os.path.join(REF, "xenomapper", "tests", "data")
    fixtures = {"se": ("test_human_in.sam", "test_mouse_in.sam"),
                "pe": ("paired_end_testdata_human.sam", "paired_end_testdata_mouse.sam")}
    for key, (a, b) in fixtures.items():
        for tag, fn in (("primary", a), ("secondary", b)):
            raw = open(os.path.join(data, fn), "rb").read()
            with gzip.GzipFile(os.path.join(HERE, "inputs", "fixture_%s_%s.sam.gz" % (key, tag)), "wb", mtime=0) as g:
                g.write(raw)
    for key in ("bam",):
        for tag, fn in (("primary", "paired_end_testdata_human.bam"), ("secondary", "paired_end_testdata_mouse.bam")):
            raw = open(os.path.join(data, fn), "rb").read()
            with gzip.GzipFile(os.path.join(HERE, "inputs", "fixture_pe_%s.bam.gz" % tag), "wb", mtime=0) as g:
                g.write(raw)

    def fixture(key):
        return [open(os.path.join(data, f), "rb").read() for f in fixtures[key]]

    # 1. the reference's own fixtures x every walk / score source / skip flag
    for key in ("se", "pe"):
        p, s = fixture(key)
        for mode in (0, 1, 2):
            for score_src in (0, 1, 2):
                for skip in (False, True):
                    for ms in (float("-inf"), 60.0):
                        if ms != float("-inf") and (score_src or skip):
                            continue
                        o = dict(mode=mode, score_src=score_src, skip_repeated=skip, min_score=ms, enabled_bins=0x3F)
                        e = run_reference(p, s, with_header=True, **o)
                        cases.append(dict(name="fixture_%s_mode%d_src%d_skip%d_min%s" % (key, mode, score_src, int(skip), ms),
                                          input=dict(kind="fixture", key=key), header=True, opts=o, expect=e, gpu=None))
    # the three known answers of test_xenomapper.py:93/:125/:158 use exactly these calls
    # 2. adversarial corpus (no headers)
    for name, p, s, o in adversarial():
        gpu = o.pop("gpu", None)
        opts = dict(mode=0, score_src=0, skip_repeated=False, min_score=float("-inf"), enabled_bins=0x3F)
        opts.update(o)
        e = run_reference(p, s, **opts)
        inp = dict(kind="inline", prim=base64.b64encode(gzip.compress(p, mtime=0)).decode(),
                   sec=base64.b64encode(gzip.compress(s, mtime=0)).decode())
        cases.append(dict(name="adv_" + name, input=inp, header=False, opts=opts, expect=e, gpu=gpu))
    # 3. seeded synthetic pairs (SURVEY 8d shapes), all walks
    for style, n, variants in (
            (synth.STYLE_SE_BOWTIE2, 30000, [dict(mode=0), dict(mode=0, skip_repeated=True), dict(mode=0, min_score=150.0),
                                             dict(mode=0, score_src=2), dict(mode=0, enabled_bins=0b010101)]),
            (synth.STYLE_PE_BOWTIE2, 30000, [dict(mode=1), dict(mode=2), dict(mode=1, min_score=200.0), dict(mode=2, score_src=2),
                                             dict(mode=0, skip_repeated=True), dict(mode=0)]),
            (synth.STYLE_PE_HISAT, 30000, [dict(mode=2, score_src=1, min_score=-18.0), dict(mode=1, score_src=1),
                                           dict(mode=2, score_src=0)])):
        for seed in (1, 7):
            p, s = synth.generate(n, seed=seed, style=style)
            p, s = p.tobytes(), s.tobytes()
            for v in variants:
                opts = dict(mode=0, score_src=0, skip_repeated=False, min_score=float("-inf"), enabled_bins=0x3F)
                opts.update(v)
                e = run_reference(p, s, **opts)
                cases.append(dict(name="synth_style%d_seed%d_%s" % (style, seed, "_".join("%s%s" % kv for kv in sorted(v.items()))),
                                  input=dict(kind="synth", style=style, seed=seed, n=n, prim_sha256=sha(p), sec_sha256=sha(s)),
                                  header=False, opts=opts, expect=e, gpu=None))

    def enc(o):
        if isinstance(o, float) and o in (float("inf"), float("-inf")):
            return "inf" if o > 0 else "-inf"
        return o
    for c in cases:
        c["opts"]["min_score"] = enc(c["opts"]["min_score"])
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(dict(reference_version=ref.__version__, cases=cases), f, indent=0, sort_keys=True)
    errs = {}
    for c in cases:
        errs[c["expect"]["error"]] = errs.get(c["expect"]["error"], 0) + 1
    print("wrote %d cases; outcomes: %s" % (len(cases), errs))


if __name__ == "__main__":
    main()
