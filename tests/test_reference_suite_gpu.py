"""The reference's own known-answer walk tests (xenomapper/tests/test_xenomapper.py:56-161), run against
xenomapper_b200.xenomapper: same calls, same assertions, the walk on the GPU."""
import hashlib
import io

import pytest

from tests import _golden as G

pytestmark = pytest.mark.gpu


def _open(key, which):
    return io.TextIOWrapper(io.BytesIO(G.fixture_bytes(key, which)))


def _six():
    return dict(primary_specific=io.StringIO(), secondary_specific=io.StringIO(), primary_multi=io.StringIO(),
                secondary_multi=io.StringIO(), unresolved=io.StringIO(), unassigned=io.StringIO())


def nlines(f):
    return len(f.getvalue().split('\n'))


def test_consistent_output_SE():
    """test_xenomapper.py:56-96"""
    from xenomapper_b200.xenomapper import process_headers, main_single_end, getReadPairs
    o = _six()
    sam1, sam2 = _open("se", "primary"), _open("se", "secondary")
    process_headers(sam1, sam2, **o)
    cat_counts = main_single_end(getReadPairs(sam1, sam2), **o)
    for k in ("primary_specific", "primary_multi", "secondary_specific", "secondary_multi", "unassigned"):
        assert cat_counts[k] == nlines(o[k]) - 4
    assert hashlib.sha224(o["primary_specific"].getvalue().encode('latin-1')).hexdigest() == \
        '381325b12dd9a9cd3afdd72eeb16b23cc92ddd16f675bb21bb21e08e'


def test_consistent_output_PE():
    """test_xenomapper.py:98-128"""
    from xenomapper_b200.xenomapper import process_headers, main_paired_end, getReadPairs
    o = _six()
    sam1, sam2 = _open("pe", "primary"), _open("pe", "secondary")
    process_headers(sam1, sam2, primary_specific=o["primary_specific"], secondary_specific=o["secondary_specific"])
    c = main_paired_end(getReadPairs(sam1, sam2), **o)
    assert sum(c[x] for x in c if 'primary_specific' in x) * 2 == nlines(o["primary_specific"]) - 30
    assert sum(c[x] for x in c if 'secondary_specific' in x and 'primary_specific' not in x) * 2 == nlines(o["secondary_specific"]) - 27
    assert sum(c[x] for x in c if 'primary_multi' in x and 'primary_specific' not in x and 'secondary_specific' not in x) * 2 == \
        nlines(o["primary_multi"]) - 1
    assert sum(c[x] for x in c if 'secondary_multi' in x and 'primary_multi' not in x and 'primary_specific' not in x
               and 'secondary_specific' not in x) * 2 == nlines(o["secondary_multi"]) - 1
    assert hashlib.sha224(o["primary_specific"].getvalue().encode('latin-1')).hexdigest() == \
        '64c0e24bf141c5aa3bb0993c73b34cdfe630a504ac424843f746918d'


def test_consistent_output_conservative_PE():
    """test_xenomapper.py:130-161"""
    from xenomapper_b200.xenomapper import process_headers, conservative_main_paired_end, getReadPairs
    o = _six()
    sam1, sam2 = _open("pe", "primary"), _open("pe", "secondary")
    process_headers(sam1, sam2, primary_specific=o["primary_specific"], secondary_specific=o["secondary_specific"])
    c = conservative_main_paired_end(getReadPairs(sam1, sam2), **o)
    assert c[('primary_specific', 'secondary_specific')] * 4 + c[('unresolved', 'unresolved')] * 4 == nlines(o["unresolved"]) - 1
    assert c[('primary_multi', 'primary_multi')] * 2 == nlines(o["primary_multi"]) - 1
    assert c[('secondary_multi', 'secondary_multi')] * 2 == nlines(o["secondary_multi"]) - 1
    assert (c[('primary_specific', 'primary_specific')] + c[('primary_specific', 'primary_multi')] +
            c[('primary_multi', 'primary_specific')]) * 2 == nlines(o["primary_specific"]) - 30
    assert (c[('secondary_specific', 'secondary_specific')] + c[('secondary_specific', 'secondary_multi')] +
            c[('secondary_multi', 'secondary_specific')]) * 2 == nlines(o["secondary_specific"]) - 27
    assert c[('unassigned', 'unassigned')] * 2 == nlines(o["unassigned"]) - 1
    assert hashlib.sha224(o["primary_specific"].getvalue().encode('latin-1')).hexdigest() == \
        'c4de3de755092c8f9ff1eb2cd360a502d74ebd4c1e65ed282515ed3e'


@pytest.mark.parametrize("name", ["fixture_se_mode0_src0_skip0_min-inf", "fixture_pe_mode1_src0_skip0_min-inf",
                                  "fixture_pe_mode2_src0_skip0_min-inf", "fixture_pe_mode1_src2_skip0_min-inf",
                                  "fixture_se_mode0_src0_skip1_min-inf", "fixture_pe_mode2_src0_skip0_min60.0"])
def test_full_files_headers_included(name):
    """six complete output files (header + records) and the summary text, byte for byte as the reference CLI writes them"""
    from xenomapper_b200 import xenomapper as xm
    case = G.BY_NAME[name]
    key, o = case["input"]["key"], case["opts"]
    outs = [io.StringIO() for _ in range(6)]
    f1, f2 = _open(key, "primary"), _open(key, "secondary")
    xm.process_headers(f1, f2, *outs)
    walk = (xm.main_single_end, xm.main_paired_end, xm.conservative_main_paired_end)[o["mode"]]
    tag = (xm.get_tag, xm.get_tag_with_ZS_as_XS, xm.get_cigarbased_AS_tag)[o["score_src"]]
    counts = walk(xm.getReadPairs(f1, f2, skip_repeated_reads=o["skip_repeated"]), *outs, min_score=o["min_score"], tag_func=tag)
    e = case["expect"]
    assert [G.sha(x.getvalue().encode()) for x in outs] == e["full_sha256"]
    summary = io.StringIO()
    xm.output_summary(counts, outfile=summary)
    assert G.sha(summary.getvalue().encode()) == e["summary_sha256"]


def test_cli_end_to_end(tmp_path):
    """the console entry point on real files, stdout default for primary_specific (xm.py:618), run as a process"""
    import os
    import subprocess
    import sys
    p, s = tmp_path / "h.sam", tmp_path / "m.sam"
    p.write_bytes(G.fixture_bytes("pe", "primary")); s.write_bytes(G.fixture_bytes("pe", "secondary"))
    un = tmp_path / "unresolved.sam"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "xenomapper_b200.xenomapper", "--primary_sam", str(p), "--secondary_sam", str(s),
                        "--paired", "--conservative", "--unresolved", str(un)], cwd=root, capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()
    e = G.BY_NAME["fixture_pe_mode2_src0_skip0_min-inf"]["expect"]
    assert G.sha(r.stdout) == e["full_sha256"][0]
    assert G.sha(un.read_bytes()) == e["full_sha256"][5]
    assert G.sha(r.stderr) == e["summary_sha256"]


@pytest.mark.parametrize("name,flags", [("fixture_se_mode0_src0_skip1_min-inf", []), ("fixture_pe_mode1_src0_skip0_min-inf", ["--paired"]),
                                        ("fixture_pe_mode1_src2_skip0_min-inf", ["--paired", "--cigar_scores"])])
@pytest.mark.parametrize("chunk", [None, "3000"])
def test_cli_on_real_files_streams_through_descriptors(tmp_path, name, flags, chunk):
    """every input and output a regular file: the walk goes through xm_classify_fds (pinned staging, write(2) per bin);
    with XM_CHUNK_BYTES=3000 in dozens of steps.  Expected: the reference CLI's six files, headers included."""
    import os
    import subprocess
    import sys
    if name not in G.BY_NAME:
        pytest.skip("no such golden")
    case = G.BY_NAME[name]
    key = case["input"]["key"]
    p, s = tmp_path / "h.sam", tmp_path / "m.sam"
    p.write_bytes(G.fixture_bytes(key, "primary")); s.write_bytes(G.fixture_bytes(key, "secondary"))
    outs = [tmp_path / (b + ".sam") for b in G.BINS]
    cmd = [sys.executable, "-m", "xenomapper_b200.xenomapper", "--primary_sam", str(p), "--secondary_sam", str(s)] + flags
    for b, o in zip(G.BINS, outs):
        cmd += ["--" + b, str(o)]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, **({"XM_CHUNK_BYTES": chunk} if chunk else {}))
    r = subprocess.run(cmd, cwd=root, capture_output=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr.decode()
    e = case["expect"]
    assert [G.sha(o.read_bytes()) for o in outs] == e["full_sha256"]
    assert G.sha(r.stderr) == e["summary_sha256"]


@pytest.mark.parametrize("name,flags", [("fixture_se_mode0_src0_skip1_min-inf", []), ("fixture_pe_mode1_src0_skip0_min-inf", ["--paired"])])
@pytest.mark.parametrize("chunk", [None, "3000"])
def test_cli_bgzf_outputs_inflate_to_the_reference_files(tmp_path, name, flags, chunk):
    """--bgzf (additive): every output is BGZF; gunzip gives the reference CLI's file, header included, and
    the file ends with the 28-byte end-of-file member"""
    import gzip
    import os
    import subprocess
    import sys
    if name not in G.BY_NAME:
        pytest.skip("no such golden")
    case = G.BY_NAME[name]
    key = case["input"]["key"]
    p, s = tmp_path / "h.sam", tmp_path / "m.sam"
    p.write_bytes(G.fixture_bytes(key, "primary")); s.write_bytes(G.fixture_bytes(key, "secondary"))
    outs = [tmp_path / (b + ".sam.gz") for b in G.BINS]
    cmd = [sys.executable, "-m", "xenomapper_b200.xenomapper", "--primary_sam", str(p), "--secondary_sam", str(s), "--bgzf"] + flags
    for b, o in zip(G.BINS, outs):
        cmd += ["--" + b, str(o)]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, **({"XM_CHUNK_BYTES": chunk} if chunk else {}))
    r = subprocess.run(cmd, cwd=root, capture_output=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr.decode()
    e = case["expect"]
    raw = [o.read_bytes() for o in outs]
    assert all(x.endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")) for x in raw)
    assert [G.sha(gzip.decompress(x)) for x in raw] == e["full_sha256"]
    assert G.sha(r.stderr) == e["summary_sha256"]


@pytest.mark.parametrize("name,flags", [("fixture_se_mode0_src0_skip1_min-inf", []), ("fixture_pe_mode1_src0_skip0_min-inf", ["--paired"])])
@pytest.mark.parametrize("chunk", [None, "3000"])
def test_cli_reads_pipes(tmp_path, name, flags, chunk):
    """both inputs are named pipes fed by `cat` (the reference needs seekable files): headers and walk
    happen inside xm_classify_streams on the descriptors; the six files equal the reference CLI's on the same text"""
    import os
    import subprocess
    import sys
    if name not in G.BY_NAME:
        pytest.skip("no such golden")
    case = G.BY_NAME[name]
    key = case["input"]["key"]
    p, s = tmp_path / "h.sam", tmp_path / "m.sam"
    p.write_bytes(G.fixture_bytes(key, "primary")); s.write_bytes(G.fixture_bytes(key, "secondary"))
    fp, fs = str(tmp_path / "p.fifo"), str(tmp_path / "s.fifo")
    os.mkfifo(fp); os.mkfifo(fs)
    outs = [tmp_path / (b + ".sam") for b in G.BINS]
    cmd = [sys.executable, "-m", "xenomapper_b200.xenomapper", "--primary_sam", fp, "--secondary_sam", fs] + flags
    for b, o in zip(G.BINS, outs):
        cmd += ["--" + b, str(o)]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, **({"XM_CHUNK_BYTES": chunk} if chunk else {}))
    feeders = [subprocess.Popen("cat %s > %s" % (src, dst), shell=True) for src, dst in ((p, fp), (s, fs))]
    r = subprocess.run(cmd, cwd=root, capture_output=True, timeout=300, env=env)
    [f.wait(timeout=60) for f in feeders]
    assert r.returncode == 0, r.stderr.decode()
    e = case["expect"]
    assert [G.sha(o.read_bytes()) for o in outs] == e["full_sha256"]
    assert G.sha(r.stderr) == e["summary_sha256"]


@pytest.mark.parametrize("bgzf", [False, True], ids=["sam_out", "bgzf_out"])
@pytest.mark.parametrize("chunk", [None, "3000"])
def test_cli_bam_inputs_to_files(tmp_path, bgzf, chunk):
    """--primary_bam / --secondary_bam (xm.py:703-713) with real output files: the BAM files are mapped, inflated and
    rendered on the GPU and the bins go to the descriptors (xm_classify_bam_fds) -- with --bgzf deflated on the GPU too.
    The six files are the reference CLI's on the SAM twins, headers included."""
    import gzip
    import os
    import subprocess
    import sys
    name = "fixture_pe_mode1_src0_skip0_min-inf"
    if name not in G.BY_NAME:
        pytest.skip("no such golden")
    case = G.BY_NAME[name]
    p, s = tmp_path / "h.bam", tmp_path / "m.bam"
    p.write_bytes(G.fixture_bytes("pe", "primary", "bam")); s.write_bytes(G.fixture_bytes("pe", "secondary", "bam"))
    outs = [tmp_path / (b + (".sam.gz" if bgzf else ".sam")) for b in G.BINS]
    cmd = [sys.executable, "-m", "xenomapper_b200.xenomapper", "--primary_bam", str(p), "--secondary_bam", str(s), "--paired"] + (["--bgzf"] if bgzf else [])
    for b, o in zip(G.BINS, outs):
        cmd += ["--" + b, str(o)]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, **({"XM_CHUNK_BYTES": chunk, "XM_BAM_WINDOW": "70000"} if chunk else {}))
    r = subprocess.run(cmd, cwd=root, capture_output=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr.decode()
    e = case["expect"]
    raw = [o.read_bytes() for o in outs]
    if bgzf:
        raw = [gzip.decompress(x) for x in raw]
    assert [G.sha(x) for x in raw] == e["full_sha256"]
    assert G.sha(r.stderr) == e["summary_sha256"]
