"""Host-side mirror of the reference API (xenomapper_b200/xenomapper.py): the parts that stay in Python.

Mirrors the reference's own unit tests (xenomapper/tests/test_xenomapper.py, cited per test) and checks
header text against the goldens produced by the unmodified reference.
"""
import io

import pytest

from tests import _golden as G
from xenomapper_b200 import xenomapper as xm

NEG_INF = float("-inf")


def _open(key, which):
    return io.TextIOWrapper(io.BytesIO(G.fixture_bytes(key, which)))


def test_process_headers():
    """test_xenomapper.py:29-54"""
    outs = [io.StringIO() for _ in range(6)]
    xm.process_headers(_open("pe", "primary"), _open("pe", "secondary"), primary_specific=outs[0], secondary_specific=outs[1],
                       primary_multi=outs[2], secondary_multi=outs[3], unresolved=outs[5], unassigned=outs[4])
    assert [len(o.getvalue()) for o in outs] == [695, 629, 708, 642, 705, 705]


@pytest.mark.parametrize("key", ["se", "pe"])
def test_headers_match_reference_bytes(key):
    """golden full outputs = header + records; the header part must be byte-identical"""
    case = G.BY_NAME["fixture_%s_mode0_src0_skip0_min-inf" % key]
    f1, f2 = _open(key, "primary"), _open(key, "secondary")
    outs = [io.StringIO() for _ in range(6)]
    xm.process_headers(f1, f2, *outs)
    e = case["expect"]
    for b in range(6):
        hdr = outs[b].getvalue().encode()
        assert len(hdr) == e["full_len"][b] - e["records_len"][b]
    # the inputs are left at their first record (xm.py:45)
    assert f1.readline().split()[0] == f2.readline().split()[0]


def test_header_errors():
    with pytest.raises(IndexError):
        xm.get_sam_header(io.StringIO("@HD\tVN:1.0\n"))          # header only
    with pytest.raises(IndexError):
        xm.get_sam_header(io.StringIO(""))
    with pytest.raises(IndexError):
        xm.add_pg_tag([])
    assert xm.add_pg_tag(["@HD\tVN:1.0"], comment="c") == ["@HD\tVN:1.0", "@PG\tID:Xenomapper\tPN:Xenomapper\tVN:1.0.2", "@CO\tc"]
    assert xm.add_pg_tag(["@PG\tID:bowtie2\tPN:bowtie2"])[-1] == "@PG\tID:Xenomapper\tPN:Xenomapper\tPP:bowtie2\tVN:1.0.2"
    with pytest.raises(ValueError):
        xm.add_pg_tag(["@HD", "oops"])


def test_get_mapping_state():
    """test_xenomapper.py:164-188"""
    table = [((200, 199, 199, 198, NEG_INF), 'primary_specific'), ((200, 200, 199, 198, NEG_INF), 'primary_multi'),
             ((199, 198, 200, 198, NEG_INF), 'secondary_specific'), ((199, 198, 200, 200, NEG_INF), 'secondary_multi'),
             ((NEG_INF, NEG_INF, NEG_INF, NEG_INF, NEG_INF), 'unassigned'), ((200, 199, 200, 198, NEG_INF), 'unresolved'),
             ((200, 199, 199, 199, NEG_INF), 'primary_specific'), ((200, 200, 199, 199, NEG_INF), 'primary_multi'),
             ((199, 199, 200, 199, NEG_INF), 'secondary_specific'), ((199, 199, 200, 200, NEG_INF), 'secondary_multi'),
             ((9, 8, 8, 8, 10), 'unassigned'), ((200, 200, 200, 200, NEG_INF), 'unresolved'),
             ((-6, NEG_INF, NEG_INF, NEG_INF, NEG_INF), 'primary_specific'), ((NEG_INF, NEG_INF, -6, NEG_INF, NEG_INF), 'secondary_specific'),
             ((-6, NEG_INF, -2, NEG_INF, NEG_INF), 'secondary_specific'), ((0, NEG_INF, -2, NEG_INF, NEG_INF), 'primary_specific'),
             ((-2, NEG_INF, 0, NEG_INF, NEG_INF), 'secondary_specific')]
    for args, want in table:
        assert xm.get_mapping_state(*args) == want
    with pytest.raises(RuntimeError):
        xm.get_mapping_state(float("nan"), 1, 2, 3)


UNMAPPED = ['HWI-ST960:63:D0CYJACXX:4:1101:21264:2228', '4', '*', '0', '0', '*', '*', '0', '0',
            'TGGTAGTATTGGTTATGGTTCATTGTCCGGAGAGTATATTGTTGAAGAGG', 'BBCBDFDDHHHGFHHIIIIIJIJJJIGJJJGIAF:CFEGHGGHEEEG@HI', 'YT:Z:UU']
BASE = ['', '', '', '', '', '50M', '', '', '', '', '']


def test_get_tag():
    """test_xenomapper.py:190-200"""
    assert xm.get_tag(UNMAPPED, 'AS') == NEG_INF
    assert xm.get_tag(BASE + ['NM:i:0', 'AS:i:101', 'XS:i:99'], 'AS') == 101
    assert xm.get_tag(BASE + ['NM:i:0', 'AS:i:100', 'XS:i:99'], 'XS') == 99
    assert xm.get_tag(BASE + ['NM:i:0', 'AS:i:100', 'XS:i:99'], 'NM') == 0
    with pytest.raises(ValueError):
        xm.get_tag(BASE + ['AS:i:1', 'RG:Z:BASS'], 'AS')


def test_get_tag_with_ZS_as_XS():
    """test_xenomapper.py:202-212"""
    line = BASE + ['NM:i:0', 'AS:i:100', 'XS:A:+', 'ZS:i:99']
    assert xm.get_tag_with_ZS_as_XS(UNMAPPED, 'AS') == NEG_INF
    assert xm.get_tag_with_ZS_as_XS(line, 'AS') == 100
    assert xm.get_tag_with_ZS_as_XS(line, 'XS') == 99
    assert xm.get_tag_with_ZS_as_XS(line, 'NM') == 0


def test_get_cigarbased_AS_tag():
    """test_xenomapper.py:214-233"""
    def line(cigar, *tags):
        return ['', '', '', '', '', cigar, '', '', '', '', ''] + list(tags)
    assert xm.get_cigarbased_AS_tag(UNMAPPED) == NEG_INF
    table = [(line('50M', 'NM:i:0'), 0), (line('1S49M', 'NM:i:0'), -2), (line('50M', 'NM:i:2'), -12),
             (line('50M', 'NM:i:0', 'AS:i:100', 'XS:i:99'), 0), (line('10M1I39M', 'NM:i:0'), -8), (line('10M1D39M', 'NM:i:0'), -8),
             (line('10M2D38M', 'NM:i:0'), -11), (line('10M1I10M1D28M', 'NM:i:0'), -16), (line('10M1234N40M', 'NM:i:0'), 0)]
    for inp, want in table:
        assert xm.get_cigarbased_AS_tag(inp) == want
    assert xm.get_cigarbased_AS_tag(line('50M', 'NM:i:0', 'AS:i:100', 'XS:i:99'), tag='XS') == 99


def test_output_summary():
    """test_xenomapper.py:235-245"""
    out = io.StringIO()
    xm.output_summary({'foo': 1, 'bar': 101}, outfile=out)
    assert out.getvalue() == ('-' * 80 + '\nRead Count Category Summary\n\n'
                              '|       Category                                     |     Count       |\n'
                              '|:--------------------------------------------------:|:---------------:|\n'
                              '|  bar                                               |            101  |\n'
                              '|  foo                                               |              1  |\n\n')


def test_unknown_tag_func_is_rejected_not_emulated():
    rec = "r1\t0\tchr1\t1\t42\t5M\t*\t0\t0\tACGTA\tFFFFF\tAS:i:10\n"
    with pytest.raises(NotImplementedError):
        xm.main_single_end(xm.getReadPairs(io.StringIO(rec), io.StringIO(rec)), primary_specific=io.StringIO(),
                           tag_func=lambda line, tag='AS': 0.0)


def test_cli_flags_match_the_reference():
    args = xm.command_line_interface(["--primary_sam", "/dev/null", "--secondary_sam", "/dev/null", "--paired", "--conservative",
                                      "--min_score", "12.5", "--use_zs", "--cigar_scores", "--unassigned", "/dev/null"])
    assert args.paired and args.conservative and args.use_zs and args.cigar_scores and args.min_score == 12.5
    assert args.primary_specific is __import__("sys").stdout and args.unresolved is None
